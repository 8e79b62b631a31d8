"""Python host mirror of the reference's inversion entry point (model/samodel.h:8-19).

``Inverter`` owns one device context of libphotic_b200.so; ``samodel()`` keeps the reference's
argument names and meaning (scene_data / gridded_data / scene_indexes ... -> the ten output grids),
so the parity tests read like calls into the reference. All compute happens in the CUDA library;
this module only marshals buffers. torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import Outputs, SceneDesc, Stats, SCALAR_PLANES, check


def _np_ptr(a, t):
    return a.ctypes.data_as(t)


class Inverter:
    """One per GPU. ``device`` is the CUDA ordinal. Raises PhoticError when no device is usable."""

    def __init__(self, device: int = 0):
        self.lib = capi.lib()
        self.ctx = C.c_void_p()
        check(self.lib.phb_ctx_create(int(device), C.byref(self.ctx)))
        self.device = device

    def close(self):
        if self.ctx:
            self.lib.phb_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers (what the samodel() shim uses) ---------------------------------------------
    def _host_outputs(self, desc: SceneDesc, want_scene_planes: bool, buffers: dict | None):
        nrows, ncols, ns = desc.nrows, desc.ncols, desc.n_scenes
        mb = max(desc.n_bands[s] for s in range(ns))
        out = buffers if buffers is not None else {}
        o = Outputs()
        for name in SCALAR_PLANES:
            if name not in out:
                out[name] = np.zeros((nrows, ncols), dtype=np.float32)
            setattr(o, name, _np_ptr(out[name], capi._fp))
        if want_scene_planes:
            shapes = {"K": (ns, mb, nrows, ncols), "P": (ns, nrows, ncols), "G": (ns, nrows, ncols),
                      "X": (ns, nrows, ncols)}
            for name, shp in shapes.items():
                if name not in out:
                    out[name] = np.zeros(shp, dtype=np.float32)
                setattr(o, name, _np_ptr(out[name], capi._fp))
        if "converged" not in out:
            out["converged"] = np.zeros((nrows, ncols), dtype=np.uint8)
        if "n_evals" not in out:
            out["n_evals"] = np.zeros((nrows, ncols), dtype=np.int32)
        o.converged = _np_ptr(out["converged"], capi._u8p)
        o.n_evals = _np_ptr(out["n_evals"], capi._ip)
        return o, out

    @staticmethod
    def _plane_ptrs(planes):
        """planes: ndarray [n_planes, nrows, ncols] float32 (C order) or a list of 2-D arrays."""
        if isinstance(planes, np.ndarray):
            assert planes.dtype == np.float32 and planes.flags["C_CONTIGUOUS"]
            lst = [planes[g] for g in range(planes.shape[0])]
        else:
            lst = [np.ascontiguousarray(p, dtype=np.float32) for p in planes]
        arr = (C.c_void_p * len(lst))(*[p.ctypes.data for p in lst])
        return arr, lst

    def invert_host(self, desc: SceneDesc, planes, prior, row_begin=0, row_end=None, scene_planes=True,
                    buffers=None, debug=False):
        """Inverts rows [row_begin,row_end) from host arrays; returns (outputs dict, stats dict)."""
        row_end = desc.nrows if row_end is None else row_end
        ptrs, keep = self._plane_ptrs(planes)
        pr = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32)
        o, out = self._host_outputs(desc, scene_planes, buffers)
        st = Stats()
        prp = C.c_void_p(None) if pr is None else C.c_void_p(pr.ctypes.data)
        if not debug:
            check(self.lib.phb_invert_host(self.ctx, C.byref(desc), ptrs, prp, row_begin, row_end, C.byref(o), C.byref(st)))
        else:
            cap = (row_end - row_begin) * desc.ncols
            rl = self.lib.phb_debug_record_len(C.byref(desc))
            rec = np.zeros((cap, rl))
            pix = np.full(cap, -1, dtype=np.int32)
            it = np.zeros((cap, 3), dtype=np.int32)
            check(self.lib.phb_invert_host_debug(self.ctx, C.byref(desc), ptrs, prp, row_begin, row_end, C.byref(o),
                                                 _np_ptr(rec, capi._dp), _np_ptr(pix, capi._ip), _np_ptr(it, capi._ip),
                                                 cap, C.byref(st)))
            n = int(st.n_valid)
            order = np.argsort(pix[:n], kind="stable")
            out["rec"], out["pix"] = rec[:n][order], pix[:n][order]
            out["rec_evals"] = it[:n, 0][order]
            out["rec_converged"] = (it[:n, 1] & 1)[order]
            out["rec_iters"] = (it[:n, 1] >> 1)[order]
            out["rec_restarts"] = it[:n, 2][order]  # nelmin's numres, summed over the H starts
        return out, st.as_dict()

    @staticmethod
    def plan_row_bands_host(desc: SceneDesc, planes, prior, n_parts: int):
        """phb_plan_row_bands: cost-balanced contiguous row bands (host only, no device). Returns (edges, row_cost)."""
        ptrs, keep = Inverter._plane_ptrs(planes)
        pr = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32)
        edges = np.zeros(n_parts + 1, dtype=np.int32)
        cost = np.zeros(desc.nrows, dtype=np.float64)
        check(capi.lib().phb_plan_row_bands(C.byref(desc), ptrs, C.c_void_p(None) if pr is None else C.c_void_p(pr.ctypes.data),
                                            n_parts, _np_ptr(edges, C.POINTER(C.c_int32)), _np_ptr(cost, capi._dp)))
        return edges, cost

    @staticmethod
    def invert_host_multi(inverters, desc: SceneDesc, planes, prior, scene_planes=True, buffers=None):
        """phb_invert_host_multi: one process, one Inverter (context) per device, cost-balanced row bands, one host
        thread per band inside the library. Returns (outputs dict, stats dict incl. 'edges' and 'per_ctx')."""
        n = len(inverters)
        ptrs, keep = Inverter._plane_ptrs(planes)
        pr = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32)
        o, out = inverters[0]._host_outputs(desc, scene_planes, buffers)
        ctxs = (C.c_void_p * n)(*[iv.ctx.value for iv in inverters])
        st, per = Stats(), (Stats * n)()
        edges = np.zeros(n + 1, dtype=np.int32)
        check(capi.lib().phb_invert_host_multi(ctxs, n, C.byref(desc), ptrs,
                                               C.c_void_p(None) if pr is None else C.c_void_p(pr.ctypes.data), C.byref(o),
                                               C.byref(st), per, _np_ptr(edges, C.POINTER(C.c_int32))))
        d = st.as_dict()
        d["edges"] = edges
        d["per_ctx"] = [per[k].as_dict() for k in range(n)]
        return out, d

    def depth_sigma_host(self, desc: SceneDesc, planes, prior, depth_neg, seed: int, n_samples: int = 128,
                         chain_mode: int = 1, max_intervals: int = 120):
        """Depth-error estimate of samodel() (samodel.c:1376-1477; phb_depth_sigma_host). ``depth_neg`` is the
        depth plane as invert_* leaves it (negated). Returns (depth_sigma plane, table, trials, stats)."""
        ptrs, keep = self._plane_ptrs(planes)
        pr = np.ascontiguousarray(prior, dtype=np.float32)
        dn = np.ascontiguousarray(depth_neg, dtype=np.float32)
        sig = np.zeros((desc.nrows, desc.ncols), dtype=np.float32)
        table = np.zeros(120)
        trials = np.zeros((120, n_samples))
        nint = C.c_int32(0)
        st = Stats()
        check(self.lib.phb_depth_sigma_host(self.ctx, C.byref(desc), ptrs, C.c_void_p(pr.ctypes.data),
                                            C.c_void_p(dn.ctypes.data), C.c_uint(seed), n_samples, chain_mode,
                                            max_intervals, _np_ptr(sig, capi._fp), _np_ptr(table, capi._dp),
                                            C.byref(nint), _np_ptr(trials, capi._dp), C.byref(st)))
        return sig, table[:nint.value], trials[:nint.value], st.as_dict()

    # ---- device buffers (resident data: bench `value`, multi-GPU shards) ---------------------------
    def invert_device(self, desc: SceneDesc, planes, prior, outputs: dict, row_begin=0, row_end=None, stream=None):
        """planes / prior / outputs are torch CUDA tensors on this device. Returns stats dict."""
        import torch

        row_end = desc.nrows if row_end is None else row_end
        o = Outputs()
        for name in SCALAR_PLANES + ("K", "P", "G", "X"):
            t = outputs.get(name)
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                setattr(o, name, C.cast(t.data_ptr(), capi._fp))
        if outputs.get("converged") is not None:
            o.converged = C.cast(outputs["converged"].data_ptr(), capi._u8p)
        if outputs.get("n_evals") is not None:
            o.n_evals = C.cast(outputs["n_evals"].data_ptr(), capi._ip)
        st = Stats()
        s = torch.cuda.current_stream(planes.device) if stream is None else stream
        assert planes.is_cuda and planes.dtype == torch.float32 and planes.is_contiguous()
        prp = C.c_void_p(None) if prior is None else C.c_void_p(prior.data_ptr())
        check(self.lib.phb_invert_device(self.ctx, C.byref(desc), C.c_void_p(planes.data_ptr()), prp, row_begin, row_end,
                                         C.byref(o), C.c_void_p(s.cuda_stream), C.byref(st)))
        return st.as_dict()

    @staticmethod
    def alloc_device_outputs(desc: SceneDesc, device, scene_planes=True):
        import torch

        nrows, ncols, ns = desc.nrows, desc.ncols, desc.n_scenes
        mb = max(desc.n_bands[s] for s in range(ns))
        out = {n: torch.zeros((nrows, ncols), dtype=torch.float32, device=device) for n in SCALAR_PLANES}
        if scene_planes:
            out["K"] = torch.zeros((ns, mb, nrows, ncols), dtype=torch.float32, device=device)
            for n in ("P", "G", "X"):
                out[n] = torch.zeros((ns, nrows, ncols), dtype=torch.float32, device=device)
        out["converged"] = torch.zeros((nrows, ncols), dtype=torch.uint8, device=device)
        out["n_evals"] = torch.zeros((nrows, ncols), dtype=torch.int32, device=device)
        return out

    # ---- known-answer hooks --------------------------------------------------------------------------
    def kat_objective(self, desc: SceneDesc, nb_active, n_regions, origin, meas, params):
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        nvec, npar = params.shape
        out = np.zeros((nvec, 6))
        check(self.lib.phb_kat_objective(self.ctx, C.byref(desc), nb_active, n_regions, origin, _np_ptr(meas, capi._dp),
                                         npar, nvec, _np_ptr(params, capi._dp), _np_ptr(out, capi._dp)))
        return out

    def eval_bench(self, desc: SceneDesc, nb_active, n_regions, origin, meas, params, team_warps=1, same_smsp=False, reps=1,
                   skew_cycles=0):
        """Mapping study (phb_eval_bench): closed-loop objective evaluations on every SM with one pixel per team of
        `team_warps` warps (1 = the product's warp-per-pixel objective). Returns (first value, evaluations/s, ms)."""
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
        first, rate, ms = C.c_double(0.0), C.c_double(0.0), C.c_float(0.0)
        check(self.lib.phb_eval_bench(self.ctx, C.byref(desc), nb_active, n_regions, origin, _np_ptr(meas, capi._dp),
                                      params.size, _np_ptr(params, capi._dp), team_warps, 1 if same_smsp else 0,
                                      int(skew_cycles), reps,
                                      C.byref(first), C.byref(rate), C.byref(ms)))
        return first.value, rate.value, ms.value

    def kat_math(self, fn: int, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        yp = C.cast(None, capi._dp)
        if y is not None:
            y = np.ascontiguousarray(y, dtype=np.float64)
            yp = _np_ptr(y, capi._dp)
        check(self.lib.phb_kat_math(self.ctx, fn, _np_ptr(x, capi._dp), yp, x.size, _np_ptr(out, capi._dp)))
        return out

    def lee_ls8_host(self, mode: int, coastal, blue, green, red, spv, theta_s: float):
        """MODEL Lee_Kd_LS8 (mode 0) / Lee_Secchi_LS8 (mode 1) on four host planes (secchi.c:13-252)."""
        pl = [np.ascontiguousarray(a, dtype=np.float32) for a in (coastal, blue, green, red)]
        out = np.zeros_like(pl[0])
        sp = np.ascontiguousarray(spv, dtype=np.float32)
        check(self.lib.phb_lee_ls8_host(self.ctx, mode, *[_np_ptr(a, capi._fp) for a in pl], _np_ptr(sp, capi._fp),
                                        float(theta_s), out.shape[0], out.shape[1], _np_ptr(out, capi._fp)))
        return out

    def lee_ls8_device(self, mode: int, coastal, blue, green, red, spv, theta_s: float, out=None, stream=None):
        """Same on torch CUDA tensors (float32, contiguous); returns the output tensor."""
        import torch
        out = torch.empty_like(coastal) if out is None else out
        sp = np.ascontiguousarray(spv, dtype=np.float32)
        s = torch.cuda.current_stream(coastal.device) if stream is None else stream
        check(self.lib.phb_lee_ls8_device(self.ctx, mode, C.c_void_p(coastal.data_ptr()), C.c_void_p(blue.data_ptr()),
                                          C.c_void_p(green.data_ptr()), C.c_void_p(red.data_ptr()), _np_ptr(sp, capi._fp),
                                          float(theta_s), coastal.numel(), C.c_void_p(out.data_ptr()),
                                          C.c_void_p(s.cuda_stream)))
        return out

    def fp64_peak(self):
        t, ms = C.c_double(0), C.c_float(0)
        check(self.lib.phb_fp64_peak(self.ctx, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    # ---- REFINE ----------------------------------------------------------------------------------------
    def refine_host(self, grid, nodata, flags, args, land=None, land_nodata=-9999.0, shallow=None,
                    shallow_nodata=-9999.0):
        g = np.ascontiguousarray(grid, dtype=np.float32)
        out = np.zeros_like(g)
        a = np.ascontiguousarray(args, dtype=np.float32)
        ld = None if land is None else np.ascontiguousarray(land, dtype=np.float32)
        sh = None if shallow is None else np.ascontiguousarray(shallow, dtype=np.float32)
        null = C.cast(None, capi._fp)
        check(self.lib.phb_refine_host(self.ctx, _np_ptr(g, capi._fp), nodata, null if ld is None else _np_ptr(ld, capi._fp),
                                       land_nodata, null if sh is None else _np_ptr(sh, capi._fp), shallow_nodata,
                                       g.shape[0], g.shape[1], flags, _np_ptr(a, capi._fp), _np_ptr(out, capi._fp)))
        return out


class _DevView:
    """A raw device pointer as something torch can wrap (__cuda_array_interface__, version 2)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Band:
    """A row band of a scene resident on one device and shareable with the other devices of the box (phb_shard_*,
    include/photic_b200.h): the device-resident form of the multi-GPU path. ``desc.nrows`` counts the band's raster
    INCLUDING its halo rows; rows [row_begin, row_end) of that raster are the band's own. The rasters live in the
    library's allocation (one CUDA IPC handle per band); ``planes`` / ``prior`` / ``outputs`` are torch views of it."""

    def __init__(self, inverter: Inverter, desc: SceneDesc, row_begin: int, row_end: int, scene_planes: bool = False):
        import torch

        self.inv, self.desc, self.row_begin, self.row_end = inverter, desc, row_begin, row_end
        self.lib = inverter.lib
        self.h = C.c_void_p()
        check(self.lib.phb_shard_create(inverter.ctx, C.byref(desc), row_begin, row_end, 1 if scene_planes else 0,
                                        C.byref(self.h)))
        dp, dpr, o = C.c_void_p(), C.c_void_p(), Outputs()
        check(self.lib.phb_shard_buffers(self.h, C.byref(dp), C.byref(dpr), C.byref(o)))
        dev = torch.device("cuda", inverter.device)
        nr, nc, ns = desc.nrows, desc.ncols, desc.n_scenes
        sb = sum(desc.n_bands[s] for s in range(ns))
        mb = max(desc.n_bands[s] for s in range(ns))

        def view(ptr, shape, typestr="<f4"):
            return torch.as_tensor(_DevView(ptr, shape, typestr), device=dev)

        self.planes = view(dp.value, (sb, nr, nc))
        self.prior = view(dpr.value, (nr, nc)) if dpr.value else None
        self.outputs = {n: view(C.cast(getattr(o, n), C.c_void_p).value, (nr, nc)) for n in SCALAR_PLANES}
        if scene_planes:
            self.outputs["K"] = view(C.cast(o.K, C.c_void_p).value, (ns, mb, nr, nc))
            for n in ("P", "G", "X"):
                self.outputs[n] = view(C.cast(getattr(o, n), C.c_void_p).value, (ns, nr, nc))
        self.outputs["converged"] = view(C.cast(o.converged, C.c_void_p).value, (nr, nc), "|u1")
        self.outputs["n_evals"] = view(C.cast(o.n_evals, C.c_void_p).value, (nr, nc), "<i4")

    def export(self) -> bytes:
        """The band's handle (512 bytes, POD): what the other devices need to take work from this band."""
        h = capi.ShardHandle()
        check(self.lib.phb_shard_export(self.h, C.byref(h)))
        return bytes(h.bytes)

    @staticmethod
    def _stream(stream, device):
        import torch
        s = torch.cuda.current_stream(device) if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    def prepare(self, stream=None):
        """Validity scan, output defaults and work queues of this band (asynchronous on the stream)."""
        check(self.lib.phb_shard_prepare(self.h, self._stream(stream, self.planes.device)))

    def solve(self, peers=(), stream=None):
        """Inverts: this band's pixels first, then pixels taken from the peers' queues (handles from export(), in the
        order to try them). Every band must be prepared before any solve starts and results are complete only once
        every device's solve has finished -- the caller synchronises. Returns this DEVICE's stats."""
        n = len(peers)
        arr = (capi.ShardHandle * max(n, 1))()
        for k, b in enumerate(peers):
            C.memmove(C.byref(arr[k]), bytes(b), 512)
        st = Stats()
        check(self.lib.phb_shard_solve(self.h, arr, n, self._stream(stream, self.planes.device), C.byref(st)))
        return st.as_dict()

    def valid(self) -> int:
        return int(self.lib.phb_shard_valid(self.h))

    def close(self):
        if self.h:
            self.planes = self.prior = None
            self.outputs = {}
            self.lib.phb_shard_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------
# The reference's call surface, in Python (model/samodel.h:8-19, model/common.h:69-84,194-218)
# --------------------------------------------------------------------------------------------------

@dataclass
class geogrid:
    """model/common.h:69-84 (fields samodel reads)."""
    array: np.ndarray
    nodata_value: float = -9999.0
    nrows: int = 0
    ncols: int = 0

    def __post_init__(self):
        self.nrows, self.ncols = self.array.shape


@dataclass
class scene:
    """model/common.h:194-218 (fields samodel reads)."""
    scene_name: str
    band_indexes: list
    wavelengths: list  # int nm
    theta_v: float
    theta_w: float
    H_tide: float = 0.0
    R_sigma: list = field(default_factory=list)

    @property
    def n_bands(self):
        return len(self.band_indexes)


_default_inverter = None


def samodel(scene_data, gridded_data, scene_indexes, nscenes, empirical_depth_present, empirical_depths,
            n_smoothing_radius, n_spatial, n_bottoms, depth, depth_sigma, model_error, bottom_albedo, bottom_sand,
            bottom_seagrass, bottom_coral, K_min, bottom_type, index_optical_depth, pagesize=8.0, background=0,
            linewidth=1, inverter: Inverter | None = None, sigma_seed: int | None = None, sigma_chain: int = 1):
    """Same arguments as the reference's samodel(); the ten output grids are filled in place.

    Differences from the reference, all documented in DESIGN.md: every valid pixel is inverted from
    a cold start (no LUT / hot start, which make the reference order dependent); the depth-error
    pass (depth_sigma, samodel.c:1376-1477) takes its seed as an argument (time() when None, like the
    reference) and needs the DEPTHS grid (0 without one).
    Returns the stats dict (the reference returns nothing).
    """
    global _default_inverter
    inv = inverter
    if inv is None:
        if _default_inverter is None:
            _default_inverter = Inverter(0)
        inv = _default_inverter
    scs = [scene_data[scene_indexes[k]] for k in range(nscenes)]
    g0 = gridded_data[scs[0].band_indexes[0]]
    for s in scs:  # every grid of the call must have the shape of the first one (the reference indexes them alike)
        for k in s.band_indexes:
            if (gridded_data[k].nrows, gridded_data[k].ncols) != (g0.nrows, g0.ncols):
                raise ValueError(f"grid {k} is {gridded_data[k].nrows}x{gridded_data[k].ncols}, expected {g0.nrows}x{g0.ncols}")
    nd = [[float(gridded_data[k].nodata_value) for k in s.band_indexes] for s in scs]
    same_nodata = all(v == nd[0][0] for row in nd for v in row)
    desc = capi.make_desc([list(s.wavelengths)[:s.n_bands] for s in scs], [s.theta_v for s in scs],
                          [s.theta_w for s in scs], [s.H_tide for s in scs], g0.nrows, g0.ncols,
                          nodata=g0.nodata_value, prior_present=bool(empirical_depth_present),
                          prior_nodata=empirical_depths.nodata_value if empirical_depth_present else -9999.0,
                          n_smooth=n_smoothing_radius, n_spatial=n_spatial, n_bottoms=n_bottoms,
                          r_sigma=[[(s.R_sigma[b] if b < len(s.R_sigma) else 0.0) for b in range(s.n_bands)] for s in scs],
                          nodata_band=None if same_nodata else nd)
    planes = [gridded_data[k].array for s in scs for k in s.band_indexes]
    buffers = {"depth": depth, "model_error": model_error, "bottom_albedo": bottom_albedo, "bottom_sand": bottom_sand,
               "bottom_seagrass": bottom_seagrass, "bottom_coral": bottom_coral, "K_min": K_min,
               "bottom_type": bottom_type, "index_optical_depth": index_optical_depth}
    for v in buffers.values():
        assert v.dtype == np.float32 and v.flags["C_CONTIGUOUS"]
    out, stats = inv.invert_host(desc, planes, empirical_depths.array if empirical_depth_present else None,
                                 buffers=buffers)
    depth_sigma[...] = 0.0
    if empirical_depth_present:
        import time
        seed = int(time.time()) if sigma_seed is None else int(sigma_seed)
        sig, table, _, st2 = inv.depth_sigma_host(desc, planes, empirical_depths.array, depth, seed & 0xFFFFFFFF, 128,
                                                  sigma_chain)
        depth_sigma[...] = sig
        stats["sigma_intervals"], stats["sigma_trials"], stats["sigma_ms"] = len(table), st2["n_valid"], st2["ms_solve"]
    samodel.last_outputs = out
    return stats
