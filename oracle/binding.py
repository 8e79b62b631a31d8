"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle libraries.

``Oracle("port")``      -> oracle/_build/libphotic_oracle.so  (plain-C restatement, photic_oracle.c)
``Oracle("reference")`` -> oracle/_ref/libphotic_ref.so       (UNMODIFIED reference hot path +
                                                                ref_harness.c; built only where
                                                                /root/reference exists)
Both expose the same entry points (prefix ``pho_`` / ``ref_``) with the same signatures.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libphotic_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libphotic_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


def build(ref: bool = True) -> None:
    """Compile the restatement (always) and the reference (only where its sources exist)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/model"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


class SceneCfg:
    """Plain description of the acquisition geometry (what SCENE ... sets up, bam.c:1310-1543)."""

    def __init__(self, wavelengths, theta_view, theta_sun, h_tide, r_sigma=None, n_smooth=1, n_spatial=2,
                 n_bottoms=3):
        wl = np.asarray(wavelengths, dtype=np.int32)
        if wl.ndim == 1:
            wl = np.tile(wl, (len(theta_sun), 1))
        self.ns, self.maxb = wl.shape
        self.wavelengths = _i(wl)
        self.n_bands = _i(np.full(self.ns, self.maxb))
        self.theta_v = _d(np.broadcast_to(np.asarray(theta_view, dtype=np.float64), (self.ns,)))
        self.theta_w = _d(theta_sun)
        self.h_tide = _d(h_tide)
        self.r_sigma = _d(np.full((self.ns, self.maxb), 1.0e-4) if r_sigma is None else r_sigma)
        self.n_smooth, self.n_spatial, self.n_bottoms = n_smooth, n_spatial, n_bottoms

    @classmethod
    def from_spec(cls, spec):
        ns = spec.n_dates
        return cls(spec.wavelengths, spec.theta_view, [spec.theta_sun(s) for s in range(ns)],
                   [spec.h_tide(s) for s in range(ns)], np.full((ns, spec.n_bands), spec.r_sigma),
                   spec.n_smoothing_radius, spec.n_spatial, spec.n_bottoms)

    def head(self):
        return (C.c_int(self.ns), C.c_int(self.maxb), _p(self.n_bands, _ip), _p(self.wavelengths, _ip),
                _p(self.theta_v, _dp), _p(self.theta_w, _dp), _p(self.h_tide, _dp), _p(self.r_sigma, _dp))


REC_FIELDS = ["depth", "Rrs_error", "bottom_albedo", "sand", "seagrass", "coral", "K_min", "iod", "bottom_type",
              "model_error", "depth_error", "bottom_error", "K_error", "n_regions", "origin", "h_prior"]


class Oracle:
    def __init__(self, kind: str = "port"):
        self.kind = kind
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run oracle.binding.build()")
        self.lib = C.CDLL(path)
        self.pre = "pho_" if kind == "port" else "ref_"
        self._fn("interp_1d").restype = C.c_double

    def _fn(self, name):
        return getattr(self.lib, self.pre + name)

    def record_len(self, cfg):
        return int(self._fn("record_len")(cfg.ns, cfg.maxb))

    def tables(self, cfg):
        out = np.zeros((cfg.ns, cfg.maxb, 4 + cfg.n_bottoms))
        out2 = np.zeros(1 + 2 * cfg.ns)
        self._fn("tables")(*cfg.head(), C.c_int(cfg.n_bottoms), _p(out, _dp), _p(out2, _dp))
        return out, out2

    def invert_pixels(self, cfg, planes, nodata, prior, prior_nodata, pix_i, pix_j, nthreads=0, variant=None):
        """Per-pixel cold-start inversion. Returns dict(rec, status, converged, n_evals, n_restarts[, n_iters]);
        n_restarts = nelmin's numres (asa047.c:493) summed over the H starts of the pixel."""
        planes = _f(planes)
        _, nrows, ncols = planes.shape
        pix_i, pix_j = _i(pix_i), _i(pix_j)
        npix = len(pix_i)
        rl = self.record_len(cfg)
        rec = np.zeros((npix, rl))
        status, conv, nev, nit = (np.zeros(npix, dtype=np.int32) for _ in range(4))
        pr = None if prior is None else _f(prior)
        prp = C.cast(None, _fp) if pr is None else _p(pr, _fp)
        tail = (C.c_int(cfg.n_smooth), C.c_int(cfg.n_spatial), C.c_int(cfg.n_bottoms), C.c_int(nrows), C.c_int(ncols),
                _p(planes, _fp), C.c_float(nodata), prp, C.c_float(prior_nodata), C.c_int(npix), _p(pix_i, _ip),
                _p(pix_j, _ip), _p(rec, _dp), _p(status, _ip), _p(conv, _ip), _p(nev, _ip))
        nres = np.zeros(npix, dtype=np.int32)
        has_numres = hasattr(self.lib, self.pre + "set_numres_out")  # a prebuilt oracle/_ref of round 1 has none
        if has_numres:
            self._fn("set_numres_out")(_p(nres, _ip))
        if variant is None:
            rc = self._fn("invert_pixels")(*cfg.head(), *tail, C.c_int(nthreads))
        else:
            assert self.kind == "port"
            rc = self.lib.pho_invert_pixels_variant(C.c_int(variant), *cfg.head(), *tail, _p(nit, _ip), C.c_int(nthreads))
        if has_numres:
            self._fn("set_numres_out")(C.cast(None, _ip))
        assert rc == 0, rc
        return {"rec": rec, "status": status, "converged": conv, "n_evals": nev, "n_iters": nit,
                "n_restarts": nres if has_numres else None}

    def error_kat(self, cfg, nb_active, n_regions, origin, meas, params):
        meas, params = _d(meas), _d(params)
        nvec, nparams = params.shape
        out = np.zeros((nvec, 6))
        rrs = np.zeros((nvec, n_regions, cfg.ns, cfg.maxb))
        K = np.zeros((nvec, cfg.ns, cfg.maxb))
        self._fn("error_kat")(*cfg.head(), C.c_int(nb_active), C.c_int(n_regions), C.c_int(origin), _p(meas, _dp),
                              C.c_int(nparams), C.c_int(nvec), _p(params, _dp), _p(out, _dp), _p(rrs, _dp), _p(K, _dp))
        return out, rrs, K

    def interp_1d(self, X, Y, x):
        X, Y = _d(X), _d(Y)
        return float(self._fn("interp_1d")(_p(X, _dp), _p(Y, _dp), C.c_int(len(X)), C.c_double(x)))

    def approx_equal(self, a, b, eps):
        return int(self._fn("approx_equal")(C.c_float(a), C.c_float(b), C.c_float(eps)))

    def nelmin_kat(self, fn_id, start, step, reqmin=1e-2, konvge=100, kcount=5000):
        start, step = _d(start), _d(step)
        n = len(start)
        xmin = np.zeros(n)
        y = C.c_double(0)
        ic, nr, ifl = C.c_int(0), C.c_int(0), C.c_int(0)
        self._fn("nelmin_kat")(C.c_int(fn_id), C.c_int(n), _p(start, _dp), _p(step, _dp), C.c_double(reqmin),
                               C.c_int(konvge), C.c_int(kcount), _p(xmin, _dp), C.byref(y), C.byref(ic), C.byref(nr),
                               C.byref(ifl))
        return xmin, y.value, ic.value, nr.value, ifl.value

    def samodel_as_is(self, cfg, planes, nodata, prior, prior_nodata):
        """The reference's samodel() exactly as shipped (LUT + hot start); reference library only."""
        assert self.kind == "reference"
        planes = _f(planes)
        _, nrows, ncols = planes.shape
        out = np.zeros((10, nrows, ncols), dtype=np.float32)
        pr = None if prior is None else _f(prior)
        prp = C.cast(None, _fp) if pr is None else _p(pr, _fp)
        self.lib.ref_samodel_as_is(*cfg.head(), C.c_int(cfg.n_smooth), C.c_int(cfg.n_spatial), C.c_int(cfg.n_bottoms),
                                   C.c_int(nrows), C.c_int(ncols), _p(planes, _fp), C.c_float(nodata), prp,
                                   C.c_float(prior_nodata), _p(out, _fp))
        return out

    def depth_sigma(self, cfg, planes, nodata, prior, prior_nodata, depth_pos, seed, n_samples=128, chain_mode=0,
                    max_intervals=120):
        """Depth-error phase of samodel() (samodel.c:1376-1477) with the seed as an argument.
        depth_pos: positive depths (0 where nothing was inverted). Returns (table, trials, depth_sigma)."""
        planes, pr, dp = _f(planes), _f(prior), _f(depth_pos)
        _, nrows, ncols = planes.shape
        table = np.zeros(max_intervals)
        trials = np.zeros((max_intervals, n_samples))
        sig = np.zeros((nrows, ncols), dtype=np.float32)
        nint = C.c_int(0)
        rc = self._fn("depth_sigma")(*cfg.head(), C.c_int(cfg.n_smooth), C.c_int(cfg.n_spatial), C.c_int(cfg.n_bottoms),
                                     C.c_int(nrows), C.c_int(ncols), _p(planes, _fp), C.c_float(nodata), _p(pr, _fp),
                                     C.c_float(prior_nodata), _p(dp, _fp), C.c_uint(seed), C.c_int(n_samples),
                                     C.c_int(chain_mode), C.c_int(max_intervals), _p(table, _dp), C.byref(nint),
                                     _p(trials, _dp), _p(sig, _fp))
        assert rc == 0, rc
        return table[:nint.value], trials[:nint.value], sig

    def lee_ls8(self, mode, coastal, blue, green, red, spv, theta_s):
        """MODEL Lee_Kd_LS8 (mode 0) / Lee_Secchi_LS8 (mode 1) on four float32 planes (secchi.c:13-252)."""
        pl = [_f(a) for a in (coastal, blue, green, red)]
        out = np.zeros_like(pl[0])
        sp = _f(spv)
        self._fn("lee_ls8")(C.c_int(mode), C.c_int(out.shape[0]), C.c_int(out.shape[1]), *[_p(a, _fp) for a in pl],
                            _p(sp, _fp), C.c_float(theta_s), _p(out, _fp))
        return out

    def nc_pack(self, grid, spval):
        """compress_2d (nc.c:271-320): int16 packing of a float grid. Returns (packed int16, add_offset, scale_factor, missing)."""
        grid = _f(grid)
        packed = np.zeros(grid.shape, dtype=np.int16)
        os2 = np.zeros(2, dtype=np.float32)
        miss = C.c_short(0)
        self._fn("nc_pack")(C.c_int(grid.shape[0]), C.c_int(grid.shape[1]), _p(grid, _fp), C.c_double(spval),
                             packed.ctypes.data_as(C.c_void_p), _p(os2, _fp), C.byref(miss))
        return packed, os2[0], os2[1], int(miss.value)

    def nc_unpack(self, packed, add_offset, scale_factor, missing, spval):
        """decompress_2d (nc.c:247-266)."""
        packed = np.ascontiguousarray(packed, dtype=np.int16)
        out = np.zeros(packed.shape, dtype=np.float32)
        self._fn("nc_unpack")(C.c_int(packed.shape[0]), C.c_int(packed.shape[1]), packed.ctypes.data_as(C.c_void_p),
                               C.c_float(add_offset), C.c_float(scale_factor), C.c_short(missing), C.c_double(spval), _p(out, _fp))
        return out

    def refine(self, grid, nodata, land, land_nodata, shallow, shallow_nodata, flags, args):
        """REFINE: pho_refine (restatement) or ref_refine (the reference's own run_refine(), refine.c:12-302)."""
        grid = _f(grid)
        out = np.zeros_like(grid)
        args = _f(args)
        ld = None if land is None else _f(land)
        sh = None if shallow is None else _f(shallow)
        self._fn("refine")(C.c_int(grid.shape[0]), C.c_int(grid.shape[1]), _p(grid, _fp), C.c_float(nodata),
                            C.cast(None, _fp) if ld is None else _p(ld, _fp), C.c_float(land_nodata),
                            C.cast(None, _fp) if sh is None else _p(sh, _fp), C.c_float(shallow_nodata),
                            C.c_int(flags), _p(args, _fp), _p(out, _fp))
        return out
