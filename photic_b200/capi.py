"""ctypes binding of the C ABI in include/photic_b200.h (libphotic_b200.so).

This is what a host application binds; nothing here computes. If the shared library is missing the
import of :func:`lib` raises -- there is no Python or CPU fallback for the inversion.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHB_LIB") or os.path.join(HERE, "csrc", "libphotic_b200.so")

MAX_SCENES, MAX_BANDS, MAX_BOTTOMS = 16, 8, 8

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_scenes", C.c_int32),
        ("n_bands", C.c_int32 * MAX_SCENES),
        ("wavelengths", (C.c_int32 * MAX_BANDS) * MAX_SCENES),
        ("theta_view", C.c_double * MAX_SCENES),
        ("theta_sun", C.c_double * MAX_SCENES),
        ("h_tide", C.c_double * MAX_SCENES),
        ("n_smoothing_radius", C.c_int32),
        ("n_spatial", C.c_int32),
        ("n_bottoms", C.c_int32),
        ("nrows", C.c_int32),
        ("ncols", C.c_int32),
        ("nodata", C.c_float),
        ("prior_present", C.c_int32),
        ("prior_nodata", C.c_float),
        ("r_sigma", (C.c_double * MAX_BANDS) * MAX_SCENES),
        ("nodata_per_band", C.c_int32),
        ("nodata_band", (C.c_float * MAX_BANDS) * MAX_SCENES),
    ]


class Outputs(C.Structure):
    _fields_ = [(n, _fp) for n in ("depth", "model_error", "bottom_albedo", "bottom_sand", "bottom_seagrass",
                                   "bottom_coral", "K_min", "bottom_type", "index_optical_depth", "K", "P", "G", "X")]
    _fields_ += [("converged", _u8p), ("n_evals", _ip)]


class Stats(C.Structure):
    _fields_ = [
        ("n_valid", C.c_int64), ("n_shallow", C.c_int64), ("n_evals", C.c_int64), ("n_iters", C.c_int64),
        ("n_converged", C.c_int64), ("alg_flops", C.c_double), ("ms_classify", C.c_float), ("ms_solve", C.c_float),
        ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("warps_per_cta", C.c_int32), ("ctas", C.c_int32),
        ("smem_bytes", C.c_int32), ("regs", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ShardHandle(C.Structure):
    """phb_shard_handle: where a row band lives (CUDA IPC handle / owner pointer) and its view as offsets. POD: ranks
    exchange it as bytes (torch.distributed all_gather of a uint8 tensor)."""
    _fields_ = [("bytes", C.c_ubyte * 512)]


_fpp = C.POINTER(_fp)


class RowOutputs(C.Structure):
    """phb_row_outputs: result grids as row pointers (the reference's float **)."""
    _fields_ = [(n, _fpp) for n in ("depth", "model_error", "bottom_albedo", "bottom_sand", "bottom_seagrass",
                                    "bottom_coral", "K_min", "bottom_type", "index_optical_depth")]
    _fields_ += [(n, C.POINTER(_fpp)) for n in ("K", "P", "G", "X")]


SCALAR_PLANES = ("depth", "model_error", "bottom_albedo", "bottom_sand", "bottom_seagrass", "bottom_coral", "K_min",
                 "bottom_type", "index_optical_depth")

_lib = None


class PhoticError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libphotic_b200.so (built by photic_b200.build). Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PhoticError(f"{LIB_PATH} not built: run `python -m photic_b200.build` (needs nvcc). "
                              "photic_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.phb_error_string.restype = C.c_char_p
        L.phb_error_string.argtypes = [C.c_int]
        L.phb_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.phb_ctx_destroy.argtypes = [C.c_void_p]
        L.phb_ctx_destroy.restype = None
        L.phb_band_tables.argtypes = [C.POINTER(SceneDesc), _dp, _dp]
        L.phb_invert_device.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.POINTER(Outputs), C.c_void_p, C.POINTER(Stats)]
        L.phb_invert_host.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_int,
                                      C.c_int, C.POINTER(Outputs), C.POINTER(Stats)]
        L.phb_debug_record_len.argtypes = [C.POINTER(SceneDesc)]
        L.phb_invert_host_debug.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p), C.c_void_p,
                                            C.c_int, C.c_int, C.POINTER(Outputs), _dp, _ip, _ip, C.c_int64,
                                            C.POINTER(Stats)]
        L.phb_kat_objective.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.c_int, C.c_int, C.c_int, _dp, C.c_int,
                                        C.c_int, _dp, _dp]
        L.phb_kat_math.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int64, _dp]
        L.phb_eval_bench.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp,
                                     C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _fp]
        L.phb_refine_minmax_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, _fp, C.c_void_p]
        L.phb_refine_device.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p,
                                        C.c_float, C.c_int64, C.c_int, _fp, _fp, C.c_void_p, C.c_void_p]
        L.phb_refine_host.argtypes = [C.c_void_p, _fp, C.c_float, _fp, C.c_float, _fp, C.c_float, C.c_int, C.c_int,
                                      C.c_int, _fp, _fp]
        L.phb_fp64_peak.argtypes = [C.c_void_p, _dp, _fp]
        L.phb_lee_ls8_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _fp, C.c_float,
                                         C.c_int64, C.c_void_p, C.c_void_p]
        L.phb_lee_ls8_host.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_float, C.c_int, C.c_int, _fp]
        L.phb_depth_sigma_host.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                           C.c_uint, C.c_int, C.c_int, C.c_int, _fp, _dp, C.POINTER(C.c_int32), _dp,
                                           C.POINTER(Stats)]
        L.phb_depth_sigma_rows.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, _dp, C.POINTER(C.c_int32), _dp,
                                           C.POINTER(Stats)]
        L.phb_nc_pack_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, _fp, _fp,
                                         C.POINTER(C.c_int16), C.c_void_p]
        L.phb_nc_unpack_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int16, C.c_double,
                                           C.c_void_p, C.c_void_p]
        L.phb_nc_pack_host.argtypes = [C.c_void_p, _fp, C.c_int, C.c_int, C.c_double, C.c_void_p, _fp, _fp, C.POINTER(C.c_int16)]
        L.phb_nc_unpack_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int16, C.c_double, _fp]
        L.phb_plan_row_bands.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_int,
                                         C.POINTER(C.c_int32), _dp]
        L.phb_invert_host_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(SceneDesc), C.POINTER(C.c_void_p),
                                            C.c_void_p, C.POINTER(Outputs), C.POINTER(Stats), C.POINTER(Stats),
                                            C.POINTER(C.c_int32)]
        L.phb_invert_rows.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(SceneDesc), C.c_void_p, C.c_void_p,
                                      C.POINTER(RowOutputs), C.POINTER(Stats), C.POINTER(Stats), C.POINTER(C.c_int32)]
        L.phb_shard_create.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.phb_shard_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(Outputs)]
        L.phb_shard_export.argtypes = [C.c_void_p, C.POINTER(ShardHandle)]
        L.phb_shard_prepare.argtypes = [C.c_void_p, C.c_void_p]
        L.phb_shard_solve.argtypes = [C.c_void_p, C.POINTER(ShardHandle), C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.phb_shard_valid.argtypes = [C.c_void_p]
        L.phb_shard_valid.restype = C.c_int64
        L.phb_shard_destroy.argtypes = [C.c_void_p]
        L.phb_shard_destroy.restype = None
        L.phb_debug_model_const.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_int64]
        L.phb_debug_model_const.restype = C.c_int64
        L.phb_jerlov_fit.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, C.c_int, C.c_float, _fp,
                                     C.POINTER(C.c_int32)]
        L.phb_jerlov_k.argtypes = [C.c_float, _fp, C.c_int, _fp]
        L.phb_jerlov_k_from_ratio.argtypes = [C.c_float, C.c_float, C.c_float, _fp, C.c_int, _fp, _fp]
        _lib = L
    return _lib


PHB_OK, PHB_EINVAL, PHB_ENODEVICE, PHB_ECUDA, PHB_ENOMEM, PHB_ENOFIT, PHB_ENOPEER = range(7)  # include/photic_b200.h

EXPORTS = ["phb_version", "phb_error_string", "phb_device_count", "phb_ctx_create", "phb_ctx_destroy",
           "phb_band_tables", "phb_invert_device", "phb_invert_host", "phb_debug_record_len", "phb_invert_host_debug",
           "phb_kat_objective", "phb_kat_math", "phb_eval_bench", "phb_refine_minmax_device", "phb_refine_device", "phb_refine_host",
           "phb_fp64_peak", "phb_depth_sigma_host", "phb_lee_ls8_device", "phb_lee_ls8_host",
           "phb_jerlov_fit", "phb_jerlov_k", "phb_jerlov_k_from_ratio",
           "phb_plan_row_bands", "phb_invert_host_multi", "phb_debug_model_const", "phb_invert_rows",
           "phb_shard_create", "phb_shard_buffers", "phb_shard_export", "phb_shard_prepare", "phb_shard_solve",
           "phb_shard_valid", "phb_shard_destroy", "phb_depth_sigma_rows",
           "phb_nc_pack_device", "phb_nc_unpack_device", "phb_nc_pack_host", "phb_nc_unpack_host"]


def check(rc: int) -> None:
    if rc != 0:
        e = PhoticError(f"photic_b200 error {rc}: {lib().phb_error_string(rc).decode()}")
        e.code = rc
        raise e


def make_desc(wavelengths, theta_view, theta_sun, h_tide, nrows, ncols, nodata=-9999.0, prior_present=True,
              prior_nodata=-9999.0, n_smooth=1, n_spatial=2, n_bottoms=3, r_sigma=1.0e-4, nodata_band=None) -> SceneDesc:
    """wavelengths: one band list for every scene, or one list PER scene -- the lists may differ in length (mixed
    sensors: the reference and the C ABI allow a different n_bands per scene, samodel.c:403-409). r_sigma: scalar or
    per scene per band (ragged allowed). nodata_band: per scene per band nodata values (every grid is tested against
    its own, samodel.c:683) or None when all grids share `nodata`."""
    ns = len(theta_sun)
    per_scene = len(wavelengths) > 0 and not np.isscalar(wavelengths[0])
    wls = [list(wavelengths[s]) for s in range(ns)] if per_scene else [list(wavelengths)] * ns
    d = SceneDesc()
    d.n_scenes = ns
    tv = np.broadcast_to(np.asarray(theta_view, dtype=np.float64), (ns,))
    for s in range(ns):
        d.n_bands[s] = len(wls[s])
        for b in range(len(wls[s])):
            d.wavelengths[s][b] = int(wls[s][b])
            if np.isscalar(r_sigma):
                d.r_sigma[s][b] = float(r_sigma)
            else:
                row = r_sigma[s]
                d.r_sigma[s][b] = float(row[b]) if b < len(row) else 0.0
            if nodata_band is not None:
                d.nodata_band[s][b] = float(nodata_band[s][b])
        d.theta_view[s] = float(tv[s])
        d.theta_sun[s] = float(theta_sun[s])
        d.h_tide[s] = float(h_tide[s])
    d.n_smoothing_radius, d.n_spatial, d.n_bottoms = n_smooth, n_spatial, n_bottoms
    d.nrows, d.ncols = nrows, ncols
    d.nodata = nodata
    d.nodata_per_band = 0 if nodata_band is None else 1
    d.prior_present = 1 if prior_present else 0
    d.prior_nodata = prior_nodata
    return d


def desc_from_spec(spec, nrows=None, prior_present=True) -> SceneDesc:
    ns = spec.n_dates
    return make_desc(spec.wavelengths, spec.theta_view, [spec.theta_sun(s) for s in range(ns)],
                     [spec.h_tide(s) for s in range(ns)], spec.nrows if nrows is None else nrows, spec.ncols,
                     prior_present=prior_present, n_smooth=spec.n_smoothing_radius, n_spatial=spec.n_spatial,
                     n_bottoms=spec.n_bottoms, r_sigma=spec.r_sigma)
