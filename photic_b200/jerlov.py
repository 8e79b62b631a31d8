"""Host mirror of the reference's COMPUTE K path (model/jerlov.c) over the C ABI.

``jerlov`` / ``compute_k`` / ``compute_k_from_jerlov`` / ``compute_k_from_ratio`` keep the reference's names and
argument meaning (jerlov.h:9-18); the arithmetic is ``phb_jerlov_*`` in libphotic_b200.so (csrc/jerlov_host.h),
bit-identical to the reference. Scene-level scalars: host code by design, no device needed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

WATER_TYPES = ("OI", "OIA", "OIB", "OII", "OIII", "C1", "C3", "C5", "C7", "C9")


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def jerlov(wlen_i, wlen_j, Lsmi, Lsmj, Li, Lj, manual_ratio=0.0):
    """``jerlov`` (jerlov.c:75-210). Returns (success, dict(ki, kj, m, c, r, water_type, n_shallow))."""
    Li = np.ascontiguousarray(Li, dtype=np.float32)
    Lj = np.ascontiguousarray(Lj, dtype=np.float32)
    assert Li.shape == Lj.shape and Li.ndim == 1
    out = np.zeros(6, dtype=np.float32)
    ns = C.c_int32(0)
    rc = capi.lib().phb_jerlov_fit(wlen_i, wlen_j, Lsmi, Lsmj, _fp(Li), _fp(Lj), len(Li), manual_ratio, _fp(out),
                                   C.byref(ns))
    if rc not in (capi.PHB_OK, capi.PHB_ENOFIT):
        capi.check(rc)
    keys = ("ki", "kj", "m", "c", "r", "water_type")
    res = {k: out[i] for i, k in enumerate(keys)}
    res["n_shallow"] = ns.value
    return rc == capi.PHB_OK, res


def compute_k(water_type, wlen):
    """``compute_k`` (jerlov.c:284-316) for one wavelength or an array of them."""
    wl = np.atleast_1d(np.asarray(wlen, dtype=np.float32))
    k = np.zeros_like(wl)
    capi.check(capi.lib().phb_jerlov_k(water_type, _fp(wl), len(wl), _fp(k)))
    return k if np.ndim(wlen) else k[0]


def compute_k_from_jerlov(water_type, alphas, spectral_indexes, wavelengths):
    """``compute_k_from_jerlov`` (jerlov.c:274-282): alphas[spectral_indexes[k]] = compute_k(water_type, wavelengths[k])."""
    alphas[np.asarray(spectral_indexes)] = compute_k(water_type, np.asarray(wavelengths, dtype=np.float32))
    return alphas


def compute_k_from_ratio(ratio, wlen_i, wlen_j, wavelengths):
    """``compute_k_from_ratio`` (jerlov.c:214-270). Returns (success, water_type, k[])."""
    wl = np.ascontiguousarray(wavelengths, dtype=np.float32)
    k = np.zeros_like(wl)
    wt = C.c_float(0.0)
    rc = capi.lib().phb_jerlov_k_from_ratio(ratio, wlen_i, wlen_j, _fp(wl), len(wl), C.byref(wt), _fp(k))
    if rc not in (capi.PHB_OK, capi.PHB_ENOFIT):
        capi.check(rc)
    return rc == capi.PHB_OK, np.float32(wt.value), k


def jerlov_water_type_str(wtype: float) -> str:
    """``jerlov_water_type_str`` (jerlov.c:429-453)."""
    i = int(np.floor(wtype))
    if not 0 <= i < len(WATER_TYPES):
        return ""
    return f"{WATER_TYPES[i]} + {wtype - i:.2f}"
