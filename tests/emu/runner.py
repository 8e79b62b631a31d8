"""TEST INFRASTRUCTURE ONLY: builds and drives tests/emu/emu_solve.cpp -- the product's solve_kernel source compiled
by g++ (-DPHB_HOST_EMU) and run on the CPU, one warp of 32 fibers (see tests/emu/include/cuda_runtime.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(ROOT, "tests", "_build")
SRC = os.path.join(HERE, "emu_solve.cpp")
DEPS = [SRC, os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "math_constants.h")] + [
    os.path.join(ROOT, "photic_b200", "csrc", f) for f in ("invert_kernel.cuh", "exact_math.cuh", "device_model.cuh",
                                                           "libm_tables.h")]


def build(defines=(), tag="") -> str:
    """g++ build of the emulation; `defines` are extra -D flags of the kernel source (its experiment switches)."""
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, f"libemu_solve{('_' + tag) if tag else ''}.so")
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in DEPS):
        cmd = ["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas",
               "-DPHB_HOST_EMU", *[f"-D{d}" for d in defines], "-I", os.path.join(HERE, "include"), "-o", so, SRC]
        subprocess.run(cmd, check=True)
    return so


class Emulator:
    def __init__(self, defines=(), tag=""):
        self.lib = C.CDLL(build(defines, tag))
        self.lib.emu_model_const_size.restype = C.c_int64

    def invert_pixels(self, desc, planes, prior, pix_i, pix_j, simplex_smem_bytes=4096):
        """Same contract as oracle.binding.Oracle.invert_pixels: full-precision records of the listed pixels, in the
        given order (they are processed in that order by the one emulated warp)."""
        from photic_b200 import capi
        L = capi.lib()
        n = L.phb_debug_model_const(C.byref(desc), None, 0)
        assert n == self.lib.emu_model_const_size(), "ModelConst differs between the product library and the emulation"
        model = (C.c_ubyte * n)()
        assert L.phb_debug_model_const(C.byref(desc), model, n) == n
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        pr = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32)
        q = (np.asarray(pix_i, dtype=np.int64) * desc.ncols + np.asarray(pix_j, dtype=np.int64)).astype(np.int32)
        nq, reclen = len(q), self.lib.emu_record_len(model)
        rec = np.zeros((nq, reclen))
        pix = np.full(nq, -1, dtype=np.int32)
        it = np.zeros((nq, 3), dtype=np.int32)
        out9 = np.full((9, desc.nrows, desc.ncols), 7.0, dtype=np.float32)
        conv = np.zeros((desc.nrows, desc.ncols), dtype=np.uint8)
        nev = np.zeros((desc.nrows, desc.ncols), dtype=np.int32)
        cnt = np.zeros(4, dtype=np.uint64)
        fl = C.c_double(0.0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu_invert(model, C.c_int64(n), vp(planes), C.c_void_p(None) if pr is None else vp(pr), vp(q), nq,
                                 int(simplex_smem_bytes), vp(rec), vp(pix), vp(it), vp(out9), vp(conv), vp(nev), vp(cnt),
                                 C.byref(fl))
        assert rc == 0, rc
        # records come back in queue order: the pixels of class 0 (all substrates) first, then the sand-only ones
        a, b = np.argsort(pix, kind="stable"), np.argsort(q, kind="stable")
        assert np.array_equal(pix[a], q[b])
        back = np.empty(nq, dtype=np.int64)
        back[b] = a
        rec, it = rec[back], it[back]
        return {"rec": rec, "n_evals": it[:, 0], "converged": it[:, 1] & 1, "n_iters": it[:, 1] >> 1, "n_restarts": it[:, 2],
                "planes": out9,
                "converged_plane": conv, "n_evals_plane": nev, "counters": cnt, "alg_flops": fl.value}

    def kat_objective(self, desc, nb_active, n_regions, origin, meas, params):
        """samodel_error on parameter vectors (contract of Inverter.kat_objective / Oracle.error_kat's first output)."""
        from photic_b200 import capi
        L = capi.lib()
        n = L.phb_debug_model_const(C.byref(desc), None, 0)
        model = (C.c_ubyte * n)()
        assert L.phb_debug_model_const(C.byref(desc), model, n) == n == self.lib.emu_model_const_size()
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        params = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros((params.shape[0], 6))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu_kat_objective(model, C.c_int64(n), int(nb_active), int(n_regions), int(origin), vp(meas),
                                        params.shape[0], vp(params), vp(out))
        assert rc == 0, rc
        return out

    def invert_raster(self, desc, planes, prior, row_begin=0, row_end=None, simplex_smem_bytes=4096):
        """The device side of phb_invert_device (classify, queue, solve) on a small raster. Returns the dict that
        Inverter.invert_host(..., debug=True) returns: planes by name, K/P/G/X, converged, n_evals, rec / pix / ..."""
        from photic_b200 import capi
        L = capi.lib()
        n = L.phb_debug_model_const(C.byref(desc), None, 0)
        model = (C.c_ubyte * n)()
        assert L.phb_debug_model_const(C.byref(desc), model, n) == n == self.lib.emu_model_const_size()
        R, Cc, ns = desc.nrows, desc.ncols, desc.n_scenes
        mb = max(desc.n_bands[s] for s in range(ns))
        row_end = R if row_end is None else row_end
        planes = np.ascontiguousarray(planes, dtype=np.float32)
        pr = None if prior is None else np.ascontiguousarray(prior, dtype=np.float32)
        out9 = np.full((9, R, Cc), 7.0, dtype=np.float32)
        K = np.full((ns, mb, R, Cc), 7.0, dtype=np.float32)
        P, G, X = (np.full((ns, R, Cc), 7.0, dtype=np.float32) for _ in range(3))
        conv = np.full((R, Cc), 7, dtype=np.uint8)
        nev = np.full((R, Cc), 7, dtype=np.int32)
        cap = (row_end - row_begin) * Cc
        reclen = self.lib.emu_record_len(model)
        rec = np.zeros((cap, reclen))
        pix = np.full(cap, -1, dtype=np.int32)
        it = np.zeros((cap, 3), dtype=np.int32)
        nv, nsh = C.c_int(0), C.c_int(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu_invert_raster(model, C.c_int64(n), vp(planes), C.c_void_p(None) if pr is None else vp(pr),
                                        int(row_begin), int(row_end), int(simplex_smem_bytes), vp(out9), vp(K), vp(P), vp(G),
                                        vp(X), vp(conv), vp(nev), vp(rec), vp(pix), vp(it), C.byref(nv), C.byref(nsh))
        assert rc == 0, rc
        nvalid = nv.value
        order = np.argsort(pix[:nvalid], kind="stable")
        out = dict(zip(capi.SCALAR_PLANES, out9))
        out.update(K=K, P=P, G=G, X=X, converged=conv, n_evals=nev, rec=rec[:nvalid][order], pix=pix[:nvalid][order],
                   rec_evals=it[:nvalid, 0][order], rec_converged=(it[:nvalid, 1] & 1)[order], rec_restarts=it[:nvalid, 2][order],
                   n_valid=nvalid,
                   n_shallow=nsh.value, queue_order=pix[:nvalid].copy())
        return out
