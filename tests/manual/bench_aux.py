#!/usr/bin/env python3
"""Measurement of the rows beside the inversion kernel (8a a9): one JSON line each.

  lee_ls8   MODEL Lee_Kd_LS8 / Lee_Secchi_LS8 kernel  -> GB/s (algorithmic: 16 B in + 4 B out per cell) vs the HBM roof
  refine    REFINE kernel (CLIP|SCALE|POWER)          -> GB/s (4 B in + 4 B out per cell)
  sigma     depth-error estimate (samodel.c:1376-1477) -> trials/s in both chain modes
each with the reference's CPU code timed beside it on a bounded sample (oracle/_ref when present, else the port).
usage: python tests/manual/bench_aux.py [--rows 3930 --cols 2858]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from photic_b200 import capi, scene
from photic_b200.samodel import Inverter
from oracle.binding import REF_SO, Oracle, SceneCfg

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=3930)
ap.add_argument("--cols", type=int, default=2858)
args = ap.parse_args()
inv = Inverter(0)
kind = "reference" if os.path.exists(REF_SO) else "port"
cpu = Oracle(kind)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6462.1))


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- Lee Kd / Secchi ---------------------------------------------------------------------------------
spec = scene.CONFIGS["exmouth"].scaled(args.rows, args.cols)
planes, prior = scene.generate(spec, device="cuda")
c, b, g, r = [planes[k].contiguous() for k in range(4)]
spv = np.full(4, scene.NODATA, dtype=np.float32)
out = torch.empty_like(c)
n = c.numel()
for mode, name in ((0, "Lee_Kd_LS8"), (1, "Lee_Secchi_LS8")):
    ms = timed(lambda: inv.lee_ls8_device(mode, c, b, g, r, spv, 28.0, out=out))
    sub = slice(0, max(1, 400_000 // args.cols))
    hc, hb, hg, hr = [t[sub].cpu().numpy() for t in (c, b, g, r)]
    t0 = time.time(); ref = cpu.lee_ls8(mode, hc, hb, hg, hr, spv, 28.0); t1 = time.time() - t0
    same = np.array_equal(out[sub].cpu().numpy().view(np.int32), ref.view(np.int32))
    gbs = 20.0 * n / (ms * 1e-3) / 1e9
    print(json.dumps({"row": "N3", "kernel": "lee_ls8_kernel", "model": name, "cells": n, "ms": ms, "Mcells_per_s": n / ms / 1e3,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                   "alg_bytes_per_cell": 20},
                      "cpu_baseline": {"kind": kind, "cores": 1, "Mcells_per_s": hc.size / t1 / 1e6, "sample": f"{hc.size} cells"},
                      "bit_identical_on_sample": bool(same)}))

# ---- REFINE ------------------------------------------------------------------------------------------
depth = (-torch.rand((args.rows, args.cols), device="cuda") * 35.0).contiguous()
depth[planes[0] == scene.NODATA] = scene.NODATA
ro = torch.empty_like(depth)
import ctypes as C
rargs = np.array([-30.0, -0.5, -40.0, 0.0, 1.3, 0.9, -0.25, -32.0, -1.0, 1.1, 0.95], dtype=np.float32)
flags = 1 | 2 | 16
def run_refine():
    capi.check(inv.lib.phb_refine_device(inv.ctx, C.c_void_p(depth.data_ptr()), C.c_float(scene.NODATA), None, C.c_float(0), None,
                                         C.c_float(0), depth.numel(), flags, rargs.ctypes.data_as(capi._fp), None,
                                         C.c_void_p(ro.data_ptr()), None))
ms = timed(run_refine)
sub = slice(0, max(1, 1_000_000 // args.cols))
hd = depth[sub].cpu().numpy()
t0 = time.time(); ref = Oracle("port").refine(hd, scene.NODATA, None, 0.0, None, 0.0, flags, rargs); t1 = time.time() - t0
gbs = 8.0 * depth.numel() / (ms * 1e-3) / 1e9
print(json.dumps({"row": "a9", "kernel": "refine_kernel", "flags": "CLIP|SCALE|POWER", "cells": depth.numel(), "ms": ms,
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "alg_bytes_per_cell": 8},
                  "cpu_baseline": {"kind": "port", "cores": 1, "Mcells_per_s": hd.size / t1 / 1e6, "sample": f"{hd.size} cells"},
                  "bit_identical_on_sample": bool(np.array_equal(ro[sub].cpu().numpy().view(np.int32), ref.view(np.int32)))}))

# ---- depth-error estimate ----------------------------------------------------------------------------
sp = scene.CONFIGS["exmouth"].scaled(256, 256)
pl, pr = scene.generate(sp)
pl, pr = pl.numpy(), pr.numpy()
desc = capi.desc_from_spec(sp)
outp, st = inv.invert_host(desc, pl, pr, scene_planes=False)
for mode, name, ns in ((1, "per-interval chains", 128), (0, "reference single chain", 8)):
    t0 = time.time()
    sig, table, trials, st2 = inv.depth_sigma_host(desc, pl, pr, outp["depth"], 4242, ns, mode)
    wall = time.time() - t0
    line = {"row": "N1", "what": "depth-error estimate", "chain_mode": name, "n_samples": ns, "scene": "exmouth-shaped 256x256, 6 dates",
            "intervals": len(table), "trials": st2["n_valid"], "kernel_ms": st2["ms_solve"], "wall_s": wall,
            "trials_per_s": st2["n_valid"] / (st2["ms_solve"] * 1e-3 + 1e-9), "evals_per_trial": st2["n_evals"] / max(1, st2["n_valid"])}
    if mode == 0:
        t0 = time.time()
        tb, tr, sg = cpu.depth_sigma(SceneCfg.from_spec(sp), pl, scene.NODATA, pr, scene.NODATA, -outp["depth"], 4242, ns, mode)
        t1 = time.time() - t0
        line["cpu_baseline"] = {"kind": kind, "cores": 1, "trials_per_s": int((tr != 0).sum()) / t1, "wall_s": t1}
        line["bit_identical_to_cpu"] = bool(np.array_equal(tr.view(np.int64), trials.view(np.int64)) and np.array_equal(sg.view(np.int32), sig.view(np.int32)))
    print(json.dumps(line))
