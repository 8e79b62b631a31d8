#!/usr/bin/env python3
"""End to end through the entry the samodel() shim calls (phb_invert_rows): a whole BASELINE.json scene held on the
host the way the reference holds it (float ** row pointers), over every visible device of the box from ONE process;
wall clock of the whole call -- rows through the pinned ring to the devices, validity scan, inversion with work sharing,
the nine grids (and, --scene-planes, the per-scene K / P / G / X grids samodel() writes to files) back into the
caller's rows -- beside the device time of the solve kernels.  usage: rows_e2e.py [--config exmouth] [--scene-planes]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from photic_b200 import capi, scene
from photic_b200.samodel import Inverter

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="exmouth")
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--cols", type=int, default=0)
ap.add_argument("--scene-planes", action="store_true")
ap.add_argument("--devices", type=int, default=0)
args = ap.parse_args()
spec = scene.CONFIGS[args.config]
if args.rows and args.cols:
    spec = spec.scaled(args.rows, args.cols)
planes, prior = scene.generate(spec, device="cuda")
pl, pr = planes.cpu().numpy(), prior.cpu().numpy()
del planes, prior
torch.cuda.empty_cache()
desc = capi.desc_from_spec(spec)
R, Cc, ns, mb = spec.nrows, spec.ncols, spec.n_dates, 4
fp = C.POINTER(C.c_float)
keep = []


def rows_of(a2d):
    base = a2d.ctypes.data
    arr = (fp * a2d.shape[0])(*[C.cast(base + r * a2d.strides[0], fp) for r in range(a2d.shape[0])])
    keep.append(arr)
    return arr


plane_rows = (C.POINTER(fp) * pl.shape[0])()
for g in range(pl.shape[0]):
    plane_rows[g] = C.cast(rows_of(pl[g]), C.POINTER(fp))
prior_rows = rows_of(pr)
out = capi.RowOutputs()
got = {}
for name in capi.SCALAR_PLANES:
    got[name] = np.zeros((R, Cc), dtype=np.float32)
    setattr(out, name, C.cast(rows_of(got[name]), C.POINTER(fp)))
if args.scene_planes:
    for name, count in (("K", ns * mb), ("P", ns), ("G", ns), ("X", ns)):
        stack = (C.POINTER(fp) * count)()
        got[name] = np.zeros((count, R, Cc), dtype=np.float32)
        for q in range(count):
            stack[q] = C.cast(rows_of(got[name][q]), C.POINTER(fp))
        keep.append(stack)
        setattr(out, name, C.cast(stack, C.POINTER(C.POINTER(fp))))
ndev = args.devices or torch.cuda.device_count()
ivs = [Inverter(k) for k in range(ndev)]
ctxs = (C.c_void_p * ndev)(*[iv.ctx.value for iv in ivs])
per = (capi.Stats * ndev)()
lib = capi.lib()


def call():
    st = capi.Stats()
    t = time.perf_counter()
    capi.check(lib.phb_invert_rows(ctxs, ndev, C.byref(desc), C.cast(plane_rows, C.c_void_p), C.cast(prior_rows, C.c_void_p),
                                   C.byref(out), C.byref(st), per, None))
    return time.perf_counter() - t, st


call()  # warm-up: kernel load on every device, band allocations, peer mappings, the pinned rings
wall, st = call()
ms = [per[k].ms_solve for k in range(ndev)]
n_out = 9 + ((ns * mb + 3 * ns) if args.scene_planes else 0)
print(json.dumps({
    "what": "phb_invert_rows (the call of the samodel() shim), float ** rows in and out, one process",
    "workload": f"{spec.name} {R}x{Cc}, {ns} dates", "devices": ndev, "valid_px": int(st.n_valid), "wall_s": wall,
    "px_per_s_wall": st.n_valid / wall, "solve_kernel_ms_max": max(ms), "px_per_s_kernel": st.n_valid / (max(ms) * 1e-3),
    "wall_over_kernel": wall / (max(ms) * 1e-3), "h2d_bytes": int(pl.nbytes + pr.nbytes), "d2h_bytes": int(n_out * R * Cc * 4),
    "ms_h2d_max": max(per[k].ms_h2d for k in range(ndev)), "ms_d2h_max": max(per[k].ms_d2h for k in range(ndev)),
    "band_balance": min(ms) / max(ms), "scene_planes": bool(args.scene_planes)}))
