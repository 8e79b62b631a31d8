#!/usr/bin/env python3
"""Mines the rare pixels the small parity scenes never contain: `converged == 0` (nelmin's `ifault = 2`,
asa047.c:217, 411-453, 481-493 -> samodel.c:2396-2402) and near-`kcount` evaluation counts, on full-size
BASELINE.json scenes inverted on the GPU.

    python tests/manual/mine_nonconverged.py exmouth 0 3930  qatar 2000 4000  ...

For every (config, row0, row1) window: generate the rows at their GLOBAL coordinates, invert, pick every valid
interior pixel with converged == 0 plus the pixels with the highest evaluation counts, and save their 3x3 input
neighbourhoods (NSPATIAL 2) side by side as a small raster: patch k occupies columns [3k, 3k+3), its centre is
(1, 3k+1). A pixel's inversion depends on its neighbourhood and its own DEPTHS prior only, so the centre of a patch
reproduces the mined pixel exactly. Output: gpurun_out/mined_<config>.npz (planes, prior, global coordinates, the
device's n_evals / converged for the centres). tests/golden/make_golden.py turns these into reference goldens.
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from photic_b200 import capi, scene
from photic_b200.samodel import Inverter

MAX_NONCONV, MAX_HIGH = 96, 48


def mine(inv, name, row0, row1):
    spec = scene.CONFIGS[name]
    planes, prior = scene.generate(spec, row0, row1, device="cuda")
    R = row1 - row0
    desc = capi.desc_from_spec(spec, nrows=R)
    o = Inverter.alloc_device_outputs(desc, "cuda", scene_planes=False)
    st = inv.invert_device(desc, planes, prior, o)
    torch.cuda.synchronize()
    valid = scene.valid_mask(planes)
    interior = torch.zeros_like(valid)
    interior[1:-1, 1:-1] = True
    ne, cv = o["n_evals"], o["converged"]
    cand = valid & interior
    nonconv = torch.nonzero((cv == 0) & cand)
    hist = torch.bincount(torch.clamp(ne[valid] // 250, max=40), minlength=41).cpu().numpy().tolist()
    # spread the non-converged picks over the window, take the highest evaluation counts among the converged
    if len(nonconv) > MAX_NONCONV:
        nonconv = nonconv[torch.linspace(0, len(nonconv) - 1, MAX_NONCONV).long()]
    hi_mask = (cv == 1) & cand & (ne > 3000)
    hi = torch.nonzero(hi_mask)
    if len(hi) > 0:
        order = torch.argsort(ne[hi[:, 0], hi[:, 1]], descending=True)[:MAX_HIGH]
        hi = hi[order]
    pick = torch.cat([nonconv, hi], dim=0)
    K = len(pick)
    SB = planes.shape[0]
    pp = torch.empty((SB, 3, 3 * K), dtype=torch.float32, device="cuda")
    pr = torch.empty((3, 3 * K), dtype=torch.float32, device="cuda")
    for k in range(K):
        i, j = int(pick[k, 0]), int(pick[k, 1])
        pp[:, :, 3 * k:3 * k + 3] = planes[:, i - 1:i + 2, j - 1:j + 2]
        pr[:, 3 * k:3 * k + 3] = prior[i - 1:i + 2, j - 1:j + 2]
    pi, pj = pick[:, 0], pick[:, 1]
    info = {"config": name, "rows": [row0, row1], "n_valid": int(st["n_valid"]), "n_converged": int(st["n_converged"]),
            "n_nonconverged_interior": int(((cv == 0) & cand).sum()), "picked_nonconverged": int(len(nonconv)),
            "picked_high": int(len(hi)), "max_evals": int(ne.max()), "evals_hist_250": hist,
            "px_per_s": st["n_valid"] / (st["ms_solve"] * 1e-3), "evals_per_px": st["n_evals"] / max(1, st["n_valid"])}
    print(json.dumps(info), flush=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"mined_{name}.npz"), planes=pp.cpu().numpy(), prior=pr.cpu().numpy(),
                        gi=(pi + row0).cpu().numpy().astype(np.int32), gj=pj.cpu().numpy().astype(np.int32),
                        dev_evals=ne[pi, pj].cpu().numpy(), dev_converged=cv[pi, pj].cpu().numpy(),
                        dev_depth=o["depth"][pi, pj].cpu().numpy(), config=name, info=json.dumps(info))
    del planes, prior, o
    torch.cuda.empty_cache()


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    inv = Inverter(0)
    a = sys.argv[1:]
    for k in range(0, len(a), 3):
        mine(inv, a[k], int(a[k + 1]), int(a[k + 2]))
