"""In-process multi-GPU through the C ABI (phb_invert_host_multi): one Exmouth-shaped raster from HOST buffers on
1 context and on one context per visible device; wall-clock of the whole call (H2D, kernels, D2H), results compared
bit for bit.  usage: python tests/manual/multi_host.py [rows cols [name]]  -> one JSON line"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from photic_b200 import capi, scene  # noqa: E402
from photic_b200.samodel import Inverter  # noqa: E402

R, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1965, 1429)
name = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "exmouth"
MULTI_ONLY = "--multi-only" in sys.argv  # large N: skip the one-device leg, check a row window against one device instead
spec = scene.CONFIGS[name].scaled(R, C)
planes, prior = scene.generate(spec)
pl, pr = planes.numpy(), prior.numpy()
desc = capi.desc_from_spec(spec)
ndev = torch.cuda.device_count()
ivs = [Inverter(k) for k in range(ndev)]
ivs[0].invert_host(capi.desc_from_spec(scene.CONFIGS[name].scaled(64, 64)), pl[:, :64, :64].copy(), pr[:64, :64].copy())  # warm-up
if MULTI_ONLY:
    t1 = time.perf_counter()
    many, stn = Inverter.invert_host_multi(ivs, desc, pl, pr, scene_planes=False)
    t2 = time.perf_counter()
    e = stn["edges"]
    lo, hi = max(0, int(e[1]) - 8), min(R, int(e[1]) + 8)  # 16 rows across the first band boundary, on one device
    part, _ = ivs[0].invert_host(desc, pl, pr, row_begin=lo, row_end=hi, scene_planes=False)
    same = all(np.array_equal(part[k][lo:hi].view(np.uint8), many[k][lo:hi].view(np.uint8)) for k in part)
    per = [p["ms_solve"] for p in stn["per_ctx"]]
    print(json.dumps({
        "workload": f"{name} {R}x{C}, {spec.n_dates} dates, host buffers in and out (pageable numpy arrays)",
        "devices": ndev, "valid_px": stn["n_valid"],
        "multi": {"wall_s": t2 - t1, "px_per_s": stn["n_valid"] / (t2 - t1), "ms_solve_per_band": per,
                  "band_balance_min_over_max": min(per) / max(per), "edges": e.tolist()},
        "rows_checked_against_one_device": [lo, hi], "bit_identical": bool(same)}))
    assert same
    sys.exit(0)
t0 = time.perf_counter()
one, st1 = ivs[0].invert_host(desc, pl, pr, scene_planes=False)
t1 = time.perf_counter()
many, stn = Inverter.invert_host_multi(ivs, desc, pl, pr, scene_planes=False)
t2 = time.perf_counter()
same = all(np.array_equal(one[k].view(np.uint8), many[k].view(np.uint8)) for k in one)
per = [p["ms_solve"] for p in stn["per_ctx"]]
print(json.dumps({
    "workload": f"{name} {R}x{C}, {spec.n_dates} dates, host buffers in and out (pageable numpy arrays)",
    "devices": ndev, "valid_px": st1["n_valid"],
    "one_device": {"wall_s": t1 - t0, "px_per_s": st1["n_valid"] / (t1 - t0), "ms_solve": st1["ms_solve"]},
    "multi": {"wall_s": t2 - t1, "px_per_s": stn["n_valid"] / (t2 - t1), "ms_solve_per_band": per,
              "band_balance_min_over_max": min(per) / max(per), "edges": stn["edges"].tolist()},
    "speedup_wall": (t1 - t0) / (t2 - t1), "bit_identical": bool(same)}))
assert same
