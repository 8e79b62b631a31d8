"""Row-band balance of the static cost models on ONE GPU: plan W bands of a scaled scene, invert each band by
itself, compare kernel times (min/max = the balance a W-GPU run would see)."""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from photic_b200 import scene, capi, sharded
from photic_b200.samodel import Inverter
name = sys.argv[1] if len(sys.argv) > 1 else "pilbara"
R, C, W = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
spec = scene.CONFIGS[name].scaled(R, C)
planes, prior = scene.generate(spec, device="cuda")
valid = scene.valid_mask(planes)
inv = Inverter(0)
halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
models = {"valid, shallow x3": sharded.row_cost(valid, prior.abs() <= 8.0), "depth-binned": sharded.row_cost_from_prior(valid, prior)}
for mname, cost in models.items():
    plan = sharded.plan_row_bands(cost.cpu().numpy(), W)
    ms = []
    for r0, r1 in plan:
        w0, w1, lb, le = sharded.window(r0, r1, halo, spec.nrows)
        d = capi.desc_from_spec(spec, nrows=w1 - w0)
        o = Inverter.alloc_device_outputs(d, "cuda", scene_planes=False)
        st = inv.invert_device(d, planes[:, w0:w1].contiguous(), prior[w0:w1].contiguous(), o, row_begin=lb, row_end=le)
        ms.append(st["ms_solve"])
    print(json.dumps({"scene": f"{name} {R}x{C}", "bands": W, "model": mname, "balance_min_over_max": min(ms) / max(ms),
                      "efficiency_mean_over_max": float(np.mean(ms) / max(ms)), "ms": [round(m, 1) for m in ms]}))
