"""Cost of a pixel inverted as sand-only (prior > 8 m, Nb = 1) vs with all substrates (Nb = NBOTTOMS): same raster,
the DEPTHS prior forced to -20 m / -3 m everywhere. Calibrates sharded.row_cost's shallow weight."""
import sys, json
sys.path.insert(0, ".")
import torch
from photic_b200 import scene, capi
from photic_b200.samodel import Inverter
name = sys.argv[1] if len(sys.argv) > 1 else "exmouth"
spec = scene.CONFIGS[name].scaled(500, 600)
planes, prior = scene.generate(spec, device="cuda")
desc = capi.desc_from_spec(spec)
inv = Inverter(0)
outs = Inverter.alloc_device_outputs(desc, "cuda", scene_planes=False)
res = {}
for tag, h in (("deep", -20.0), ("shallow", -3.0)):
    pr = torch.where(prior == scene.NODATA, prior, torch.full_like(prior, h))
    for _ in range(2):
        st = inv.invert_device(desc, planes, pr, outs)
    res[tag] = {"us_per_px": 1e3 * st["ms_solve"] / st["n_valid"], "evals_per_px": st["n_evals"] / st["n_valid"], "n": st["n_valid"]}
res["ratio"] = res["shallow"]["us_per_px"] / res["deep"]["us_per_px"]
print(json.dumps({name: res}))
