"""Small fixed workload for ncu: one inversion of an Exmouth-shaped (6-date) raster."""
import sys, json
sys.path.insert(0, ".")
from photic_b200 import scene, capi
from photic_b200.samodel import Inverter
import torch
R, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 200)
name = sys.argv[3] if len(sys.argv) > 3 else "exmouth"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
spec = scene.CONFIGS[name].scaled(R, C)
planes, prior = scene.generate(spec, device="cuda")
desc = capi.desc_from_spec(spec)
inv = Inverter(0)
outs = Inverter.alloc_device_outputs(desc, "cuda", scene_planes=False)
for _ in range(reps):
    st = inv.invert_device(desc, planes, prior, outs)
torch.cuda.synchronize()
print(json.dumps(st))
print("px/s", st["n_valid"] / (st["ms_solve"] * 1e-3), "TFLOP/s", st["alg_flops"] / (st["ms_solve"] * 1e-3) / 1e12)
