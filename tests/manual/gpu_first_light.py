"""First-light check on a GPU box: exact-math KAT, per-pixel parity vs the CPU oracle, timings."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from photic_b200 import scene, capi
from photic_b200.samodel import Inverter
from oracle.binding import Oracle, SceneCfg

inv = Inverter(0)
print("fp64 peak TFLOP/s, ms:", inv.fp64_peak(), flush=True)
rng = np.random.default_rng(0)
import ctypes as C
host = C.CDLL("tests/_build/libexactmath_host.so")
for f in (host.phm_host_exp, host.phm_host_log, host.phm_host_pow): f.restype = C.c_double
n = 200000
for fn, x, y in ((0, rng.uniform(-800, 20, n), None), (1, np.exp(rng.uniform(-30, 5, n)), None),
                 (2, rng.uniform(0.3, 1.5, n), rng.uniform(-3, 3, n))):
    got = inv.kat_math(fn, x, y)
    import math
    ref = np.array([math.exp(v) for v in x]) if fn == 0 else (np.array([math.log(v) for v in x]) if fn == 1 else np.array([math.pow(a, b) for a, b in zip(x, y)]))
    print("math fn", fn, "mismatch vs host libm:", int((got.view(np.int64) != ref.view(np.int64)).sum()), flush=True)

name, R, Cc = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spec = scene.CONFIGS[name].scaled(R, Cc)
planes, prior = scene.generate(spec)
desc = capi.desc_from_spec(spec)
t = time.time()
out, st = inv.invert_host(desc, planes.numpy(), prior.numpy(), debug=True)
print("gpu %.2fs" % (time.time() - t), json.dumps(st), flush=True)
pix = out["pix"]; ii, jj = pix // spec.ncols, pix % spec.ncols
cfg = SceneCfg.from_spec(spec)
port = Oracle("port")
npx = min(len(pix), int(sys.argv[4]) if len(sys.argv) > 4 else 2000)
sel = np.linspace(0, len(pix) - 1, npx).astype(int)
t = time.time()
a = port.invert_pixels(cfg, planes.numpy(), scene.NODATA, prior.numpy(), scene.NODATA, ii[sel], jj[sel], nthreads=0)
print("oracle %.2fs for %d px" % (time.time() - t, npx), flush=True)
g = out["rec"][sel]
same = (g.view(np.int64) == a["rec"].view(np.int64)).all(axis=1)
print("records bit-identical: %d / %d" % (same.sum(), npx))
print("evals identical:", int((out["rec_evals"][sel] == a["n_evals"]).sum()), "conv identical:", int((out["rec_converged"][sel] == a["converged"]).sum()))
if not same.all():
    k = np.nonzero(~same)[0][0]
    print("first mismatch pixel", ii[sel][k], jj[sel][k]); print(" gpu", g[k]); print(" cpu", a["rec"][k])
    print("max |dH|", np.abs(g[:, 0] - a["rec"][:, 0]).max())
print("px/s solve-only:", st["n_valid"] / (st["ms_solve"] * 1e-3), "alg TFLOP/s:", st["alg_flops"] / (st["ms_solve"] * 1e-3) / 1e12)
