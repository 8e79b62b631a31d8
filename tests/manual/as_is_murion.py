"""BASELINE.json configs[0] on the CPU: the reference's samodel() EXACTLY as shipped (LUT + hot start + its own OpenMP
team, oracle/_ref) on the Murion-shaped 1040x305, 4-date scene, wall clock, beside the per-pixel cold-start harness
(the like-for-like path the CUDA kernel reproduces) on a pixel subsample. CPU only; needs oracle/_ref.
    python tests/manual/as_is_murion.py [rows cols]   -> one JSON line (stdout of the reference goes to /dev/null)
The as-is output is schedule dependent: it is timed, never compared."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle.binding import Oracle, SceneCfg  # noqa: E402
from photic_b200 import scene  # noqa: E402

spec = scene.CONFIGS["murion"]
if len(sys.argv) > 2:
    spec = spec.scaled(int(sys.argv[1]), int(sys.argv[2]))
planes, prior = scene.generate(spec)
pl, pr = planes.numpy(), prior.numpy()
cfg = SceneCfg.from_spec(spec)
valid = scene.valid_mask(planes).numpy()
n_valid = int(valid.sum())
ref = Oracle("reference")
cores = len(os.sched_getaffinity(0))

sys.stdout.flush()
saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
os.dup2(devnull, 1)
try:
    t0 = time.perf_counter()
    out = ref.samodel_as_is(cfg, pl, scene.NODATA, pr, scene.NODATA)
    t_as_is = time.perf_counter() - t0
finally:
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(devnull)

ii, jj = np.nonzero(valid)
k = max(1, len(ii) // 4000)
sel = np.arange(0, len(ii), k)
t0 = time.perf_counter()
res = ref.invert_pixels(cfg, pl, scene.NODATA, pr, scene.NODATA, ii[sel], jj[sel], nthreads=cores)
t_cold = time.perf_counter() - t0
print(json.dumps({
    "workload": f"murion {spec.nrows}x{spec.ncols}, {spec.n_dates} dates (BASELINE.json configs[0]), DEPTHS prior, "
                "NSPATIAL=2 NSMOOTH=1 NBOTTOMS=3", "host_cores": cores, "valid_px": n_valid,
    "samodel_as_is": {"wall_s": t_as_is, "px_per_s_whole_call": n_valid / t_as_is,
                      "note": "LUT + hot start + depth-error phase + per-thread redundant phases, gcc -O3 -fopenmp; "
                              "inverted cells: %d" % int((out[0] != 0).sum())},
    "cold_start_per_pixel": {"sample_px": int(len(sel)), "every_kth_valid_pixel": int(k), "wall_s": t_cold,
                             "px_per_s": len(sel) / t_cold, "mean_evals": float(res["n_evals"].mean())}}))
