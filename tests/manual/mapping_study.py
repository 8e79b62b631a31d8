#!/usr/bin/env python3
"""Mapping study: objective evaluations per second with one pixel per warp (the product) against one pixel per team of
2 / 4 / 8 co-operating warps (photic_b200/csrc/aux_kernels.cuh: objective_team, eval_bench_kernel), 16 warps per SM on
every SM either way, each evaluation depending on the previous one as in the simplex. Prints one JSON line per
configuration; with --ncu-list only the configurations that are worth an ncu capture (one launch each) are run.

    python tests/manual/mapping_study.py [--reps 4000]
    ncu --set full --clock-control none -k regex:eval_bench -o gpurun_out/r02_mapping python tests/manual/mapping_study.py --ncu-list --reps 600
"""
import argparse, json, os, sys
from dataclasses import replace
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from photic_b200 import capi, scene
from photic_b200.samodel import Inverter

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=4000)
ap.add_argument("--ncu-list", action="store_true")
ap.add_argument("--skews", default="0,900", help="start-time skew between the teams of an SM, cycles (0: all in step)")
args = ap.parse_args()
k = np.load(os.path.join(ROOT, "tests", "golden", "kat_objective.npz"))
inv = Inverter(0)
# golden b: 6 dates, sand only, 9 regions (n = 45: the class of 79 % of the Exmouth pixels); a: 4 dates, 3 substrates,
# 9 regions (n = 75); c: 8 dates, 3 substrates, 6 regions (n = 66)
cases = [("b", "6 dates, sand only, 9 regions"), ("a", "4 dates, 3 substrates, 9 regions"), ("c", "8 dates, 3 substrates, 6 regions")]
maps = [(1, False), (2, False), (2, True), (4, False), (4, True), (8, False)]
if args.ncu_list:
    cases, maps = cases[:1], [(1, False), (2, True), (4, False), (4, True)]
for tag, what in cases:
    ns, nb, nr, origin = (int(v) for v in k[f"{tag}_meta"])
    desc = capi.desc_from_spec(replace(scene.CONFIGS["murion"], n_dates=ns))
    for skew in [int(v) for v in args.skews.split(",")]:
        base = None
        for tw, same in maps:
            # one evaluation takes ~14 k cycles: `skew` cycles per team spread the teams of an SM over an evaluation
            sk = skew * tw
            if not args.ncu_list:
                inv.eval_bench(desc, nb, nr, origin, k[f"{tag}_meas"], k[f"{tag}_params"][0], tw, same, reps=50)  # warm-up
            first, rate, ms = inv.eval_bench(desc, nb, nr, origin, k[f"{tag}_meas"], k[f"{tag}_params"][0], tw, same, reps=args.reps,
                                             skew_cycles=sk)
            ok = np.float64(first).view(np.int64) == np.float64(k[f"{tag}_out"][0, 0]).view(np.int64)
            if tw == 1:
                base = rate
            print(json.dumps({"case": what, "warps_per_pixel": tw, "team_on_one_scheduler": same, "pixels_in_flight_per_sm": 16 // tw,
                              "skew_cycles_between_teams": sk, "evals_per_s": rate, "vs_warp_per_pixel": rate / base if base else None,
                              "ms": ms, "reps": args.reps, "first_value_equals_reference": bool(ok)}), flush=True)
