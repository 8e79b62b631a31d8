#!/usr/bin/env python3
"""One whole BASELINE.json scene through the sharded path, timed end to end on the devices.

    python tests/manual/run_scene.py --config exmouth                                   (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \\
        tests/manual/run_scene.py --config pilbara [--refine]                            (8 GPUs)

Every rank starts with an equal split of the scene's rows resident in HBM (synthetic, generated in place). Timed:
halo exchange between neighbouring bands (NCCL point to point over NVLink) -> the band's rasters into its shareable
allocation -> validity scan + work queues -> per-band inversion, every device taking pixels from its neighbours' queues
once its own is empty (no collective and no cost model on the data path) -> [REFINE with all-reduced min/max] -> gather
of the nine result planes on rank 0. --no-share: cost-balanced contiguous bands (round 1), every rank on its own band.
Rank 0 prints one JSON line (also written to gpurun_out/ when that directory exists).
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from photic_b200 import capi, scene, sharded
from photic_b200.samodel import Inverter

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="exmouth")
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--cols", type=int, default=0)
ap.add_argument("--refine", action="store_true", help="REFINE SCALE+POWER on the depth plane (global min/max all-reduced)")
ap.add_argument("--check", type=int, default=0, help="pixels per rank to check against the CPU oracle (bit equality)")
ap.add_argument("--no-share", action="store_true", help="round-1 path: cost-balanced bands, no work sharing")
ap.add_argument("--tag", default="")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
inv = Inverter(local)
spec = scene.CONFIGS[args.config]
if args.rows and args.cols:
    spec = spec.scaled(args.rows, args.cols)
halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
eq = sharded.plan_row_bands(np.ones(spec.nrows), world)
e0, e1 = eq[rank]
t_gen = time.time()
planes_eq, prior_eq = scene.generate(spec, e0, e1, device=dev)
torch.cuda.synchronize()
t_gen = time.time() - t_gen
peak, _ = inv.fp64_peak()
if world > 1:
    dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
torch.cuda.synchronize()
wall0 = time.time()
ev[0].record()
if args.no_share:
    # 1. cost estimate -> balanced bands; 2. re-deal rows
    cost = torch.zeros(spec.nrows, device=dev)
    cost[e0:e1] = sharded.row_cost_from_prior(scene.valid_mask(planes_eq), prior_eq)
    if world > 1:
        dist.all_reduce(cost)
    plan = sharded.plan_row_bands(cost.cpu().numpy(), world)
    planes = sharded.repartition_rows(planes_eq, eq, plan, rank, world)
    prior = sharded.repartition_rows(prior_eq[None], eq, plan, rank, world)[0]
else:
    plan, planes, prior = eq, planes_eq, prior_eq   # equal rows: the devices share the work at run time
r0, r1 = plan[rank]
del planes_eq, prior_eq
w0, w1, lb, le = sharded.window(r0, r1, halo, spec.nrows)
desc = capi.desc_from_spec(spec, nrows=w1 - w0)
group = sharded.BandGroup(inv, desc, lb, le, rank, world, share=not args.no_share)   # allocation + handle exchange
band = group.band
# 2. halo rows from the neighbouring bands, straight into the band's rasters
band.planes.copy_(sharded.exchange_halo(planes, plan, halo, rank, world))
band.prior.copy_(sharded.exchange_halo(prior[None], plan, halo, rank, world)[0])
win, prw = band.planes, band.prior
del planes, prior
ev[1].record()
# 3. inversion of this band (+ pixels of the neighbours' bands once its own queue is empty)
st = group.step()
own_valid = band.valid()
outs = band.outputs
ev[2].record()
# 4. optional REFINE on the depth plane (refine.c:215-301): global min/max over all bands
if args.refine:
    import ctypes as C
    d = outs["depth"][lb:le].contiguous()
    mm = np.zeros(2, dtype=np.float32)
    capi.check(inv.lib.phb_refine_minmax_device(inv.ctx, C.c_void_p(d.data_ptr()), d.numel(), C.c_float(0.0), mm.ctypes.data_as(capi._fp), None))
    lo, hi = (float(mm[0]), float(mm[1]))
    if world > 1:
        lo, hi = sharded.allreduce_minmax(lo, hi, dev)
    mm[:] = (lo, hi)
    rargs = np.array([0, 0, -40.0, 0.0, 1.2, 1, 0, 0, 0, 1.05, 0.98], dtype=np.float32)
    refined = torch.empty_like(d)
    capi.check(inv.lib.phb_refine_device(inv.ctx, C.c_void_p(d.data_ptr()), C.c_float(0.0), None, C.c_float(0), None, C.c_float(0),
                                         d.numel(), 2 | 16, rargs.ctypes.data_as(capi._fp), mm.ctypes.data_as(capi._fp),
                                         C.c_void_p(refined.data_ptr()), None))
ev[3].record()
# 5. gather the nine planes on rank 0
stack = torch.stack([outs[n][lb:le] for n in capi.SCALAR_PLANES])
full = sharded.gather_bands(stack, plan, rank, world)
ev[4].record()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.time() - wall0
ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(4)]
t = torch.tensor([ev[0].elapsed_time(ev[4]), ms[0], ms[1], ms[2], ms[3], st["ms_solve"]], dtype=torch.float64, device=dev)
agg = torch.tensor([float(st["n_valid"]), st["alg_flops"], float(st["n_evals"]), float(st["n_converged"]), float(st["n_shallow"]),
                    float(own_valid)], dtype=torch.float64, device=dev)
share_t = torch.tensor([float(st["n_valid"]) - float(own_valid)], dtype=torch.float64, device=dev)  # pixels taken from (+) / given to (-) neighbours
share_all = [torch.zeros_like(share_t) for _ in range(world)]
if world > 1:
    dist.all_gather(share_all, share_t)
else:
    share_all = [share_t]
tmin = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(agg)
ok = None
if args.check > 0:
    from oracle.binding import Oracle, SceneCfg
    vm = scene.valid_mask(win[:, lb:le])
    idx = torch.nonzero(vm.reshape(-1)).reshape(-1).cpu().numpy()
    pick = idx[np.random.default_rng(rank).permutation(len(idx))[:args.check]]
    ii, jj = pick // spec.ncols + lb, pick % spec.ncols
    ref = Oracle("port").invert_pixels(SceneCfg.from_spec(spec), win.cpu().numpy(), scene.NODATA, prw.cpu().numpy(), scene.NODATA, ii, jj)
    got = outs["depth"].cpu().numpy()[ii, jj]
    okl = bool(np.array_equal(got.view(np.int32), (-ref["rec"][:, 0].astype(np.float32)).view(np.int32))
               and np.array_equal(outs["n_evals"].cpu().numpy()[ii, jj], ref["n_evals"]))
    okt = torch.tensor([1.0 if okl else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    ok = bool(okt[0] > 0)
if rank == 0:
    total_ms = float(t[0]); npx = float(agg[0])
    line = {"what": "whole scene, sharded", "config": args.config, "scene": f"{spec.nrows}x{spec.ncols}, {spec.n_dates} dates", "n_gpus": world,
            "valid_pixels": int(npx), "px_per_s": npx / (total_ms * 1e-3), "seconds_device": total_ms * 1e-3, "seconds_wall": wall,
            "work_sharing": bool(st.get("shared", False)), "pixels_taken_from_neighbours_per_rank": [int(x[0]) for x in share_all],
            "phases_ms_max_over_ranks": {"plan+redeal+halo": float(t[1]), "invert": float(t[2]), "refine": float(t[3]), "gather": float(t[4]),
                                         "solve_kernel": float(t[5])},
            "solve_kernel_ms_min_over_ranks": float(tmin[5]), "band_balance": float(tmin[5]) / float(t[5]),
            "alg_tflops_total": float(agg[1]) / (float(t[5]) * 1e-3) / 1e12, "fp64_peak_tflops_per_gpu": peak,
            "frac_of_fp64_peak": float(agg[1]) / (float(t[5]) * 1e-3) / 1e12 / (peak * world),
            "evals_per_pixel": float(agg[2]) / npx, "converged_fraction": float(agg[3]) / npx, "shallow_fraction": float(agg[4]) / npx,
            "refine": bool(args.refine), "gathered_shape": list(full.shape) if full is not None else None,
            "oracle_check_bit_identical": ok, "scene_generation_s": t_gen, "plan": plan}
    print(json.dumps(line))
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", f"scene_{args.config}_n{world}{args.tag}.json"), "w") as f:
            f.write(json.dumps(line) + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
