#!/usr/bin/env python3
"""bench.py -- pixel inversions/s of the per-pixel semi-analytical inversion (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3                      (our arm, one B200)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               (the reference's CPU code)

Workload (config.workload): BASELINE.json configs[1], the Exmouth-Gulf-shaped 3930x2858, 6-date
synthetic Landsat-8 scene, cut into `--batches` row batches. One STEP inverts one batch per GPU
(weak scaling: N GPUs take N batches per step, stacked into one raster that is sharded by row bands
with a real NCCL halo exchange and a final gather of the output planes to rank 0). Every step sees
a different batch (134 MB of reflectance planes per batch > the 126 MB L2), so nothing is cached
between timed iterations.

Printed JSON line (rank 0): `value` = valid pixels inverted / device time with inputs resident in
HBM; `e2e` = the same through the host-buffer C-ABI call (pinned host memory in, host planes out,
H2D and D2H inside the timed region); `roofline` = algorithmic FP64 FLOP/s of the solve kernel
against the DFMA peak measured on this device in this run (the binding roof is the FP64 pipe, not
HBM: ~430 B vs ~15 MFLOP per pixel; MEASURED_PEAKS.json has no FP64 figure); `cpu_baseline` = the
reference's own CPU code (oracle/_ref, else the oracle port) on a bounded pixel sample on this host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pixel inversions/s"
UNIT = "px/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="exmouth")
    ap.add_argument("--batches", type=int, default=0,
                    help="row batches the scene is cut into (one per GPU per step); default = --steps, so that the "
                         "timed steps cover every batch of the scene exactly N times at N GPUs")
    ap.add_argument("--rows", type=int, default=0, help="debug: shrink the scene")
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 8 for k in range(4) if r[4 + k].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def scene_spec(args):
    from photic_b200 import scene
    spec = scene.CONFIGS[args.config]
    if args.rows and args.cols:
        spec = spec.scaled(args.rows, args.cols)
    return spec


def batch_rows(spec, args, b):
    rb = -(-spec.nrows // args.batches)
    r0 = (b % args.batches) * rb
    return r0, min(spec.nrows, r0 + rb)


def cpu_baseline(spec, args, planes_np, prior_np, target_s, steps=1):
    """The reference's CPU inversion on a bounded, deterministic pixel sample (every k-th valid pixel)."""
    from oracle.binding import REF_SO, Oracle, SceneCfg
    from photic_b200 import scene
    import torch
    kind = "reference" if os.path.exists(REF_SO) else "port"
    orc = Oracle(kind)
    cores = len(os.sched_getaffinity(0))
    vm = scene.valid_mask(torch.from_numpy(planes_np)).numpy()
    ii, jj = np.nonzero(vm)
    n = int(max(64, min(len(ii), cores * 25 * target_s)))  # ~25 px/s/core at 6 dates
    k = max(1, len(ii) // n)
    sel = np.arange(0, len(ii), k)[:n]
    cfg = SceneCfg.from_spec(spec)
    times = []
    for _ in range(steps):
        t = time.time()
        out = orc.invert_pixels(cfg, planes_np, scene.NODATA, prior_np, scene.NODATA, ii[sel], jj[sel], nthreads=cores)
        times.append(time.time() - t)
    pxs = len(sel) / float(np.mean(times))
    return {"value": pxs, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"every {k}th valid pixel of batch 0 ({len(sel)} px, {np.mean(times):.1f} s/step, "
                      f"omp schedule(dynamic), gcc -O3, mean {out['n_evals'].mean():.0f} evals/px)"}, times, len(sel)


def run_reference(args, emit=print):
    """--impl reference: the reference's own CPU implementation, all host threads, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from photic_b200 import scene
    spec = scene_spec(args)
    r0, r1 = batch_rows(spec, args, 0)
    planes, prior = scene.generate(spec, r0, r1)
    planes_np, prior_np = planes.numpy(), prior.numpy()
    sub = spec.scaled(r1 - r0, spec.ncols)
    per_step = max(4.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    base, times, npx = cpu_baseline(sub, args, planes_np, prior_np, per_step, steps=args.warmup + args.steps)
    t = times[args.warmup:]
    val = npx * len(t) / float(np.sum(t))
    base["value"] = val
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(t)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(spec, args, 1, extra={"sample_px_per_step": npx}),
        "cpu_baseline": base, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(spec, args, world, extra=None):
    rb = -(-spec.nrows // args.batches)
    c = {"workload": f"{spec.name} {spec.nrows}x{spec.ncols}, {spec.n_dates} dates x {spec.n_bands} bands "
                     f"(BASELINE.json configs[1]), NSPATIAL={spec.n_spatial} NSMOOTH={spec.n_smoothing_radius} "
                     f"NBOTTOMS={spec.n_bottoms}, DEPTHS prior",
         "step": f"one batch of {rb} rows x {spec.ncols} cols per GPU ({args.batches} batches per scene), a new batch every step",
         "l2": "inputs larger than L2: every step reads a different batch (>=134 MB of planes)",
         "parallelism": f"row bands x{world}" + (", cost-balanced re-deal + halo exchange (NCCL p2p) + gather" if world > 1 else "")}
    if extra:
        c.update(extra)
    return c


def main():
    args = parse()
    if args.batches <= 0:
        args.batches = max(1, args.steps)
    # stdout carries exactly ONE line, the JSON result: everything else that libraries print there (NCCL's
    # "NCCL version ..." banner, OpenMP notices) is sent to stderr for the duration of the run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    from photic_b200 import capi, scene, sharded
    from photic_b200.samodel import Inverter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; photic_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    inv = Inverter(local)
    spec = scene_spec(args)
    K, W = args.steps, args.warmup
    halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
    rb = -(-spec.nrows // args.batches)
    plan_eq = [(r * rb, (r + 1) * rb) for r in range(world)]  # how the stacked per-step raster arrives

    # ---- resident inputs: the batches this rank will see, generated on the device ----------------
    def batch_of(step, r=None):
        return (step * world + (rank if r is None else r)) % args.batches

    need = sorted({batch_of(s) for s in range(W + K)})
    data, cost = {}, {}
    for b in need:
        r0, r1 = batch_rows(spec, args, b)
        p, pr = scene.generate(spec, r0, r1, device=dev)
        if r1 - r0 < rb:  # last batch of the scene may be short: pad with nodata rows (land)
            padp = torch.full((p.shape[0], rb - (r1 - r0), spec.ncols), scene.NODATA, device=dev)
            p = torch.cat([p, padp], dim=1).contiguous()
            pr = torch.cat([pr, torch.full((rb - (r1 - r0), spec.ncols), scene.NODATA, device=dev)], dim=0).contiguous()
        data[b] = (p, pr)
        # estimated work per row: a shallow-water pixel (all substrates) costs ~3x a sand-only one
        cost[b] = sharded.row_cost_from_prior(scene.valid_mask(p), pr)
    torch.cuda.synchronize()
    peak_tflops, _ = inv.fp64_peak()
    gather_names = capi.SCALAR_PLANES
    out_cache = {}

    def step(s):
        """One step: N batches stacked into one raster; rows are re-dealt to cost-balanced bands (NCCL p2p),
        halo rows exchanged, every rank inverts its band, the 9 result planes are gathered on rank 0."""
        p, pr = data[batch_of(s)]
        plan = plan_eq
        if world > 1:
            c = torch.zeros(rb * world, device=dev)
            c[rank * rb:(rank + 1) * rb] = cost[batch_of(s)]
            dist.all_reduce(c)
            plan = sharded.plan_row_bands(c.cpu().numpy(), world)
            p = sharded.repartition_rows(p, plan_eq, plan, rank, world)
            pr = sharded.repartition_rows(pr[None], plan_eq, plan, rank, world)[0]
        win = sharded.exchange_halo(p, plan, halo, rank, world)
        prw = sharded.exchange_halo(pr[None], plan, halo, rank, world)[0]
        w0, w1, lb, le = sharded.window(plan[rank][0], plan[rank][1], halo, rb * world)
        if w1 - w0 not in out_cache:
            d = capi.desc_from_spec(spec, nrows=w1 - w0)
            out_cache[w1 - w0] = (d, Inverter.alloc_device_outputs(d, dev, scene_planes=False))
        desc, outs = out_cache[w1 - w0]
        st = inv.invert_device(desc, win.contiguous(), prw.contiguous(), outs, row_begin=lb, row_end=le)
        if world > 1:
            stack = torch.stack([outs[n][lb:le] for n in gather_names])
            sharded.gather_bands(stack, plan, rank, world)
        return st

    for s in range(W):
        step(s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    stats = [step(W + s) for s in range(K)]
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    px = torch.tensor([float(sum(s["n_valid"] for s in stats))], dtype=torch.float64, device=dev)
    agg = torch.tensor([sum(s["alg_flops"] for s in stats), sum(s["ms_solve"] for s in stats),
                        float(sum(s["n_evals"] for s in stats)), float(sum(s["n_iters"] for s in stats)),
                        float(sum(s["n_shallow"] for s in stats)), float(sum(s["n_converged"] for s in stats))],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(px)
        dist.all_reduce(agg)
    total_ms, total_px = float(ms[0]), float(px[0])
    value = total_px / (total_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hp, hpr = {}, {}
        for b in need:
            p, pr = data[b]
            hp[b] = torch.empty(p.shape, dtype=torch.float32, pin_memory=True).copy_(p).numpy()
            hpr[b] = torch.empty(pr.shape, dtype=torch.float32, pin_memory=True).copy_(pr).numpy()
        hdesc = capi.desc_from_spec(spec, nrows=rb)
        hbuf = {n: torch.empty((rb, spec.ncols), dtype=torch.float32, pin_memory=True).numpy() for n in capi.SCALAR_PLANES}
        hbuf["converged"] = torch.empty((rb, spec.ncols), dtype=torch.uint8, pin_memory=True).numpy()
        hbuf["n_evals"] = torch.empty((rb, spec.ncols), dtype=torch.int32, pin_memory=True).numpy()

        def host_step(s):
            b = batch_of(s)
            _, st = inv.invert_host(hdesc, hp[b], hpr[b], scene_planes=False, buffers=hbuf)
            return st

        host_step(0)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        hst = [host_step(W + s) for s in range(K)]
        t1 = time.perf_counter()
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        hpx = torch.tensor([float(sum(s["n_valid"] for s in hst))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(hpx)
        plane_bytes = rb * spec.ncols * 4
        e2e = {"value": float(hpx[0]) / float(tt[0]), "unit": UNIT,
               "h2d_bytes_per_step": world * plane_bytes * (spec.n_planes + 1),
               "d2h_bytes_per_step": world * (plane_bytes * len(capi.SCALAR_PLANES) + rb * spec.ncols * 5),
               "ms_per_step": 1e3 * float(tt[0]) / K}

    if rank == 0:
        alg_flops, ms_solve = float(agg[0]), float(agg[1]) / world
        achieved = alg_flops / (ms_solve * 1e-3) / 1e12 / world  # per-GPU TFLOP/s of the solve kernel
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(spec, args, world),
            "pixels_per_step": total_px / K,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "traffic": None,
                         "peak_source": "DFMA-chain kernel measured on this device in this run (phb_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry; B200 spec 37 TFLOP/s",
                         "alg_flops_per_pixel": alg_flops / total_px, "evals_per_pixel": float(agg[2]) / total_px,
                         "iters_per_pixel": float(agg[3]) / total_px, "kernel": "phb::solve_kernel",
                         "kernel_ms_per_step": ms_solve / K, "hbm_gbs_algorithmic": None},
            "stats": {"shallow_fraction": float(agg[4]) / total_px, "converged_fraction": float(agg[5]) / total_px,
                      "warps_per_cta": stats[0]["warps_per_cta"], "ctas": stats[0]["ctas"],
                      "smem_bytes": stats[0]["smem_bytes"], "regs": stats[0]["regs"]},
            "e2e": e2e, "gpu_launches": 3 * K * world, "clocks": clocks,
        }
        # measured DRAM traffic of the kernel: bytes per pixel from the committed `ncu --set full` capture x pixels per launch
        tpath = os.path.join(ROOT, "profiles", "r01_solve_kernel_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            line["roofline"]["traffic"] = tj["dram_bytes_per_pixel"] * total_px / (K * world)
            line["roofline"]["traffic_source"] = "profiles/r01_solve_kernel_traffic.json (ncu dram__bytes_read+write per pixel) x pixels per launch"
        # algorithmic HBM traffic (reported, not binding): planes + prior in, 9 planes + flags out
        bytes_px = spec.n_planes * 4 + 4 + 9 * 4 + 5
        line["roofline"]["hbm_gbs_algorithmic"] = bytes_px * rb * spec.ncols * K * world / (total_ms * 1e-3) / 1e9
        if not args.no_cpu_baseline and world == 1:
            p, pr = data[need[0]]
            base, _, _ = cpu_baseline(spec.scaled(rb, spec.ncols), args, p.cpu().numpy(), pr.cpu().numpy(), args.cpu_seconds)
            line["cpu_baseline"] = base
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
