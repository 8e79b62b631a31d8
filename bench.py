#!/usr/bin/env python3
"""bench.py -- pixel inversions/s of the per-pixel semi-analytical inversion (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3                      (our arm, one B200)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               (the reference's CPU code)

Workload (config.workload): BASELINE.json configs[1], the Exmouth-Gulf-shaped 3930x2858, 6-date
synthetic Landsat-8 scene, cut into `--batches` row batches (default 4: 983 rows, 1.2 - 1.8 M valid pixels each).
One STEP inverts one batch -- the SAME raster at every N (strong scaling): at N GPUs the batch is held
as N equal row bands, one per rank, halo rows are exchanged point to point (NCCL), the solve kernels
share the work at run time over NVLink (a device that runs out of pixels takes them from its
neighbours' queues: photic_b200.sharded.BandGroup), and the nine result planes are gathered on rank 0.
Consecutive steps see different batches (>= 270 MB of reflectance planes each > the 126 MB L2), so
nothing is cached between timed iterations.

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
    ap.add_argument("--batches", type=int, default=4,
                    help="row batches the scene is cut into; one batch = the raster of one step, at every N")
    ap.add_argument("--no-share", action="store_true", help="N > 1: every rank works on its own band only")
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


def cpu_sample(planes_np, n, offset=0):
    """every k-th valid pixel, starting at `offset`: distinct pixels for distinct offsets < k"""
    from photic_b200 import scene
    import torch
    vm = scene.valid_mask(torch.from_numpy(planes_np)).numpy()
    ii, jj = np.nonzero(vm)
    n = int(max(64, min(len(ii), n)))
    k = max(1, len(ii) // n)
    sel = np.arange(offset % k, len(ii), k)[:n]
    return ii[sel], jj[sel], k


def cpu_baseline(spec, args, planes_np, prior_np, target_s):
    """The reference's CPU inversion on a bounded, deterministic pixel sample (every k-th valid pixel)."""
    from oracle.binding import REF_SO, Oracle, SceneCfg
    from photic_b200 import scene
    kind = "reference" if os.path.exists(REF_SO) else "port"
    orc = Oracle(kind)
    cores = len(os.sched_getaffinity(0))
    ii, jj, k = cpu_sample(planes_np, cores * 25 * target_s)  # ~25 px/s/core at 6 dates
    t = time.time()
    out = orc.invert_pixels(SceneCfg.from_spec(spec), planes_np, scene.NODATA, prior_np, scene.NODATA, ii, jj, nthreads=cores)
    dt = time.time() - t
    return {"value": len(ii) / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"every {k}th valid pixel of batch 0 ({len(ii)} px, {dt:.1f} s, omp schedule(dynamic), gcc -O3, "
                      f"mean {out['n_evals'].mean():.0f} evals/px)"}


def run_reference(args, emit=print):
    """--impl reference: the reference's own CPU implementation (oracle/_ref: the unmodified samodel.c / asa047.c /
    common.c per-pixel cold start under omp schedule(dynamic)), all host threads, same workload. Every step times a
    different sample of a different batch; the timed steps together cover >= 50 000 distinct pixels."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import REF_SO, Oracle, SceneCfg
    from photic_b200 import scene
    spec = scene_spec(args)
    kind = "reference" if os.path.exists(REF_SO) else "port"
    orc = Oracle(kind)
    cores = len(os.sched_getaffinity(0))
    K, W = args.steps, args.warmup
    per_step = int(min(20000, max(2000, -(-50000 // max(1, K)))))
    data = {}
    for b in sorted({s % args.batches for s in range(W + K)}):
        r0, r1 = batch_rows(spec, args, b)
        planes, prior = scene.generate(spec, r0, r1)
        data[b] = (planes.numpy(), prior.numpy(), spec.scaled(r1 - r0, spec.ncols))
    times, npx, evals = [], [], []
    for s in range(W + K):
        pl, pr, sub = data[s % args.batches]
        ii, jj, k = cpu_sample(pl, per_step if s >= W else min(per_step, 2000), offset=s // args.batches)
        t = time.time()
        out = orc.invert_pixels(SceneCfg.from_spec(sub), pl, scene.NODATA, pr, scene.NODATA, ii, jj, nthreads=cores)
        times.append(time.time() - t); npx.append(len(ii)); evals.append(float(out["n_evals"].mean()))
    t, n = times[W:], npx[W:]
    val = float(np.sum(n)) / float(np.sum(t))
    base = {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{int(np.sum(n))} distinct valid pixels in {K} timed steps ({n[0]} per step, every k-th valid pixel of "
                      f"batch s mod {args.batches}, start offset s div {args.batches}), {np.mean(t):.1f} s/step, "
                      f"omp schedule(dynamic) on {cores} threads, gcc -O3, mean {np.mean(evals[W:]):.0f} evals/px"}
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * float(np.mean(t)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(spec, args, args.gpus),
        "cpu_baseline": base, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(spec, args, world, extra=None):
    rb = -(-spec.nrows // args.batches)
    c = {"workload": f"{spec.name} {spec.nrows}x{spec.ncols}, {spec.n_dates} dates x {spec.n_bands} bands "
                     f"(BASELINE.json configs[1]), NSPATIAL={spec.n_spatial} NSMOOTH={spec.n_smoothing_radius} "
                     f"NBOTTOMS={spec.n_bottoms}, DEPTHS prior",
         "step": f"one batch of {rb} rows x {spec.ncols} cols ({args.batches} batches per scene), the same raster at every "
                 f"GPU count, another batch every step",
         "l2": "inputs larger than L2: consecutive steps read different batches (>= 270 MB of planes each)",
         "parallelism": f"row bands x{world}"}
    if extra:
        c.update(extra)
    return c


def main():
    args = parse()
    args.batches = max(1, args.batches)
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
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")  # one node; the container's hostname may not resolve
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers around the one-process e2e leg (see below)
    inv = Inverter(local)
    spec = scene_spec(args)
    K, W = args.steps, args.warmup
    halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
    rb = -(-spec.nrows // args.batches)          # rows of one batch = the raster of one step
    plan = sharded.equal_row_bands(rb, world)    # equal row bands: the kernels share the work at run time
    r0b, r1b = plan[rank]
    w0, w1, lb, le = sharded.window(r0b, r1b, halo, rb)

    # ---- resident inputs: this rank's rows of every batch it will see, generated on the device --------
    def batch_of(step):
        return step % args.batches

    need = sorted({batch_of(s) for s in range(W + K)})
    data = {}
    for b in need:
        g0, g1 = batch_rows(spec, args, b)
        a0, a1 = min(g1, g0 + r0b), min(g1, g0 + r1b)
        if a1 > a0:
            p, pr = scene.generate(spec, a0, a1, device=dev)
        else:
            p = torch.empty((spec.n_planes, 0, spec.ncols), device=dev)
            pr = torch.empty((0, spec.ncols), device=dev)
        if a1 - a0 < r1b - r0b:  # the last batch of the scene may be short: pad with nodata rows (land)
            padn = (r1b - r0b) - (a1 - a0)
            p = torch.cat([p, torch.full((p.shape[0], padn, spec.ncols), scene.NODATA, device=dev)], dim=1).contiguous()
            pr = torch.cat([pr, torch.full((padn, spec.ncols), scene.NODATA, device=dev)], dim=0).contiguous()
        data[b] = (p, pr)
    torch.cuda.synchronize()
    peak_tflops, _ = inv.fp64_peak()
    gather_names = capi.SCALAR_PLANES
    group = sharded.BandGroup(inv, capi.desc_from_spec(spec, nrows=w1 - w0), lb, le, rank, world,
                              share=not args.no_share)
    band = group.band

    def step(s):
        """One step: the batch is held as N equal row bands; halo rows are exchanged (NCCL p2p), every rank inverts its
        band and then takes pixels from its neighbours' queues, the 9 result planes are gathered on rank 0."""
        p, pr = data[batch_of(s)]
        band.planes.copy_(sharded.exchange_halo(p, plan, halo, rank, world))
        band.prior.copy_(sharded.exchange_halo(pr[None], plan, halo, rank, world)[0])
        st = group.step()
        if world > 1:
            stack = torch.stack([band.outputs[n][lb:le] for n in gather_names])
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
    # N = 1: phb_invert_host. N > 1: phb_invert_host_multi from ONE process (rank 0) over the N devices of the box --
    # the plugin's own multi-GPU entry point, what the samodel() shim calls -- while the other ranks wait.
    # The waiting ranks must wait on the HOST (gloo): an NCCL barrier would leave a spinning kernel on their device, and
    # rank 0's kernels on that device -- another process, another context -- would be time-sliced against it.
    e2e = None
    if not args.no_e2e:
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)
        if rank == 0:
            hp, hpr = {}, {}
            for b in need:
                g0, g1 = batch_rows(spec, args, b)
                p, pr = scene.generate(spec, g0, g1)
                hp[b] = torch.full((spec.n_planes, rb, spec.ncols), scene.NODATA, dtype=torch.float32).pin_memory()
                hpr[b] = torch.full((rb, spec.ncols), scene.NODATA, dtype=torch.float32).pin_memory()
                hp[b][:, : g1 - g0].copy_(p)
                hpr[b][: g1 - g0].copy_(pr)
                hp[b], hpr[b] = hp[b].numpy(), hpr[b].numpy()
            hdesc = capi.desc_from_spec(spec, nrows=rb)
            hbuf = {n: torch.empty((rb, spec.ncols), dtype=torch.float32, pin_memory=True).numpy() for n in capi.SCALAR_PLANES}
            hbuf["converged"] = torch.empty((rb, spec.ncols), dtype=torch.uint8, pin_memory=True).numpy()
            hbuf["n_evals"] = torch.empty((rb, spec.ncols), dtype=torch.int32, pin_memory=True).numpy()
            ivs = [inv] + [Inverter(k) for k in range(1, world)]

            def host_step(s):
                b = batch_of(s)
                if world == 1:
                    _, st = inv.invert_host(hdesc, hp[b], hpr[b], scene_planes=False, buffers=hbuf)
                else:
                    _, st = Inverter.invert_host_multi(ivs, hdesc, hp[b], hpr[b], scene_planes=False, buffers=hbuf)
                return st

            host_step(0)
            t0 = time.perf_counter()
            hst = [host_step(W + s) for s in range(K)]
            t1 = time.perf_counter()
            for iv in ivs[1:]:
                iv.close()
            plane_bytes = rb * spec.ncols * 4
            e2e = {"value": float(sum(s["n_valid"] for s in hst)) / (t1 - t0), "unit": UNIT,
                   "h2d_bytes_per_step": plane_bytes * (spec.n_planes + 1),
                   "d2h_bytes_per_step": plane_bytes * len(capi.SCALAR_PLANES) + rb * spec.ncols * 5,
                   "ms_per_step": 1e3 * (t1 - t0) / K,
                   "api": "phb_invert_host (pinned host planes in, host planes out)" if world == 1 else
                          f"phb_invert_host_multi from one process over {world} devices (pinned host planes in, host planes out)"}
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank == 0:
        alg_flops, ms_solve = float(agg[0]), float(agg[1]) / world
        achieved = alg_flops / (ms_solve * 1e-3) / 1e12 / world  # per-GPU TFLOP/s of the solve kernel
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
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
            "work_sharing": bool(stats[0].get("shared", False)),
            "e2e": e2e, "gpu_launches": 2 * K * world, "clocks": clocks,
        }
        # measured DRAM traffic of the kernel: bytes per pixel from the committed `ncu --set full` capture x pixels per launch
        for tname in ("r02_solve_kernel_traffic.json", "r01_solve_kernel_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                line["roofline"]["traffic"] = tj["dram_bytes_per_pixel"] * total_px / (K * world)
                line["roofline"]["traffic_source"] = f"profiles/{tname} (ncu dram__bytes_read+write per pixel) x pixels per launch"
                break
        # algorithmic HBM traffic (reported, not binding): planes + prior in, 9 planes + flags out
        bytes_px = spec.n_planes * 4 + 4 + 9 * 4 + 5
        line["roofline"]["hbm_gbs_algorithmic"] = bytes_px * rb * spec.ncols * K / (total_ms * 1e-3) / 1e9
        if not args.no_cpu_baseline and world == 1:
            p, pr = data[need[0]]
            line["cpu_baseline"] = cpu_baseline(spec.scaled(rb, spec.ncols), args, p.cpu().numpy(), pr.cpu().numpy(), args.cpu_seconds)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
