#!/usr/bin/env python3
"""Prints the handful of ncu metrics this project tracks from a .ncu-rep (first kernel in the report)."""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg.per_second",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second"]
d = dict(zip(hdr, zip(vals, units)))
for k in want:
    if k in d:
        print(f"{k:80s} {d[k][0]:>18s} {d[k][1]}")
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(d[k][0])
        if v > 0.05:
            print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
