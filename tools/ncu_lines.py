#!/usr/bin/env python3
"""Top source lines of solve_kernel by stall samples / instructions, straight from an .ncu-rep captured with
--import-source on (no need for the matching build).  usage: ncu_lines.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:solve_kernel"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = []; cur_file = "?"; cols = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": cols = {n: i for i, n in enumerate(r)}; continue
    if cols is None or len(r) < 10 or r[0] == "": continue
    try:
        ln = int(r[0]); smp = float(r[cols["# Samples"]]); ins = float(r[cols["Instructions Executed"]])
    except ValueError:
        continue
    g = lambda k: float(r[cols[k]]) if r[cols[k]] not in ("-", "") else 0.0
    out.append((cur_file, ln, r[1].strip(), smp, ins, g("stall_wait"), g("stall_long_sb"), g("stall_short_sb"), g("stall_no_inst"),
                g("stall_math"), g("stall_branch_resolving")))
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts:.0f}  instructions {ti:.3e}")
print(f"{'file:line':28s} {'smp%':>6} {'inst%':>6} {'wait':>5} {'long':>5} {'shrt':>5} {'noin':>5} {'math':>5} {'brch':>5}  source")
for o in sorted(out, key=lambda o: -o[3])[:N]:
    print(f"{o[0][:20]+':'+str(o[1]):28s} {100*o[3]/ts:6.2f} {100*o[4]/ti:6.2f} " + " ".join(f"{100*v/ts:5.2f}" for v in o[5:]) + "  " + o[2][:90])
