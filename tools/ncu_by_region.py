#!/usr/bin/env python3
"""Dynamic instruction / stall-sample share per code region (objective phases, centroid, ...)."""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "photic_b200", "csrc", "libphotic_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
kern = sys.argv[2] if len(sys.argv) > 2 else "solve_kernelILi3ELi32ELb0"
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
lines, cur, ins, labels = [], ("?", 0), [], {}
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*(\.L_x_\d+):", l)
    if m:
        labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur)); ins.append(m.group(2).strip())
# the objective-evaluation loop of each pixel class (round 2: two classes in one kernel): instruction index ranges
back = []
for i, t in enumerate(ins):
    m = re.search(r"BRA\s+.*?`\((\.L_x_\d+)\)", t)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        back.append((labels[m.group(1)], i))
terms = [(a, b) for (a, b) in back if sum(1 for x in ins[a:b + 1] if "LDS.128" in x) >= 16 and b - a < 600]
evals = []
for (a, b) in terms:
    enc = [(x, y) for (x, y) in back if x <= a and y >= b and (y - x) > (b - a) + 500 and (y - x) < 5000]
    if enc:
        evals.append(min(enc, key=lambda z: z[1] - z[0]))
addr_of = [a for a, _ in lines]
def class_of(addr):
    for k, (x, y) in enumerate(evals[:2]):
        if addr_of[x] <= addr <= addr_of[y]:
            return f"c{k}:"
    return "  :"
src = open(os.path.join(root, "photic_b200", "csrc", "invert_kernel.cuh")).read().splitlines()
# region markers: comment tags in the source
tags = [("prepass", "(scene,band) pre-pass"), ("terms", "forward model, one (region"), ("sqsum", "squared residuals added"),
        ("pen_depth", "depth continuity, samodel"), ("pen_bottom", "bottom continuity, samodel"), ("pen_K", "K penalties, samodel"),
        ("obj_tail", "if (final_pass) {"), ("nm_helpers", "Nelder-Mead helpers"), ("setup", "per-pixel set-up"),
        ("kernel_prologue", "Persistent solve kernel"), ("switch", "switch (phase)"), ("transitions", "resolve the transitions"),
        ("iter_begin(centroid)", "next == NX_ITER_BEGIN"), ("other_transitions", "next == NX_SIMPLEX"), ("outputs", "derived outputs, samodel")]
marks = []
for name, pat in tags:
    for i, l in enumerate(src, 1):
        if pat in l:
            marks.append((i, name)); break
marks.sort()
def region(c):
    f, ln = c
    if f != "invert_kernel.cuh":
        return f
    nm = "head"
    for a, n in marks:
        if a <= ln:
            nm = n
    return nm
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:solve_kernel"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = {n: i for i, n in enumerate(rows[h])}
data = [r for r in rows[h + 1:] if len(r) > 5]
base = int(data[0][0], 16)
byoff = dict(lines)
agg = defaultdict(lambda: defaultdict(float))
for r in data:
    c = byoff.get(int(r[0], 16) - base, ("?", 0))
    g = class_of(int(r[0], 16) - base) + region(c)
    for k in ("# Samples", "Instructions Executed", "stall_wait", "stall_long_sb", "stall_no_inst", "stall_short_sb"):
        agg[g][k] += float(r[cols[k]])
    agg[g]["sass"] += 1
ti = sum(a["Instructions Executed"] for a in agg.values()); ts = sum(a["# Samples"] for a in agg.values())
print(f"{'region (c0/c1: inside the evaluation loop of pixel class 0/1)':24s} {'inst%':>6} {'smp%':>6} {'wait%':>6} {'long%':>6} {'noin%':>6} {'shrt%':>6} {'sass':>6}")
for g, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    print(f"{g:24s} {100*a['Instructions Executed']/ti:6.2f} {100*a['# Samples']/ts:6.2f} {100*a['stall_wait']/ts:6.2f} "
          f"{100*a['stall_long_sb']/ts:6.2f} {100*a['stall_no_inst']/ts:6.2f} {100*a['stall_short_sb']/ts:6.2f} {int(a['sass']):6d}")
