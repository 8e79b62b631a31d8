#!/usr/bin/env python3
"""Aggregates an ncu SASS-level source page by CUDA source line.

usage: tools/ncu_by_line.py <report.ncu-rep> <kernel regex> [top N]
Needs the .so the report was taken from (photic_b200/csrc/libphotic_b200.so, built with -lineinfo).
ncu's CSV export of the CUDA view carries no metrics, so the SASS rows (in program order) are aligned
with `nvdisasm -g` of the same cubin, which annotates every instruction with file:line.
"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "photic_b200", "csrc", "libphotic_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel's .text section, in order
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kre, l))
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), m.group(2).strip(), cur))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = {n: i for i, n in enumerate(rows[hdr])}
data = [r for r in rows[hdr + 1:] if len(r) > 5]
base = int(data[0][0], 16)
byoff = {o: (t, c) for o, t, c in lines}
agg = defaultdict(lambda: defaultdict(float))
keys = ["# Samples", "Instructions Executed", "stall_wait", "stall_long_sb", "stall_no_inst", "stall_short_sb", "stall_math",
        "stall_branch_resolving", "L1 Wavefronts Shared", "L2 Theoretical Sectors Global", "L2 Theoretical Sectors Local"]
miss = 0
for r in data:
    off = int(r[0], 16) - base
    t, c = byoff.get(off, (None, ("?", 0)))
    if t is None:
        miss += 1
    for k in keys:
        try:
            agg[c][k] += float(r[cols[k]])
        except Exception:
            pass
    agg[c]["n_sass"] += 1
tot = {k: sum(a[k] for a in agg.values()) for k in keys}
print(f"kernel rows {len(data)}, disasm instr {len(lines)}, unmatched {miss}")
print("totals:", {k: f"{v:.3g}" for k, v in tot.items()})
src_cache = {}
def src(c):
    f, ln = c
    for d in ("photic_b200/csrc",):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < ln <= len(src_cache[p]):
                return src_cache[p][ln - 1].strip()[:90]
    return ""
print(f"{'inst%':>6} {'smp%':>6} {'wait':>6} {'longsb':>6} {'noinst':>6} {'sass':>5}  line")
for c, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print(f"{100*a['Instructions Executed']/tot['Instructions Executed']:6.2f} {100*a['# Samples']/tot['# Samples']:6.2f} "
          f"{100*a['stall_wait']/max(1,tot['# Samples']):6.2f} {100*a['stall_long_sb']/max(1,tot['# Samples']):6.2f} "
          f"{100*a['stall_no_inst']/max(1,tot['# Samples']):6.2f} {int(a['n_sass']):5d}  {c[0]}:{c[1]}  {src(c)}")
