#!/usr/bin/env python3
"""Static view of the forward-model term loop of solve_kernel<3,32,false>: size, opcode mix, spill traffic and the
longest run of FP64 instructions between other-pipe instructions.  usage: term_loop.py lib.so [-v]"""
import re, subprocess, os, sys, tempfile, collections
so = os.path.abspath(sys.argv[1])
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and "solve_kernelILi3ELi32ELb0" in l)
ins = []; labels = {}
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r"\s*(\.L_x_\d+):", l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append(m.group(2).strip())
print("kernel instructions", len(ins))
best = None
for i, s in enumerate(ins):
    m = re.search(r"BRA\s+.*?`\((\.L_x_\d+)\)", s)
    if not m or m.group(1) not in labels: continue
    t = labels[m.group(1)]
    if t >= i: continue
    body = ins[t:i + 1]
    if sum(1 for b in body if "LDS.128" in b) >= 16 and (best is None or len(body) < len(best)): best = body
c = collections.Counter()
for b in best:
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", b); c[mm.group(2).split(".")[0]] += 1
print("term loop size", len(best), dict(c.most_common(40)))
print("spill ops in loop:", sum(1 for b in best if re.search(r"\b(STL|LDL)\b", b)))
if "-v" in sys.argv:
    for b in best: print("   ", b)
