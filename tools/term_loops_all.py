#!/usr/bin/env python3
"""Every forward-model term loop of solve_kernel<3,32,false> (one per pixel class since round 2): size, opcode mix, spill
traffic inside it, and where the kernel's spill instructions sit.  usage: term_loops_all.py lib.so [kernel-substring]"""
import re, subprocess, os, sys, tempfile, collections
so = os.path.abspath(sys.argv[1])
kern = sys.argv[2] if len(sys.argv) > 2 else "solve_kernelILi3ELi32ELb0"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
ins = []; labels = {}
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r"\s*(\.L_x_\d+):", l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append(m.group(2).strip())
print("kernel instructions", len(ins))
loops = []
for i, s in enumerate(ins):
    m = re.search(r"BRA\s+.*?`\((\.L_x_\d+)\)", s)
    if not m or m.group(1) not in labels: continue
    t = labels[m.group(1)]
    if t >= i: continue
    body = ins[t:i + 1]
    if sum(1 for b in body if "LDS.128" in b) >= 16 and len(body) < 600: loops.append((t, i))
for (t, i) in loops:
    body = ins[t:i + 1]
    c = collections.Counter()
    for b in body:
        mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", b); c[mm.group(2).split(".")[0]] += 1
    fp = sum(c[k] for k in ("DADD", "DFMA", "DMUL", "DSETP"))
    print(f"term loop [{t},{i}] size {len(body)} fp64 {fp} spill ops {sum(1 for b in body if re.search(r'(STL|LDL)', b))}", dict(c.most_common(12)))
sp = [k for k, b in enumerate(ins) if re.search(r"\b(STL|LDL)", b)]
print("spill instructions:", len(sp), "at", sp[:80])
# outer evaluation loops: smallest loop containing each term loop entirely and > 2000 instructions
for (t, i) in loops:
    best = None
    for j, s in enumerate(ins):
        m = re.search(r"BRA\s+.*?`\((\.L_x_\d+)\)", s)
        if not m or m.group(1) not in labels: continue
        tt = labels[m.group(1)]
        if tt <= t and j >= i and (j - tt) > (i - t) + 500 and (best is None or j - tt < best[1] - best[0]): best = (tt, j)
    if best:
        n_sp = sum(1 for k in sp if best[0] <= k <= best[1])
        print(f"  enclosing loop of [{t},{i}]: [{best[0]},{best[1]}] size {best[1]-best[0]+1}, spill ops inside {n_sp}")
