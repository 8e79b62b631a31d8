#!/usr/bin/env python3
"""SASS instructions of each objective-evaluation loop of a solve kernel (one per pixel class), by source function and
by coarse source block.  usage: eval_loop_regions.py lib.so [kernel-substring] [src-root]"""
import re, subprocess, os, sys, tempfile, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.abspath(sys.argv[1])
kern = sys.argv[2] if len(sys.argv) > 2 else "solve_kernelILi3ELi32ELb0"
srcroot = sys.argv[3] if len(sys.argv) > 3 else root
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
src = open(os.path.join(srcroot, "photic_b200", "csrc", "invert_kernel.cuh")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"__(?:device|global)__.*?\b(\w+)\(", l)
    if m and not l.strip().endswith(";"): marks.append((i, m.group(1)))
def region(f, ln):
    if f != "invert_kernel.cuh": return f
    name = "?"
    for a, nm in marks:
        if a <= ln: name = nm
    return name
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
ins = []; labels = {}; cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*(\.L_x_\d+):", l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((m.group(2).strip(), cur))
back = []
for i, (s, _) in enumerate(ins):
    m = re.search(r"BRA\s+.*?`\((\.L_x_\d+)\)", s)
    if m and m.group(1) in labels and labels[m.group(1)] < i: back.append((labels[m.group(1)], i))
terms = [(t, i) for (t, i) in back if sum(1 for b, _ in ins[t:i + 1] if "LDS.128" in b) >= 16 and i - t < 600]
for (t, i) in terms:
    enc = [(a, b) for (a, b) in back if a <= t and b >= i and (b - a) > (i - t) + 500]
    if not enc: continue
    a, b = min(enc, key=lambda x: x[1] - x[0])
    if b - a > 5000: continue
    c = collections.Counter(); cl = collections.Counter()
    for s, (f, ln) in ins[a:b + 1]:
        c[region(f, ln)] += 1
    print(f"evaluation loop [{a},{b}] size {b - a + 1}; term loop {i - t + 1}")
    for k, v in c.most_common(14): print(f"   {v:6d} {k}")
