#!/usr/bin/env python3
"""SASS instruction count of solve_kernel by source region (I-cache footprint tracking)."""
import re, subprocess, os, tempfile, collections, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "photic_b200", "csrc", "libphotic_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
src = open(os.path.join(root, "photic_b200", "csrc", "invert_kernel.cuh")).read().splitlines()
marks = []  # (line, name) from function definitions
for i, l in enumerate(src, 1):
    m = re.match(r"__(?:device|global)__.*?\b(\w+)\(", l)
    if m and not l.strip().endswith(";"):
        marks.append((i, m.group(1)))
def region(f, ln):
    if f != "invert_kernel.cuh":
        return f
    name = "?"
    for a, nm in marks:
        if a <= ln:
            name = nm
    return name
for sect in [l for l in dis if l.startswith(".text.")]:
    if "solve_kernel" not in sect and "first_m" not in sect:
        continue
    start = dis.index(sect)
    cnt = collections.Counter(); cur = ("?", 0); total = 0
    for l in dis[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
            cnt[region(*cur)] += 1; total += 1
    print(sect[:70], "total", total, "=", total * 16 // 1024, "KB")
    for k, v in cnt.most_common(12):
        print(f"   {v:6d} {k}")
