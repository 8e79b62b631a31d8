#!/usr/bin/env python3
"""Static SASS instruction count per source line range of invert_kernel.cuh for one solve_kernel instantiation.
usage: sass_by_line.py [lib.so] [kernel substring]"""
import re, subprocess, os, sys, tempfile, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(root, "photic_b200", "csrc", "libphotic_b200.so")
kern = sys.argv[2] if len(sys.argv) > 2 else "solve_kernelILi3ELi32ELb0"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
cnt = collections.Counter(); cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): cnt[cur] += 1
src = open(os.path.join(root, "photic_b200", "csrc", "invert_kernel.cuh")).read().splitlines()
lines = sorted((ln, c) for (f, ln), c in cnt.items() if f == "invert_kernel.cuh")
other = sum(c for (f, ln), c in cnt.items() if f != "invert_kernel.cuh")
print("total", sum(cnt.values()), "other files", other)
for ln, c in lines:
    print(f"{ln:5d} {c:4d}  {src[ln-1].strip()[:110]}")
