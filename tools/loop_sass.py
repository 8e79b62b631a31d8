#!/usr/bin/env python3
"""Static size of the hot loops of solve_kernel<3>: finds backward branches, prints per-loop instruction
class counts for loops whose body contains the given opcode pattern.  usage: loop_sass.py [lib.so] [pattern] [mincount]"""
import re, subprocess, os, sys, tempfile, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "photic_b200", "csrc", "libphotic_b200.so")
pat = sys.argv[2] if len(sys.argv) > 2 else "LDS.128"
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
kern = sys.argv[4] if len(sys.argv) > 4 else "solve_kernelILi3E"
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
def cls(op):
    if op in ("DADD", "DFMA", "DMUL", "DSETP"): return "fp64"
    if op in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDTM", "STTM", "LDC", "LDCU"): return "mem"
    if op in ("BRA", "BSSY", "BSYNC", "WARPSYNC", "CALL", "RET", "NOP", "BREAK"): return "ctrl"
    if op in ("UMOV", "MOV") : return "mov"
    return "int"
print("kernel instructions:", len(ins))
for i, s in enumerate(ins):
    m = re.search(r"BRA\s+.*?`\((\.L_x_\d+)\)", s)
    if not m or m.group(1) not in labels: continue
    t = labels[m.group(1)]
    if t >= i: continue
    body = ins[t:i + 1]
    if sum(1 for b in body if pat in b) < minc: continue
    c = collections.Counter(); ops = collections.Counter()
    for b in body:
        mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", b)
        op = mm.group(2); base = op.split(".")[0]
        if op.startswith("IMAD.MOV"): c["mov"] += 1
        else: c[cls(base)] += 1
        ops[op.split(".")[0] if not op.startswith("IMAD.MOV") else "IMAD.MOV"] += 1
    print(f"loop [{t},{i}] size {len(body)}: " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
    print("   " + " ".join(f"{k}:{v}" for k, v in ops.most_common(30)))
    if "-v" in sys.argv:
        for b in body: print("      ", b)
