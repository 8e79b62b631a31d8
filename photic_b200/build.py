"""Builds the product library in-tree: photic_b200/csrc/libphotic_b200.so (sm_100a only).

    python -m photic_b200.build            # or __graft_entry__.build()

-fmad=false is part of the correctness contract (bit parity with the reference's non-FMA x86-64
build; the FMAs that ARE wanted -- inside the glibc exp/log/pow port -- are explicit __fma_rn).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libphotic_b200.so")
HOST_SHIM = os.path.join(HERE, "host", "libsamodel_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_variant(tag: str, defines: list[str]) -> str:
    """Experimental build with extra -D flags -> csrc/libphotic_b200_<tag>.so (select with PHB_LIB=<path>)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(CSRC, f"libphotic_b200_{tag}.so")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, os.path.join(CSRC, "photic_b200.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    if r.returncode:
        print(log, file=sys.stderr)
        raise RuntimeError("nvcc failed")
    for line in log.splitlines():
        if "solve_kernel" in line or "Used" in line and "128" in line or "spill" in line and "solve" in line:
            pass
    import re
    m = re.search(r"solve_kernel.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", log, re.S)
    if m:
        print(f"{tag}: stack {m.group(1)} spill {m.group(2)} regs {m.group(3)}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(HERE, "..", "include", f) for f in ("photic_b200.h", "photic_spectra.h")]
    if force or _stale(LIB, srcs):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "photic_b200.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(CSRC, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if verbose or r.returncode:
            print(log, file=sys.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed; see photic_b200/csrc/build.log")
    return LIB


def build_host_shim(force: bool = False) -> str:
    """The C host layer that keeps the reference's samodel() symbol (photic_b200/host/samodel_b200.c)."""
    src = os.path.join(HERE, "host", "samodel_b200.c")
    deps = [src, os.path.join(HERE, "host", "photic_abi.h"), os.path.join(HERE, "..", "include", "photic_b200.h")]
    build()
    if force or _stale(HOST_SHIM, deps + [LIB]):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(HERE, "..", "include"), "-I",
               os.path.join(HERE, "host"), "-o", HOST_SHIM, src, "-L", CSRC, "-lphotic_b200", "-Wl,-rpath," + CSRC]
        subprocess.run(cmd, check=True)
    return HOST_SHIM


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host_shim(force="--force" in sys.argv))
