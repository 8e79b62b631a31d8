"""photic_b200/csrc/exact_math.cuh, host build, must equal the live libm (the one the reference links)
bit for bit: the device evaluates the same operation sequence (tests/test_gpu_parity.py checks that)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_build", "libexactmath_host.so")
SRC = os.path.join(ROOT, "tests", "csrc", "exact_math_host.cpp")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    deps = [SRC, os.path.join(ROOT, "photic_b200", "csrc", "exact_math.cuh"),
            os.path.join(ROOT, "photic_b200", "csrc", "libm_tables.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC, "-lm"], check=True)
    lib = C.CDLL(SO)
    for f in (lib.phm_check_exp, lib.phm_check_log, lib.phm_check_pow, lib.phm_check_log10):
        f.restype = C.c_longlong
    return lib


def _bad1(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    first = C.c_longlong(-1)
    return fn(x.ctypes.data_as(_dp), C.c_longlong(x.size), C.byref(first)), (x[first.value] if first.value >= 0 else None)


def _bad2(lib, x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    first = C.c_longlong(-1)
    n = lib.phm_check_pow(x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), C.c_longlong(x.size), C.byref(first))
    return n, ((x[first.value], y[first.value]) if first.value >= 0 else None)


SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308, 1e-310,
                    1.7976931348623157e308, 0.5, 2.0, 3.0, -2.0, -3.0, -0.5, 709.78, 709.79, -708.4, -745.1, -745.2,
                    -746.0, 1e-20, -1e-20, 512.0, -512.0, 1024.0, -1024.0, 0.9375, 1.064697265625, 0.93749999999999989])


def test_libm_uses_fma_variant():
    """The disassembled routines are glibc's FMA ifunc variants; they are selected when the CPU has FMA+AVX2."""
    flags = open("/proc/cpuinfo").read()
    assert " fma " in flags and " avx2 " in flags


def test_exp(host):
    rng = np.random.default_rng(1)
    n = 2_000_000
    for x in (SPECIAL, rng.uniform(-50, 5, n), rng.uniform(-760, 720, n), -np.exp(rng.uniform(-45, 7, n)),
              rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64)):
        assert _bad1(host.phm_check_exp, x) == (0, None)


def test_log(host):
    rng = np.random.default_rng(2)
    n = 2_000_000
    for x in (SPECIAL, rng.uniform(1e-6, 2, n), rng.uniform(0.9, 1.1, n), np.exp(rng.uniform(-700, 700, n)),
              rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64)):
        assert _bad1(host.phm_check_log, x) == (0, None)


def test_log10(host):
    """glibc's log10 (fdlibm wrapper around log; used by the Lee Kd / Secchi model, secchi.c:155)."""
    rng = np.random.default_rng(4)
    n = 2_000_000
    for x in (SPECIAL, rng.uniform(1e-6, 20, n), rng.uniform(0.9, 1.1, n), np.exp(rng.uniform(-700, 700, n)),
              rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64)):
        assert _bad1(host.phm_check_log10, x) == (0, None)


def test_pow(host):
    rng = np.random.default_rng(3)
    n = 2_000_000
    sx, sy = np.meshgrid(SPECIAL, SPECIAL)
    cases = [(sx.ravel(), sy.ravel()), (rng.uniform(0.5, 1.2, n), rng.uniform(0, 2.5, n)),  # (440/lambda)^Y
             (rng.uniform(0.2, 5, n), np.full(n, -1.7)),                                      # P start
             (np.exp(rng.uniform(-50, 50, n)), rng.uniform(-20, 20, n)), (rng.uniform(-5, 5, n), np.round(rng.uniform(-6, 6, n))),
             (rng.uniform(-5, 5, n), rng.uniform(-6, 6, n)),
             (rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64), rng.integers(0, 2**64, n, dtype=np.uint64).view(np.float64)),
             (np.exp(rng.uniform(-1, 1, n)), np.exp(rng.uniform(-160, 50, n)) * rng.choice([-1, 1], n)),
             (rng.uniform(0, 1, n) * 1e-308, rng.uniform(-3, 3, n))]
    for x, y in cases:
        assert _bad2(host, x, y) == (0, None)
