"""int16 scale/offset packing of a grid: compress_2d / decompress_2d of model/nc.c:247-320.
Goldens (tests/golden/nc_pack.npz) come from the UNMODIFIED nc.c (oracle/ref_nc.c includes it where it lies;
tests/golden/make_golden.py nc_pack). The restatement is checked on the CPU, the device kernels on the GPU; both bit
for bit, including the order-dependent range of the reference (the maximum is only tested when a value did not lower
the running minimum and starts at FLT_MIN) and its wrapped shorts."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_golden

CASES = ["random", "decreasing", "increasing", "first_is_max", "flat", "all_spval", "nan_inf", "positive", "many_chunks",
         "decreasing_many_chunks"]


def _same_f32(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float32).view(np.int32), np.asarray(b, dtype=np.float32).view(np.int32))


@pytest.mark.parametrize("name", CASES)
def test_restatement_equals_reference_golden(oracle_port, name):
    g = load_golden("nc_pack")
    packed, off, sc, miss = oracle_port.nc_pack(g[f"{name}_grid"], -9999.0)
    assert np.array_equal(packed, g[f"{name}_packed"]) and miss == int(g[f"{name}_missing"][0]) == -32768
    assert _same_f32([off, sc], g[f"{name}_meta"])
    assert _same_f32(oracle_port.nc_unpack(packed, off, sc, miss, -9999.0), g[f"{name}_unpacked"])


def test_goldens_hold_the_reference_quirks():
    g = load_golden("nc_pack")
    # every value of a decreasing grid lowers the minimum: the maximum stays FLT_MIN, so the shorts wrap
    d = g["decreasing_grid"]
    assert g["decreasing_meta"][0] == d.min() and g["decreasing_meta"][1] == np.float32((np.float32(1.17549435e-38) - d.min()) / np.float32(32767.0))
    assert g["decreasing_packed"].min() < 0
    # the first value is never a candidate for the maximum
    f = g["first_is_max_grid"]
    assert f[0, 0] == 100.0 and np.array_equal(g["first_is_max_meta"], g["random_meta"])
    # in the long decreasing run only the one value that did not lower the minimum sets the maximum
    assert g["decreasing_many_chunks_meta"][1] == np.float32((np.float32(19.5) - np.float32(-20.0)) / np.float32(32767.0))


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference sources absent")
@pytest.mark.parametrize("name", ["random", "many_chunks"])
def test_golden_is_reproducible_from_the_reference(oracle_ref, name):
    g = load_golden("nc_pack")
    packed, off, sc, miss = oracle_ref.nc_pack(g[f"{name}_grid"], -9999.0)
    assert np.array_equal(packed, g[f"{name}_packed"]) and _same_f32([off, sc], g[f"{name}_meta"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_packing_equals_reference_golden(inverter, name):
    from photic_b200 import capi
    g = load_golden("nc_pack")
    grid = np.ascontiguousarray(g[f"{name}_grid"], dtype=np.float32)
    packed = np.zeros(grid.shape, dtype=np.int16)
    off, sc, miss = C.c_float(0), C.c_float(0), C.c_int16(0)
    capi.check(inverter.lib.phb_nc_pack_host(inverter.ctx, grid.ctypes.data_as(capi._fp), grid.shape[0], grid.shape[1],
                                             C.c_double(-9999.0), packed.ctypes.data_as(C.c_void_p), C.byref(off), C.byref(sc),
                                             C.byref(miss)))
    assert _same_f32([off.value, sc.value], g[f"{name}_meta"]) and miss.value == -32768
    assert np.array_equal(packed, g[f"{name}_packed"])
    back = np.zeros(grid.shape, dtype=np.float32)
    capi.check(inverter.lib.phb_nc_unpack_host(inverter.ctx, packed.ctypes.data_as(C.c_void_p), grid.shape[0], grid.shape[1],
                                               off, sc, miss, C.c_double(-9999.0), back.ctypes.data_as(capi._fp)))
    assert _same_f32(back, g[f"{name}_unpacked"])


@pytest.mark.gpu
def test_device_packing_of_a_large_plane_equals_the_restatement(inverter, oracle_port):
    """A depth plane of 1500 x 2858 cells (1047 chunks of the device passes) against the CPU restatement."""
    import torch
    from photic_b200 import capi
    rng = np.random.default_rng(3)
    grid = (-rng.gamma(2.0, 6.0, (1500, 2858))).astype(np.float32)
    grid[rng.uniform(size=grid.shape) < 0.45] = -9999.0
    want = oracle_port.nc_pack(grid, -9999.0)
    d = torch.from_numpy(grid).cuda()
    out = torch.zeros(grid.shape, dtype=torch.int16, device="cuda")
    off, sc, miss = C.c_float(0), C.c_float(0), C.c_int16(0)
    capi.check(inverter.lib.phb_nc_pack_device(inverter.ctx, C.c_void_p(d.data_ptr()), d.numel(), C.c_double(-9999.0),
                                               C.c_void_p(out.data_ptr()), C.byref(off), C.byref(sc), C.byref(miss),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert _same_f32([off.value, sc.value], [want[1], want[2]]) and np.array_equal(out.cpu().numpy(), want[0])
