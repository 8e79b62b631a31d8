"""The C host layer (photic_b200/host/samodel_b200.c) that keeps the reference's samodel() symbol."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT

REF = "/root/reference/model"


class Geogrid(C.Structure):  # photic_abi.h / model/common.h:69-84
    _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("cellsize", C.c_float), ("wlon", C.c_float),
                ("slat", C.c_float), ("elon", C.c_float), ("nlat", C.c_float), ("nodata_value", C.c_float),
                ("lambda_", C.c_float), ("theta_v", C.c_float), ("theta_w", C.c_float),
                ("array", C.POINTER(C.POINTER(C.c_float)))]


class Scene(C.Structure):  # photic_abi.h / model/common.h:194-218
    _fields_ = [("scene_name", C.c_char * 2048), ("n_bands", C.c_int), ("nrows", C.c_int), ("ncols", C.c_int),
                ("band_indexes", C.c_int * 265), ("wavelengths", C.c_int * 265), ("theta_w", C.c_double),
                ("theta_v", C.c_double), ("H_tide", C.c_double), ("R_inf", C.c_double * 265),
                ("R_sigma", C.c_double * 265), ("K", C.c_double * 265), ("K_sigma", C.c_double * 265),
                ("ratio_min", C.c_double), ("ratio_max", C.c_double), ("slope_min", C.c_double),
                ("slope_max", C.c_double), ("pgx_present", C.c_int), ("ps_p", C.c_int), ("ps_g", C.c_int),
                ("ps_x", C.c_int)]


PROBE = r'''
#include <stdio.h>
#include <stddef.h>
%s
int main(void) {
  printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\n", sizeof(GEO), offsetof(GEO, array), offsetof(GEO, nodata_value),
         sizeof(SCN), offsetof(SCN, band_indexes), offsetof(SCN, wavelengths), offsetof(SCN, theta_w), offsetof(SCN, H_tide));
  return 0;
}
'''


def _probe(header_block, flags):
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "p.c"), os.path.join(td, "p")
        open(src, "w").write(PROBE % header_block)
        subprocess.run(["gcc", "-w", *flags, src, "-o", exe], check=True)
        return [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]


def test_abi_mirror_matches_ctypes():
    mine = _probe('#include "photic_abi.h"\n#define GEO photic_geogrid\n#define SCN photic_scene',
                  ["-I", os.path.join(ROOT, "photic_b200", "host")])
    assert mine[0] == C.sizeof(Geogrid) and mine[3] == C.sizeof(Scene)
    assert mine[1] == Geogrid.array.offset and mine[4] == Scene.band_indexes.offset and mine[7] == Scene.H_tide.offset


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference headers absent")
def test_abi_mirror_matches_reference_headers():
    mine = _probe('#include "photic_abi.h"\n#define GEO photic_geogrid\n#define SCN photic_scene',
                  ["-I", os.path.join(ROOT, "photic_b200", "host")])
    ref = _probe('#include "common.h"\n#define GEO geogrid\n#define SCN scene', ["-fcommon", "-I", REF])
    assert mine == ref


def test_shim_exports_reference_symbol(product_lib):
    from photic_b200 import build
    so = build.build_host_shim()
    lib = C.CDLL(so)
    assert hasattr(lib, "samodel")


def _build_mini_model(tmp_path, so, extra=()):
    exe = str(tmp_path / "mini_model")
    subprocess.run(["gcc", "-O1", "-Wall", *extra, "-I", os.path.join(ROOT, "photic_b200", "host"),
                    os.path.join(ROOT, "tests", "csrc", "mini_model.c"), "-o", exe, "-L", os.path.dirname(so),
                    "-lsamodel_b200", "-Wl,-rpath," + os.path.dirname(so)], check=True)
    return exe


def test_c_program_linking_the_drop_in_fails_loudly_without_a_device(product_lib, tmp_path):
    """A C caller built like the reference's REPL (tests/csrc/mini_model.c) linked against the drop-in: on a machine
    without a CUDA device samodel() must print the error and exit(1) -- the reference's own failure convention
    (common.h:62-67) -- and never compute anything on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the no-device path cannot be observed")
    from photic_b200 import build
    so = build.build_host_shim()
    exe = _build_mini_model(tmp_path, so)
    for env_extra in ({}, {"PHOTIC_B200_DEVICES": "all"}, {"PHOTIC_B200_DEVICES": "0,1"}):
        r = subprocess.run([exe], capture_output=True, text=True, env={**os.environ, **env_extra}, timeout=120)
        assert r.returncode == 1, (env_extra, r.returncode, r.stdout[-400:], r.stderr[-400:])
        assert "ERROR: photic_b200" in r.stdout and "no CPU fallback" in r.stdout
        assert "mini_model: depth" not in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources absent")
def test_shim_builds_and_links_inside_the_reference_tree(product_lib, tmp_path):
    """INTEGRATION.md section 1, executed: the shim compiled INSTEAD of samodel.c with -DPHOTIC_REFERENCE_TREE against the
    reference's own headers (its `scene`, `geogrid`, `bool`, MAX_STRING_LEN, trim()), linked with the reference's
    common.c and a caller that uses the reference's types, and libphotic_b200. netcdf.h / cpgplot.h are the two empty
    stand-in headers the oracle build uses (the libraries are not in this image). Without a CUDA device the program
    must stop the reference's way: message + exit(1)."""
    import torch
    from photic_b200 import build
    lib = build.build()
    exe = str(tmp_path / "mini_ref")
    inc = ["-I", os.path.join(ROOT, "oracle", "ref_shim"), "-I", REF, "-I", os.path.join(ROOT, "include")]
    cmd = ["gcc", "-O1", "-fcommon", "-w", "-DPHOTIC_REFERENCE_TREE", *inc,
           os.path.join(ROOT, "photic_b200", "host", "samodel_b200.c"), os.path.join(ROOT, "tests", "csrc", "mini_model.c"),
           os.path.join(REF, "common.c"), "-o", exe, "-L", os.path.dirname(lib), "-lphotic_b200",
           "-Wl,-rpath," + os.path.dirname(lib), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # the stricter compile of the shim alone must be warning-free too
    r = subprocess.run(["gcc", "-c", "-fcommon", "-Wall", "-DPHOTIC_REFERENCE_TREE", *inc, "-o", str(tmp_path / "shim.o"),
                        os.path.join(ROOT, "photic_b200", "host", "samodel_b200.c")], capture_output=True, text=True)
    own = [l for l in r.stderr.splitlines() if "samodel_b200.c" in l and "warning" in l]
    assert r.returncode == 0 and not own, r.stderr[-2000:]
    if not torch.cuda.is_available():
        run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert run.returncode == 1 and "ERROR: photic_b200" in run.stdout and "no CPU fallback" in run.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["", "0,0,0"])
def test_shim_samodel_equals_library(inverter, devices):
    """Call the C samodel() the way bam.c:3236 does (row-pointer grids, geogrid by value); once on one device and
    once through the in-process multi-device path (PHOTIC_B200_DEVICES: three row bands, here all on device 0)."""
    from photic_b200 import build, capi, scene
    lib = C.CDLL(build.build_host_shim())
    lib.samodel_b200_shutdown()  # contexts are cached between calls: reopen under this test's environment
    os.environ.pop("PHOTIC_B200_DEVICES", None)
    if devices:
        os.environ["PHOTIC_B200_DEVICES"] = devices
    spec = scene.CONFIGS["murion"].scaled(18, 14)
    planes, prior = scene.generate(spec)
    planes, prior = planes.numpy(), prior.numpy()
    R, Cc, ns = spec.nrows, spec.ncols, spec.n_dates
    keep = []

    def rows(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        ptrs = (C.POINTER(C.c_float) * R)(*[a[r].ctypes.data_as(C.POINTER(C.c_float)) for r in range(R)])
        keep.extend([a, ptrs])
        return ptrs, a

    grids = (Geogrid * (spec.n_planes + 1))()
    for g in range(spec.n_planes + 1):
        p, _ = rows(planes[g] if g < spec.n_planes else prior)
        grids[g].nrows, grids[g].ncols, grids[g].nodata_value, grids[g].cellsize = R, Cc, scene.NODATA, 30.0
        grids[g].array = C.cast(p, C.POINTER(C.POINTER(C.c_float)))
    scenes = (Scene * ns)()
    for s in range(ns):
        scenes[s].scene_name = f"date{s}".encode()
        scenes[s].n_bands, scenes[s].nrows, scenes[s].ncols = 4, R, Cc
        for b in range(4):
            scenes[s].band_indexes[b] = 4 * s + b
            scenes[s].wavelengths[b] = spec.wavelengths[b]
        scenes[s].theta_v, scenes[s].theta_w, scenes[s].H_tide = spec.theta_view, spec.theta_sun(s), spec.h_tide(s)
        for b in range(4):
            scenes[s].R_sigma[b] = spec.r_sigma
    idx = (C.c_int * ns)(*range(ns))
    os.environ["PHOTIC_B200_SIGMA_SEED"] = "77"  # the reference seeds the depth-error pass with time(NULL)
    os.environ["PHOTIC_B200_SIGMA_CHAIN"] = "interval"  # one chain per depth interval (the shim's default is the reference's single chain)
    outs = [rows(np.full((R, Cc), 7.0, dtype=np.float32)) for _ in range(10)]
    lib.samodel.restype = None
    lib.samodel.argtypes = [C.POINTER(Scene), C.POINTER(Geogrid), C.POINTER(C.c_int), C.c_int, C.c_int, Geogrid, C.c_int,
                            C.c_int, C.c_int] + [C.POINTER(C.POINTER(C.c_float))] * 10 + [C.c_float, C.c_int, C.c_int]
    try:
        lib.samodel(scenes, grids, idx, ns, 1, grids[spec.n_planes], 1, 2, 3, *[o[0] for o in outs], 8.0, 0, 1)
    finally:
        lib.samodel_b200_shutdown()
        os.environ.pop("PHOTIC_B200_DEVICES", None)
    exp, st = inverter.invert_host(capi.desc_from_spec(spec), planes, prior)
    sigma, table, _, st2 = inverter.depth_sigma_host(capi.desc_from_spec(spec), planes, prior, exp["depth"], 77, 128, 1)
    order = ["depth", None, "model_error", "bottom_albedo", "bottom_sand", "bottom_seagrass", "bottom_coral", "K_min",
             "bottom_type", "index_optical_depth"]
    for k, name in enumerate(order):
        got = outs[k][1]
        if name is None:               # depth_sigma: the depth-error pass with the same seed
            assert np.array_equal(got.view(np.int32), sigma.view(np.int32)) and st2["n_valid"] > 0
        else:
            assert np.array_equal(got.view(np.int32), exp[name].view(np.int32)), name
    assert st["n_valid"] > 50


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["", "0,0"])
def test_c_program_runs_samodel_on_the_device_and_writes_the_reference_files(inverter, tmp_path, devices):
    """tests/csrc/mini_model.c as a linked C PROGRAM (row-pointer grids, geogrid by value, its own write_nc): the ten
    grids must equal the library's results bit for bit and write_nc must be called for exactly the files of
    samodel.c:1513-1687 in the reference's order -- per scene K_coastal, K_blue, K_green, K_red, P, G, X (the _D file
    is written under DELTA only, which is 0: samodel.c:1586-1596), then modelled_H, H_sigma, error, albedo,
    bottom_sand, bottom_seagrass, bottom_coral, min_K, bottom_type, index_optical_depth -- with the trimmed scene
    names (samodel.c:1516) and the geometry of the first grid."""
    from photic_b200 import build, capi, scene
    so = build.build_host_shim()
    exe = _build_mini_model(tmp_path, so)
    spec = scene.CONFIGS["murion"].scaled(16, 13)
    planes, prior = scene.generate(spec)
    planes, prior = planes.numpy(), prior.numpy()
    R, Cc, ns = spec.nrows, spec.ncols, spec.n_dates
    fin, fout, flog = (str(tmp_path / n) for n in ("in.bin", "out.bin", "log.txt"))
    with open(fin, "wb") as f:
        f.write(np.array([R, Cc, ns], dtype=np.int32).tobytes())
        f.write(planes.astype(np.float32).tobytes())
        f.write(prior.astype(np.float32).tobytes())
    env = {**os.environ, "PHOTIC_B200_SIGMA_SEED": "77", "PHOTIC_B200_SIGMA_CHAIN": "interval"}
    env.pop("PHOTIC_B200_DEVICES", None)
    if devices:
        env["PHOTIC_B200_DEVICES"] = devices
    r = subprocess.run([exe, fin, fout, flog], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "scene 0 is now named 'date0'" in r.stdout          # trim() works in place, as in the reference
    got = np.fromfile(fout, dtype=np.float32).reshape(10, R, Cc)
    desc = capi.make_desc(spec.wavelengths, 0.0, [25.0 + 3.0 * s for s in range(ns)], [0.1 * s for s in range(ns)], R, Cc,
                          r_sigma=1.0e-4)
    exp, st = inverter.invert_host(desc, planes, prior)
    sigma, _, _, st2 = inverter.depth_sigma_host(desc, planes, prior, exp["depth"], 77, 128, 1)
    order = ["depth", None, "model_error", "bottom_albedo", "bottom_sand", "bottom_seagrass", "bottom_coral", "K_min",
             "bottom_type", "index_optical_depth"]
    for k, name in enumerate(order):
        want = sigma if name is None else exp[name]
        assert np.array_equal(got[k].view(np.int32), want.view(np.int32)), name or "depth_sigma"
    assert st["n_valid"] > 40 and st2["n_valid"] > 0
    lines = [l.split() for l in open(flog).read().splitlines()]
    names = [l[0] for l in lines]
    want_names = []
    for s in range(ns):
        want_names += [f"date{s}_K_{b}.nc" for b in ("coastal", "blue", "green", "red")] + [f"date{s}_{v}.nc" for v in "PGX"]
    want_names += [f"modelled_{n}.nc" for n in ("H", "H_sigma", "error", "albedo", "bottom_sand", "bottom_seagrass",
                                                "bottom_coral", "min_K", "bottom_type", "index_optical_depth")]
    assert names == want_names and len(names) == 7 * ns + 10
    for l in lines:
        assert int(l[1]) == Cc and int(l[2]) == R and float(l[3]) == -9999.0
        assert float(l[5]) == np.float32(100.0 + 30.0 * (Cc - 1)) and float(l[6]) == np.float32(-20.0 + 30.0 * (R - 1))
    sums = {l[0]: float(l[4]) for l in lines}
    seq = lambda a: float(np.cumsum(a.ravel().astype(np.float64))[-1])  # noqa: E731  (row-major running sum, as the C loop)
    assert sums["modelled_H.nc"] == seq(exp["depth"]) and sums["modelled_H_sigma.nc"] == seq(sigma)
    for s in range(ns):
        for b, bn in enumerate(("coastal", "blue", "green", "red")):
            assert sums[f"date{s}_K_{bn}.nc"] == seq(exp["K"][s, b])
        assert sums[f"date{s}_P.nc"] == seq(exp["P"][s]) and sums[f"date{s}_X.nc"] == seq(exp["X"][s])
