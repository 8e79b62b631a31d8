import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle_port():
    """The plain-C restatement (oracle/photic_oracle.c), compiled on demand."""
    from oracle import binding
    if not os.path.exists(binding.PORT_SO):
        binding.build(ref=False)
    return binding.Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref():
    """The unmodified reference (oracle/_ref), only where it has been built."""
    from oracle import binding
    if not os.path.exists(binding.REF_SO):
        if os.path.isdir("/root/reference/model"):
            binding.build(ref=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return binding.Oracle("reference")


@pytest.fixture(scope="session")
def product_lib():
    """libphotic_b200.so; built with nvcc when missing (cross-compiles without a GPU)."""
    from photic_b200 import build, capi
    if not os.path.exists(capi.LIB_PATH):
        build.build()
    return capi.lib()


@pytest.fixture(scope="session")
def inverter(product_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from photic_b200.samodel import Inverter
    return Inverter(0)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def cfg_from_golden(g):
    from oracle.binding import SceneCfg
    ns = len(g["theta_sun"])
    return SceneCfg(g["wavelengths"], float(g["theta_view"]), g["theta_sun"], g["h_tide"], None, int(g["n_smooth"]),
                    int(g["n_spatial"]), int(g["n_bottoms"]))


def desc_from_golden(g):
    from photic_b200 import capi
    _, nrows, ncols = g["planes"].shape
    return capi.make_desc(g["wavelengths"], float(g["theta_view"]), g["theta_sun"], g["h_tide"], nrows, ncols,
                          nodata=float(g["nodata"]), prior_present=bool(g["use_prior"]), prior_nodata=float(g["nodata"]),
                          n_smooth=int(g["n_smooth"]), n_spatial=int(g["n_spatial"]), n_bottoms=int(g["n_bottoms"]))


def bits_equal(a, b):
    """Bit equality of float64 arrays, any-NaN == any-NaN."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return (a.view(np.int64) == b.view(np.int64)) | (np.isnan(a) & np.isnan(b))


SCENE_FIXTURES = ["scene_murion", "scene_exmouth", "scene_qatar", "scene_noprior", "scene_nspatial1",
                  "scene_nsmooth2_nb2", "scene_nspatial3"]

# Reference goldens for the rare pixels (tests/golden/make_golden.py: make_mined, make_restarts): nelmin's kcount
# exhaustion (converged = 0), evaluation counts at the kcount edge, restarts after a failed factorial test. Patch k of
# the raster = columns [3k, 3k+3); the reference inverted the patch centres (centre_i, centre_j).
MINED_FIXTURES = ["mined_exmouth", "mined_qatar", "mined_abudhabi", "mined_pilbara", "mined_restart_exmouth",
                  "mined_restart_qatar"]


def refine_cases(g):
    """(flags, land, land_nodata, shallow, shallow_nodata, args, expected) of tests/golden/refine.npz: the outputs of the
    reference's own run_refine() (refine.c:12-302) for every flag set x {no mask, both, LAND only, SHALLOW only} x SHAPE."""
    grid, land, shallow, args = g["grid"], g["land"], g["shallow"], g["args"]
    for flags in (int(f) for f in g["flags"]):
        for mk, (ld, sh) in enumerate(((None, None), (land, shallow), (land, None), (None, shallow))):
            for shape1 in (0, 1):
                key = f"out_{flags}_{mk}_{shape1}"
                if key not in g.files:
                    continue
                a2 = args.copy()
                if shape1:
                    a2[4] = 1.0
                yield flags, ld, -9999.0, sh, -7777.0, a2, g[key]
