"""The product's solve_kernel SOURCE run on the CPU (tests/emu: g++ build of photic_b200/csrc/invert_kernel.cuh, one
warp = 32 fibers) against the CPU oracle: every retrieved parameter, evaluation count and convergence flag must be bit
identical. Covers what can be wrong in the kernel's LOGIC without a GPU -- optimiser state machine, the simplex tiers
(global slab with centroid checkpoints / shared memory), ordered sums, penalties, neighbourhood gather at raster
edges, derived outputs and stores, the compile-time variants -- on every CPU run of the suite. What only a device can
show (PTX fast paths, tensor memory, nvcc's code) stays with the `-m gpu` tests."""
import os
import sys
from dataclasses import replace

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from runner import Emulator  # noqa: E402

from conftest import bits_equal  # noqa: E402


def _case(cfg_name, R, C, over, n_pix, stride_seed=0):
    from oracle.binding import SceneCfg
    from photic_b200 import capi, scene
    spec = replace(scene.CONFIGS[cfg_name].scaled(R, C), **over)
    planes, prior = scene.generate(spec)
    valid = scene.valid_mask(planes).numpy()
    ii, jj = np.nonzero(valid)
    step = max(1, len(ii) // n_pix)
    sel = np.arange(stride_seed % step, len(ii), step)[:n_pix]
    return spec, planes.numpy(), prior.numpy(), ii[sel], jj[sel], capi, scene, SceneCfg


def _check(emu, oracle_port, cfg_name, R, C, over, n_pix, smem, use_prior=True):
    spec, pl, pr, pi, pj, capi, scene, SceneCfg = _case(cfg_name, R, C, over, n_pix)
    desc = capi.desc_from_spec(spec, prior_present=use_prior)
    got = emu.invert_pixels(desc, pl, pr if use_prior else None, pi, pj, simplex_smem_bytes=smem)
    ref = oracle_port.invert_pixels(SceneCfg.from_spec(spec), pl, scene.NODATA, pr if use_prior else None, scene.NODATA, pi, pj)
    assert len(pi) >= min(2, n_pix)
    assert np.array_equal(got["n_evals"], ref["n_evals"]), (got["n_evals"], ref["n_evals"])
    assert np.array_equal(got["converged"], ref["converged"])
    eq = bits_equal(got["rec"], ref["rec"])
    assert eq.all(), np.argwhere(~eq)[:6]
    # the float planes are what samodel() leaves: depth negated, Rrs error in model_error (samodel.c:1120-1160, 1486)
    rec = ref["rec"]
    assert np.array_equal(got["planes"][0][pi, pj].view(np.int32), (-rec[:, 0].astype(np.float32)).view(np.int32))
    assert np.array_equal(got["planes"][1][pi, pj].view(np.int32), rec[:, 1].astype(np.float32).view(np.int32))
    assert np.array_equal(got["n_evals_plane"][pi, pj], ref["n_evals"])
    assert int(got["counters"][3]) == len(pi) and int(got["counters"][2]) == int(ref["converged"].sum())
    return got


@pytest.fixture(scope="module")
def emu(product_lib):
    return Emulator()


@pytest.mark.parametrize("cfg_name,R,C,over,n_pix,smem", [
    ("murion", 10, 8, {}, 10, 4096),                       # 4 dates, edges and corners of the raster (Nr = 4, 6, 9)
    ("exmouth", 12, 10, {}, 6, 0),                         # 6 dates, the whole simplex in the global slab (checkpoints)
    ("exmouth", 12, 10, {}, 4, 1 << 20),                   # the whole simplex in shared memory
    ("abudhabi", 9, 9, {}, 3, 20000),                      # 8 dates: 32 (scene, band) slots exactly
    ("qatar", 12, 12, {}, 4, 8192),                        # noisy mixed deep / shallow pixels
    ("murion", 9, 8, {"n_spatial": 1}, 4, 2048),           # one region
    ("murion", 8, 8, {"n_spatial": 3, "n_dates": 2}, 2, 30000),   # 25 regions
    ("murion", 9, 8, {"n_smoothing_radius": 2, "n_bottoms": 2}, 3, 4096),  # box smoothing, run-time substrate count
    ("murion", 8, 8, {"n_dates": 10}, 1, 16384),           # 40 (scene, band) slots: the 128-slot table stride
    ("murion", 8, 8, {"n_dates": 16}, 1, 30000),           # the maximum number of dates: 111 parameters, 4 per lane
    ("murion", 9, 9, {"n_spatial": 3, "n_dates": 8}, 1, 60000),   # 25 regions x 8 dates: 199 parameters, 800 terms
    ("murion", 8, 8, {"n_bottoms": 8, "n_dates": 3}, 1, 30000),   # every substrate of the table
    ("murion", 8, 8, {"n_bottoms": 1}, 2, 3000),           # sand only everywhere
    ("murion", 8, 8, {"n_dates": 1}, 3, 3000),             # a single date
])
def test_emulated_kernel_equals_oracle(emu, oracle_port, cfg_name, R, C, over, n_pix, smem):
    _check(emu, oracle_port, cfg_name, R, C, over, n_pix, smem)


def test_emulated_kernel_without_depth_prior(emu, oracle_port):
    """No DEPTHS grid: up to eight H starts per pixel with the early exit of samodel.c:2404."""
    _check(emu, oracle_port, "murion", 8, 8, {}, 2, 4096, use_prior=False)


@pytest.mark.parametrize("define", ["", "PHB_PIPELINE_TERMS=1", "PHB_COLD_OUT=1", "generic", "PHB_UNIFIED_PENALTY=0"])
def test_emulated_objective_known_answers(product_lib, define, monkeypatch):
    """samodel_error of the kernel source on the reference's own known answers (tests/golden/kat_objective*.npz),
    including the extreme parameter vectors that send lanes through the out-of-range fallbacks. Each substrate count
    runs through the instantiation the solve kernel uses for it (compile-time classes for 3 and 1, where the sand-only
    class takes the depth and substrate penalties in one pass); "generic": the run-time-count code for every count."""
    from conftest import load_golden
    from photic_b200 import capi, scene
    if define == "generic":
        monkeypatch.setenv("PHB_ONE_CLASS", "1")
        define = ""
    e = Emulator((define,), tag=define.split("=")[0].lower()) if define else Emulator()
    for gname, tags in (("kat_objective", "abcd"), ("kat_objective_extreme", "abc")):
        k = load_golden(gname)
        for tag in tags:
            ns, nb, nr, origin = (int(v) for v in k[f"{tag}_meta"])
            spec = replace(scene.CONFIGS["murion"], n_dates=ns)
            got = e.kat_objective(capi.desc_from_spec(spec), nb, nr, origin, k[f"{tag}_meas"], k[f"{tag}_params"])
            eq = bits_equal(got, k[f"{tag}_out"])
            assert eq.all(), (gname, tag, np.argwhere(~eq)[:6])


@pytest.mark.parametrize("define", ["PHB_PIPELINE_TERMS=1", "PHB_COLD_OUT=1"])
def test_emulated_variants_equal_oracle(product_lib, oracle_port, define):
    """The off-by-default experiment switches of the kernel are the same arithmetic in another order."""
    e = Emulator((define,), tag=define.split("=")[0].lower())
    _check(e, oracle_port, "exmouth", 12, 10, {}, 3, 6000)


@pytest.mark.parametrize("name", ["scene_nspatial1", "scene_nsmooth2_nb2"])
def test_emulated_device_pipeline_equals_reference_golden(emu, name):
    """classify_kernel -> concat_queue_kernel -> solve_kernel on a whole (small) raster, against the committed outputs
    of the UNMODIFIED reference: records, evaluation counts, convergence-flag map, the float planes samodel() leaves
    (depth negated, K / P / G / X per scene) and the defaults where nothing is inverted (samodel.c:819-829)."""
    from conftest import desc_from_golden, load_golden
    g = load_golden(name)
    desc = desc_from_golden(g)
    _, R, C = g["planes"].shape
    out = emu.invert_raster(desc, g["planes"], g["prior"] if bool(g["use_prior"]) else None, simplex_smem_bytes=3000)
    ok = g["status"] == 1
    pix = np.nonzero(ok)[0]
    assert out["n_valid"] == ok.sum() and np.array_equal(out["pix"], pix)
    assert np.array_equal(out["rec_evals"], g["n_evals"][ok]) and np.array_equal(out["rec_converged"], g["converged"][ok])
    assert bits_equal(out["rec"], g["rec"][ok]).all()
    rec, ns = out["rec"], len(g["theta_sun"])
    i, j = pix // C, pix % C
    assert np.array_equal(out["depth"][i, j], -(rec[:, 0].astype(np.float32)))
    for col, plane in ((1, "model_error"), (2, "bottom_albedo"), (3, "bottom_sand"), (4, "bottom_seagrass"),
                       (5, "bottom_coral"), (6, "K_min"), (7, "index_optical_depth"), (8, "bottom_type")):
        assert np.array_equal(out[plane][i, j], rec[:, col].astype(np.float32)), plane
    K = rec[:, 16:16 + ns * 4].reshape(-1, ns, 4).astype(np.float32)
    assert np.array_equal(out["K"][:, :, i, j].transpose(2, 0, 1), K)
    pgx = rec[:, 16 + ns * 4:16 + ns * 4 + 3 * ns].reshape(-1, ns, 3).astype(np.float32)
    for k, plane in enumerate("PGX"):
        assert np.array_equal(out[plane][:, i, j].T, pgx[:, :, k]), plane
    assert np.array_equal(out["converged"].ravel(), g["converged"].astype(np.uint8))
    assert np.array_equal(out["n_evals"].ravel(), g["n_evals"])
    bad = ~ok.reshape(R, C)
    assert (out["bottom_sand"][bad] == -9999.0).all() and (out["bottom_type"][bad] == -9999.0).all()
    assert (out["depth"][bad] == 0.0).all() and np.signbit(out["depth"][bad]).all() and (out["K_min"][bad] == 0.0).all()
    assert (out["K"][:, :, bad] == 0.0).all() and (out["P"][:, bad] == 0.0).all()
    # the work queue: every shallow-water pixel (all substrates) before every sand-only one
    q = out["queue_order"]
    deep = np.abs(np.where(g["prior"].ravel()[q] > -1.0, -1.0, g["prior"].ravel()[q])) > 8.0 if bool(g["use_prior"]) else np.zeros(len(q), bool)
    assert out["n_shallow"] == int((~deep).sum()) and not deep[:out["n_shallow"]].any() and deep[out["n_shallow"]:].all()


def test_pixels_whose_objective_is_never_a_number(emu, oracle_port):
    """An all-zero spectrum in the neighbourhood makes Rrs440/Rrs490 = 0/0 and with it every objective value NaN: no
    start ever beats `lowest`, and the reference hands back uninitialised memory with n_iterations = 0
    (samodel.c:2385-2413: `params` is only written when ynewlo < lowest). Here that case is defined: zero evaluations,
    not converged, depth 0 -- and the oracle agrees on the counts. Huge and tiny reflectances stay ordinary pixels."""
    from oracle.binding import SceneCfg
    from photic_b200 import capi, scene
    spec = scene.CONFIGS["murion"].scaled(9, 9)
    planes, prior = scene.generate(spec)
    pl, pr = planes.numpy().copy(), prior.numpy().copy()
    pl[:] = np.where(pl < 0, np.float32(0.004), pl)  # no land
    pl[:, 2, 2] = 0.0
    pl[:, 6, 6] = 1e30
    pl[:, 1, 5] = 1e-30
    pi, pj = np.array([2, 3, 6, 1]), np.array([2, 3, 6, 5])
    got = emu.invert_pixels(capi.desc_from_spec(spec), pl, pr, pi, pj)
    ref = oracle_port.invert_pixels(SceneCfg.from_spec(spec), pl, scene.NODATA, pr, scene.NODATA, pi, pj)
    assert np.array_equal(got["n_evals"], ref["n_evals"]) and np.array_equal(got["converged"], ref["converged"])
    assert got["n_evals"][:2].tolist() == [0, 0] and got["converged"][:2].tolist() == [0, 0]
    assert (got["planes"][0][pi[:2], pj[:2]] == 0.0).all() and (got["n_evals"][2:] > 100).all()
    assert bits_equal(got["rec"][2:], ref["rec"][2:]).all()


@pytest.mark.parametrize("name,take", [("mined_exmouth", (0, 20, 45)), ("mined_restart_exmouth", (0, 2)),
                                       ("mined_restart_qatar", (0,))])
def test_emulated_kernel_on_mined_reference_pixels(emu, name, take):
    """The kernel source on the reference's goldens of the rare exits (tests/golden/make_golden.py make_mined /
    make_restarts): kcount exhausted (`ifault = 2`, converged 0, asa047.c:217, 411-453), the budget hit exactly with a
    passing factorial test, and a restart after a failed one (numres = 1, asa047.c:481-493)."""
    from conftest import desc_from_golden, load_golden
    g = load_golden(name)
    desc = desc_from_golden(g)
    k = np.array(take)
    got = emu.invert_pixels(desc, g["planes"], g["prior"], g["centre_i"][k], g["centre_j"][k], simplex_smem_bytes=8192)
    assert np.array_equal(got["n_evals"], g["n_evals"][k]) and np.array_equal(got["converged"], g["converged"][k])
    assert np.array_equal(got["n_restarts"], g["n_restarts"][k])
    assert bits_equal(got["rec"], g["rec"][k]).all()
    if name == "mined_exmouth":
        assert got["converged"].tolist() == [0, 0, 1] and (got["n_evals"] > 5000).all()
    else:
        assert (got["n_restarts"] == 1).all()


def test_emulated_kernel_tests_every_grid_against_its_own_nodata(emu, oracle_port):
    """Per-grid nodata values (phb_scene_desc.nodata_band; samodel.c:683, 941, 2999-3003): the nodata cells of three grids
    re-coded to 0, -1 and 12345 and declared per grid give the records of the uniformly coded scene."""
    from oracle.binding import SceneCfg
    from photic_b200 import capi, scene
    spec, pl, pr, pi, pj, capi, scene, SceneCfg = _case("murion", 10, 8, {}, 6)
    ref = oracle_port.invert_pixels(SceneCfg.from_spec(spec), pl, scene.NODATA, pr, scene.NODATA, pi, pj)
    pl2 = pl.copy()
    nd = [[float(scene.NODATA)] * 4 for _ in range(spec.n_dates)]
    for (s, b), v in {(0, 1): 0.0, (1, 3): -1.0, (3, 0): 12345.0}.items():
        g = 4 * s + b
        assert (pl2[g] == scene.NODATA).any() and not (pl2[g] == v).any()
        pl2[g][pl2[g] == scene.NODATA] = v
        nd[s][b] = v
    desc = capi.make_desc(spec.wavelengths, spec.theta_view, [spec.theta_sun(s) for s in range(spec.n_dates)],
                          [spec.h_tide(s) for s in range(spec.n_dates)], spec.nrows, spec.ncols, nodata_band=nd,
                          r_sigma=spec.r_sigma)
    got = emu.invert_pixels(desc, pl2, pr, pi, pj)
    assert np.array_equal(got["n_evals"], ref["n_evals"]) and bits_equal(got["rec"], ref["rec"]).all()
    assert (ref["rec"][:, 13] < 9).any()  # some of the pixels do lose neighbours to nodata


def test_emulated_generic_kernel_on_the_default_substrate_count(oracle_port, monkeypatch):
    """PHB_ONE_CLASS=1: NBOTTOMS = 3 through the run-time-substrate-count kernel and ONE queue (the path every other
    NBOTTOMS takes) instead of the two compile-time pixel classes; same bits."""
    monkeypatch.setenv("PHB_ONE_CLASS", "1")
    _check(Emulator(), oracle_port, "exmouth", 12, 10, {}, 4, 6000)


def test_ragged_band_lists_reach_the_model_constants(product_lib):
    """Scenes with different band counts (mixed sensors, samodel.c:403-409) through make_desc -> phb_band_tables."""
    from photic_b200 import capi
    d = capi.make_desc([[443, 482, 561, 655], [443, 561, 655], [482, 561]], 0.0, [25.0, 28.0, 31.0], [0.0, 0.1, 0.2], 4, 5,
                       r_sigma=[[1e-4] * 4, [2e-4] * 3, [3e-4] * 2])
    assert [d.n_bands[s] for s in range(3)] == [4, 3, 2] and d.r_sigma[2][1] == 3e-4 and d.wavelengths[1][2] == 655
    tab = np.zeros((3, capi.MAX_BANDS, 4 + capi.MAX_BOTTOMS))
    aux = np.zeros(1 + 2 * 3)
    capi.check(product_lib.phb_band_tables(d, tab.ctypes.data_as(capi._dp), aux.ctypes.data_as(capi._dp)))
    assert np.array_equal(tab[0, 0], tab[1, 0]) and np.array_equal(tab[0, 2], tab[1, 1]) and np.array_equal(tab[0, 1], tab[2, 0])
    assert (tab[1, 3] == 0).all() and (tab[2, 2:] == 0).all() and (tab[2, :2, :4] > 0).all()
