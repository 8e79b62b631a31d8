"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden vectors of the reference. The bar is BIT equality of every retrieved parameter in double
precision, identical evaluation counts and identical convergence flags; that implies the stated
tolerance (|dH| <= 1e-3 m, relative <= 1e-4 on the other parameters for >= 99.9 % of pixels)."""
from dataclasses import replace

import numpy as np
import pytest

from conftest import MINED_FIXTURES, SCENE_FIXTURES, bits_equal, cfg_from_golden, desc_from_golden, load_golden

pytestmark = pytest.mark.gpu

H_TOL, REL_TOL = 1.0e-3, 1.0e-4  # north_star tolerance, for the float32 output planes


def _full_grid(R, C):
    ii, jj = np.meshgrid(np.arange(R), np.arange(C), indexing="ij")
    return ii.ravel(), jj.ravel()


def _check_planes_against_records(out, rec, pix, ncols, ns):
    """The float32 planes the shim hands back are the float casts of the records (samodel.c:1120-1160)."""
    i, j = pix // ncols, pix % ncols
    assert np.array_equal(out["depth"][i, j], -(rec[:, 0].astype(np.float32)))
    for col, name in ((1, "model_error"), (2, "bottom_albedo"), (3, "bottom_sand"), (4, "bottom_seagrass"),
                      (5, "bottom_coral"), (6, "K_min"), (7, "index_optical_depth"), (8, "bottom_type")):
        assert np.array_equal(out[name][i, j], rec[:, col].astype(np.float32)), name
    K = rec[:, 16:16 + ns * 4].reshape(-1, ns, 4).astype(np.float32)
    assert np.array_equal(out["K"][:, :, i, j].transpose(2, 0, 1), K)
    pgx = rec[:, 16 + ns * 4:16 + ns * 4 + 3 * ns].reshape(-1, ns, 3).astype(np.float32)
    for k, name in enumerate("PGX"):
        assert np.array_equal(out[name][:, i, j].T, pgx[:, :, k]), name


@pytest.mark.parametrize("name", SCENE_FIXTURES)
def test_golden_scenes_bit_exact(inverter, name):
    """Committed outputs of the UNMODIFIED reference (tests/golden/make_golden.py)."""
    g = load_golden(name)
    desc = desc_from_golden(g)
    _, R, C = g["planes"].shape
    prior = g["prior"] if bool(g["use_prior"]) else None
    out, st = inverter.invert_host(desc, g["planes"], prior, debug=True)
    ok = g["status"] == 1
    pix = np.nonzero(ok)[0]
    assert st["n_valid"] == ok.sum() and np.array_equal(out["pix"], pix)
    assert np.array_equal(out["rec_evals"], g["n_evals"][ok])
    assert np.array_equal(out["rec_converged"], g["converged"][ok])
    assert bits_equal(out["rec"], g["rec"][ok]).all()
    ns = len(g["theta_sun"])
    _check_planes_against_records(out, out["rec"], pix, C, ns)
    # identical convergence-flag map, defaults where nothing is inverted (samodel.c:819-829)
    assert np.array_equal(out["converged"].ravel(), g["converged"].astype(np.uint8))
    assert np.array_equal(out["n_evals"].ravel(), g["n_evals"])
    bad = ~ok.reshape(R, C)
    assert (out["bottom_sand"][bad] == -9999.0).all() and (out["bottom_type"][bad] == -9999.0).all()
    assert (out["depth"][bad] == 0.0).all() and (out["K_min"][bad] == 0.0).all()


@pytest.mark.parametrize("env", [{"PHB_WARPS_PER_CTA": "8"}, {"PHB_CTAS_PER_SM": "2"}, {"PHB_CTAS_PER_SM": "4", "PHB_ALIGN": "1"},
                                 {"PHB_ALIGN": "1"}, {"PHB_TMEM": "0"}, {"PHB_SIMPLEX_SMEM_BYTES": "0"}, {"PHB_CT_LAYOUT": "0"}])
def test_launch_geometry_does_not_change_a_bit(inverter, env, monkeypatch):
    """The tuning / profiling knobs of launch_solve (warps per CTA, CTAs per SM, aligned evaluations, tensor memory off,
    no simplex rows in shared memory: every split of the simplex over its three storage tiers; the run-time-layout
    instantiation instead of the compile-time one) only move work and data around: records, evaluation counts and flags
    stay the reference's."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name in ("scene_exmouth", "scene_noprior"):
        g = load_golden(name)
        prior = g["prior"] if bool(g["use_prior"]) else None
        out, st = inverter.invert_host(desc_from_golden(g), g["planes"], prior, debug=True)
        ok = g["status"] == 1
        assert st["n_valid"] == ok.sum()
        assert np.array_equal(out["rec_evals"], g["n_evals"][ok]) and np.array_equal(out["rec_converged"], g["converged"][ok])
        assert bits_equal(out["rec"], g["rec"][ok]).all(), (env, name)


@pytest.mark.parametrize("name", MINED_FIXTURES)
def test_mined_non_converged_and_restart_pixels_bit_exact(inverter, name):
    """Reference goldens of the rare pixels (tests/golden/make_golden.py make_mined / make_restarts): the device's
    `ifault = 2` exit (kcount exhausted -> converged 0 with more than 5000 evaluations, asa047.c:217, 411-453 ->
    samodel.c:2396-2402), the budget hit exactly with a passing factorial test, and nelmin restarts (numres >= 1,
    asa047.c:481-493) must equal the UNMODIFIED reference: full record, evaluation count, flag, restart count."""
    g = load_golden(name)
    desc = desc_from_golden(g)
    _, R, C = g["planes"].shape
    out, st = inverter.invert_host(desc, g["planes"], g["prior"], debug=True)
    centres = g["centre_i"].astype(np.int64) * C + g["centre_j"]
    sel = np.searchsorted(out["pix"], centres)
    assert np.array_equal(out["pix"][sel], centres)
    assert np.array_equal(out["rec_converged"][sel], g["converged"])
    assert np.array_equal(out["rec_evals"][sel], g["n_evals"])
    assert np.array_equal(out["rec_restarts"][sel], g["n_restarts"])
    assert bits_equal(out["rec"][sel], g["rec"]).all()
    # the flag / count planes the shim hands back carry the same values
    assert np.array_equal(out["converged"][g["centre_i"], g["centre_j"]], g["converged"].astype(np.uint8))
    assert np.array_equal(out["n_evals"][g["centre_i"], g["centre_j"]], g["n_evals"])
    if "restart" in name:
        assert (out["rec_restarts"][sel] >= 1).any()
    else:
        assert (out["rec_converged"][sel] == 0).sum() >= 10
    assert st["n_converged"] == int(out["rec_converged"].sum())


@pytest.mark.parametrize("cfg_name,R,C,over", [
    ("murion", 40, 31, {}), ("exmouth", 37, 29, {}), ("abudhabi", 24, 20, {}), ("qatar", 30, 30, {}),
    ("pilbara", 16, 33, {}), ("murion", 1, 23, {}), ("murion", 19, 1, {}), ("murion", 12, 12, {"n_spatial": 0}),
    ("exmouth", 12, 12, {"n_bottoms": 1}), ("murion", 10, 10, {"n_dates": 1}), ("murion", 10, 10, {"n_bottoms": 8}),
    ("murion", 8, 8, {"n_dates": 9}),    # 36 (scene,band) slots: the SBP = 128 instantiations of the kernel
    ("murion", 7, 7, {"n_dates": 16}),   # PHB_MAX_SCENES: n = 111 parameters, four coordinates per lane, T = 576
])
def test_seeded_scenes_vs_oracle(inverter, oracle_port, cfg_name, R, C, over):
    """Fresh seeded scenes (ragged shapes, single row / column, 1 / 9 / 16 dates, 1 and 8 substrates, NSPATIAL 0)."""
    from oracle.binding import SceneCfg
    from photic_b200 import capi, scene
    spec = replace(scene.CONFIGS[cfg_name].scaled(R, C), **over)
    planes, prior = scene.generate(spec)
    planes, prior = planes.numpy(), prior.numpy()
    out, st = inverter.invert_host(capi.desc_from_spec(spec), planes, prior, debug=True)
    ii, jj = _full_grid(R, C)
    ref = oracle_port.invert_pixels(SceneCfg.from_spec(spec), planes, scene.NODATA, prior, scene.NODATA, ii, jj)
    ok = ref["status"] == 1
    assert ok.sum() > 0 and st["n_valid"] == ok.sum()
    assert np.array_equal(out["pix"], np.nonzero(ok)[0])
    assert np.array_equal(out["rec_evals"], ref["n_evals"][ok]) and np.array_equal(out["rec_converged"], ref["converged"][ok])
    assert bits_equal(out["rec"], ref["rec"][ok]).all()
    assert st["n_evals"] >= ref["n_evals"].sum() and st["n_converged"] == ref["converged"].sum()
    # the stated tolerance on the float32 planes, spelled out (implied by bit equality above)
    d = -out["depth"].ravel()[ok]
    assert (np.abs(d - ref["rec"][ok, 0]) <= H_TOL).mean() >= 0.999
    k = out["K_min"].ravel()[ok]
    assert (np.abs(k - ref["rec"][ok, 6]) <= REL_TOL * np.abs(ref["rec"][ok, 6]) + 1e-12).mean() >= 0.999


def test_empty_and_all_land_scenes(inverter):
    from photic_b200 import capi, scene
    spec = scene.CONFIGS["murion"].scaled(9, 7)
    planes = np.full((spec.n_planes, 9, 7), scene.NODATA, dtype=np.float32)
    prior = np.full((9, 7), scene.NODATA, dtype=np.float32)
    out, st = inverter.invert_host(capi.desc_from_spec(spec), planes, prior)
    assert st["n_valid"] == 0 and (out["bottom_sand"] == -9999.0).all() and (out["depth"] == 0.0).all()
    planes[:] = -0.5  # negative reflectance everywhere: valid nowhere (samodel.c:939)
    out, st = inverter.invert_host(capi.desc_from_spec(spec), planes, prior)
    assert st["n_valid"] == 0


def test_row_band_shards_equal_whole_scene(inverter):
    """Size-independent property used at full scale: inverting row bands with their halo rows gives
    exactly the unsharded planes (edge clamping happens at the global edge only)."""
    from photic_b200 import capi, scene, sharded
    spec = scene.CONFIGS["abudhabi"].scaled(30, 22)
    planes, prior = scene.generate(spec)
    planes, prior = planes.numpy(), prior.numpy()
    whole, _ = inverter.invert_host(capi.desc_from_spec(spec), planes, prior)
    halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
    cost = scene.valid_mask(__import__("torch").from_numpy(planes)).sum(dim=1).numpy()
    for world in (2, 4):
        plan = sharded.plan_row_bands(cost, world)
        for r0, r1 in plan:
            if r1 <= r0:
                continue
            w0, w1, lb, le = sharded.window(r0, r1, halo, spec.nrows)
            part, _ = inverter.invert_host(capi.desc_from_spec(spec, nrows=w1 - w0), np.ascontiguousarray(planes[:, w0:w1]),
                                           np.ascontiguousarray(prior[w0:w1]), row_begin=lb, row_end=le)
            for name in ("depth", "K_min", "bottom_albedo", "index_optical_depth", "converged", "n_evals", "P", "K"):
                a = part[name][..., lb:le, :]
                b = whole[name][..., r0:r1, :]
                assert np.array_equal(a, b), (name, world, r0, r1)


def test_run_to_run_determinism_and_device_entry(inverter):
    """Queue order is nondeterministic (atomics); results must not be. Also exercises phb_invert_device."""
    import torch
    from photic_b200 import capi, scene
    from photic_b200.samodel import Inverter
    spec = scene.CONFIGS["qatar"].scaled(48, 40)
    planes, prior = scene.generate(spec, device="cuda")
    desc = capi.desc_from_spec(spec)
    outs = []
    for _ in range(2):
        o = Inverter.alloc_device_outputs(desc, "cuda")
        st = inverter.invert_device(desc, planes, prior, o)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items()})
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k
    host, st_h = inverter.invert_host(desc, planes.cpu().numpy(), prior.cpu().numpy())
    for k in outs[0]:
        assert np.array_equal(outs[0][k], host[k]), k
    assert st["n_valid"] == st_h["n_valid"] > 0 and st["alg_flops"] == st_h["alg_flops"] > 0


def test_full_size_exmouth_sampled_against_oracle(inverter, oracle_port):
    """BASELINE.json configs[1] at FULL size (3930x2858, 6 dates, ~6 M valid pixels) through the device entry:
    size-independent properties (every valid pixel inverted exactly once, defaults elsewhere, flags consistent)
    plus a random sample of pixels checked against the CPU oracle on the same full-size raster, bit for bit
    at float32 output precision (what samodel() stores) and in the evaluation counts / convergence flags."""
    import torch
    from oracle.binding import SceneCfg
    from photic_b200 import capi, scene
    from photic_b200.samodel import Inverter
    spec = scene.CONFIGS["exmouth"]
    planes, prior = scene.generate(spec, device="cuda")
    desc = capi.desc_from_spec(spec)
    o = Inverter.alloc_device_outputs(desc, "cuda", scene_planes=False)
    st = inverter.invert_device(desc, planes, prior, o)
    torch.cuda.synchronize()
    valid = scene.valid_mask(planes)
    n_valid = int(valid.sum())
    assert st["n_valid"] == n_valid > 5_000_000
    depth = o["depth"]
    assert bool((depth[valid] < 0).all()) and bool((depth[~valid] == 0).all())          # negated depth on valid pixels only
    assert bool((o["bottom_type"][~valid] == -9999.0).all()) and bool((o["n_evals"][~valid] == 0).all())
    assert bool((o["n_evals"][valid] > 100).all()) and int(o["converged"].sum()) == st["n_converged"]
    assert st["n_converged"] > 0.999 * n_valid
    idx = torch.nonzero(valid.reshape(-1)).reshape(-1)
    g = torch.Generator().manual_seed(7)
    pick = idx[torch.randperm(idx.numel(), generator=g)[:160].to(idx.device)].cpu().numpy()
    ii, jj = pick // spec.ncols, pick % spec.ncols
    ref = oracle_port.invert_pixels(SceneCfg.from_spec(spec), planes.cpu().numpy(), scene.NODATA, prior.cpu().numpy(),
                                    scene.NODATA, ii, jj, nthreads=0)
    got = {k: o[k].cpu().numpy()[ii, jj] for k in ("depth", "model_error", "K_min", "bottom_albedo", "index_optical_depth",
                                                   "bottom_sand", "n_evals", "converged")}
    rec = ref["rec"]
    assert np.array_equal(got["depth"].view(np.int32), (-rec[:, 0].astype(np.float32)).view(np.int32))
    for name, col in (("model_error", 1), ("bottom_albedo", 2), ("bottom_sand", 3), ("K_min", 6), ("index_optical_depth", 7)):
        assert np.array_equal(got[name].view(np.int32), rec[:, col].astype(np.float32).view(np.int32)), name
    assert np.array_equal(got["n_evals"], ref["n_evals"]) and np.array_equal(got["converged"], ref["converged"].astype(np.uint8))


@pytest.mark.parametrize("generic", [False, True])
def test_objective_known_answers_on_device(inverter, generic, monkeypatch):
    """samodel_error / samodel_Rrs on random parameter vectors: reference's own outputs (golden). Through the
    instantiation the solve kernel uses for the substrate count (compile-time classes for 3 and 1) and, generic, through
    the run-time-count instantiation for every count (PHB_ONE_CLASS=1)."""
    from photic_b200 import capi, scene
    if generic:
        monkeypatch.setenv("PHB_ONE_CLASS", "1")
    k = load_golden("kat_objective")
    for tag in "abcd":
        ns, nb, nr, origin = (int(v) for v in k[f"{tag}_meta"])
        spec = replace(scene.CONFIGS["murion"], n_dates=ns)
        got = inverter.kat_objective(capi.desc_from_spec(spec), nb, nr, origin, k[f"{tag}_meas"], k[f"{tag}_params"])
        assert bits_equal(got, k[f"{tag}_out"]).all(), tag


@pytest.mark.parametrize("generic", [False, True])
def test_objective_extreme_parameters_take_the_fallback_paths(inverter, generic, monkeypatch):
    """Parameter vectors outside the guaranteed range of the branch-free division / sqrt / exp (H = 0, 1e-300, 150 m,
    1e300; zero, tiny and huge IOPs and albedos; 0/0 mixing weights; inf and NaN coordinates), in all or only some
    regions so that fallback lanes sit next to fast-path lanes: still the reference's bits, NaNs where it has NaNs."""
    from photic_b200 import capi, scene
    if generic:
        monkeypatch.setenv("PHB_ONE_CLASS", "1")
    k = load_golden("kat_objective_extreme")
    for tag in "abc":
        ns, nb, nr, origin = (int(v) for v in k[f"{tag}_meta"])
        spec = replace(scene.CONFIGS["murion"], n_dates=ns)
        got = inverter.kat_objective(capi.desc_from_spec(spec), nb, nr, origin, k[f"{tag}_meas"], k[f"{tag}_params"])
        eq = bits_equal(got, k[f"{tag}_out"])
        assert eq.all(), (tag, np.argwhere(~eq)[:8], got[~eq][:8], k[f"{tag}_out"][~eq][:8])


def test_pinned_host_buffers_equal_pageable_ones(inverter):
    """Page-locked caller buffers are read and written by the copy engine directly, pageable ones go through the pinned
    staging ring (photic_b200.cu: band_upload / band_download): same results either way, row window included."""
    import torch
    from photic_b200 import capi, scene
    spec = scene.CONFIGS["exmouth"].scaled(41, 37)
    planes, prior = scene.generate(spec)
    desc = capi.desc_from_spec(spec)
    want, st0 = inverter.invert_host(desc, planes.numpy(), prior.numpy(), scene_planes=False)
    pp, pr = planes.clone().pin_memory(), prior.clone().pin_memory()
    buf = {n: torch.zeros((spec.nrows, spec.ncols), dtype=torch.float32).pin_memory().numpy() for n in capi.SCALAR_PLANES}
    buf["converged"] = torch.zeros((spec.nrows, spec.ncols), dtype=torch.uint8).pin_memory().numpy()
    buf["n_evals"] = torch.zeros((spec.nrows, spec.ncols), dtype=torch.int32).pin_memory().numpy()
    got, st1 = inverter.invert_host(desc, pp.numpy(), pr.numpy(), scene_planes=False, buffers=buf)
    assert st1["n_valid"] == st0["n_valid"] > 300 and st1["n_evals"] == st0["n_evals"]
    for name in list(capi.SCALAR_PLANES) + ["converged", "n_evals"]:
        assert np.array_equal(got[name].view(np.uint8), want[name].view(np.uint8)), name
    # a row window: only its rows are written
    for name in buf:
        buf[name][...] = 0
    got, st2 = inverter.invert_host(desc, pp.numpy(), pr.numpy(), row_begin=9, row_end=30, scene_planes=False, buffers=buf)
    for name in list(capi.SCALAR_PLANES) + ["converged", "n_evals"]:
        assert np.array_equal(got[name][9:30].view(np.uint8), want[name][9:30].view(np.uint8)), name
        assert not got[name][:9].any() and not got[name][30:].any(), name


def test_team_objective_equals_warp_objective(inverter):
    """Mapping study (aux_kernels.cuh): the objective evaluated by a team of 2 / 4 / 8 warps -- terms dealt over the
    team's lanes, ordered sum and penalties on one warp -- gives the bits of the one-warp objective, which are the
    reference's (golden), in both warp-to-scheduler placements and on every SM."""
    from photic_b200 import capi, scene
    k = load_golden("kat_objective")
    seen = 0
    for tag in "abcd":
        ns, nb, nr, origin = (int(v) for v in k[f"{tag}_meta"])
        if nb not in (1, 3) or nr > 16 or ns * 4 > 32:
            continue
        spec = replace(scene.CONFIGS["murion"], n_dates=ns)
        desc = capi.desc_from_spec(spec)
        for v in (0, 1, 7):
            want = k[f"{tag}_out"][v, 0]
            for tw in (1, 2, 4, 8):
                for same in (False, True):
                    got, rate, ms = inverter.eval_bench(desc, nb, nr, origin, k[f"{tag}_meas"], k[f"{tag}_params"][v], tw, same, reps=3)
                    assert np.float64(got).view(np.int64) == np.float64(want).view(np.int64), (tag, v, tw, same, got, want)
                    seen += 1
    assert seen >= 48


def test_device_math_equals_host_libm(inverter):
    """exp/log/pow on the device == the libm the reference links (same process, math module / numpy ufunc free)."""
    import math
    rng = np.random.default_rng(5)
    n = 100_000
    x = np.concatenate([rng.uniform(-760, 30, n), [0.0, -0.0, -745.2, 709.0, np.inf, -np.inf]])
    ref = np.array([math.exp(v) if v < 709.7 else np.inf for v in x])
    assert bits_equal(inverter.kat_math(0, x), ref).all()
    x = np.exp(rng.uniform(-40, 10, n))
    assert bits_equal(inverter.kat_math(1, x), np.array([math.log(v) for v in x])).all()
    x, y = rng.uniform(0.3, 1.5, n), rng.uniform(-3, 3, n)
    assert bits_equal(inverter.kat_math(2, x, y), np.array([math.pow(a, b) for a, b in zip(x, y)])).all()


def test_branch_free_div_sqrt_exp_equal_ieee_operators(inverter):
    """The hot loop uses nvcc's div/sqrt FAST PATHS without their range branch (one check per term instead).
    Inside the guarded range they must equal the ordinary operators bit for bit; exp_main must equal exp."""
    rng = np.random.default_rng(9)
    n = 4_000_000
    for lo, hi in ((-380.0, 380.0), (-30.0, 30.0), (-3.0, 3.0)):
        a = np.exp2(rng.uniform(lo, hi, n)) * rng.choice([-1.0, 1.0], n)
        b = np.exp2(rng.uniform(lo / 2, hi / 2, n)) * rng.choice([-1.0, 1.0], n)
        assert bits_equal(inverter.kat_math(3, a, b), inverter.kat_math(5, a, b)).all()
        x = np.abs(a)
        assert bits_equal(inverter.kat_math(4, x), inverter.kat_math(6, x)).all()
    u = rng.uniform(0.0, 1.0, n)   # the forward model's actual arguments
    for c in (2.4, 5.4):
        x = 1.0 + c * u
        assert bits_equal(inverter.kat_math(4, x), inverter.kat_math(6, x)).all()
    rho = rng.uniform(1e-6, 1.0, n)
    assert bits_equal(inverter.kat_math(3, rho, np.full(n, 3.141592653589793)), inverter.kat_math(5, rho, np.full(n, 3.141592653589793))).all()
    x = -np.exp(rng.uniform(np.log(2.0 ** -53), np.log(511.9), n))
    assert bits_equal(inverter.kat_math(7, x), inverter.kat_math(0, x)).all()
    # rho / pi with the reciprocal refined once per context (div_by_pi) == the IEEE quotient
    pi = np.full(n, 3.141592653589793)
    for lo, hi in ((-380.0, 380.0), (-30.0, 3.0)):
        a = np.exp2(rng.uniform(lo, hi, n)) * rng.choice([-1.0, 1.0], n)
        assert bits_equal(inverter.kat_math(8, a), inverter.kat_math(5, a, pi)).all()


def test_guarded_division_and_sqrt_equal_ieee_everywhere(inverter):
    """div_by / sqrt_guarded (fast path inside its range, the ordinary operator outside) == a / b and sqrt(x) for ANY
    operands: random bit patterns (NaN, inf, subnormals, zeros included) and the magnitudes the objective produces."""
    rng = np.random.default_rng(12)
    n = 2_000_000
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1.0, 120.0])
    a = np.concatenate([rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64), np.repeat(special, len(special)),
                        np.exp(rng.uniform(-30, 30, n))])
    b = np.concatenate([rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64), np.tile(special, len(special)),
                        rng.choice([9.0, 81.0, 45.0, 216.0, 120.0, 0.0123], n)])
    assert bits_equal(inverter.kat_math(13, a, b), inverter.kat_math(5, a, b)).all()
    assert bits_equal(inverter.kat_math(15, a, b), inverter.kat_math(5, a, b)).all()
    assert bits_equal(inverter.kat_math(14, a), inverter.kat_math(6, a)).all()


def test_range_predicates_of_the_hot_loop(inverter):
    """The guards of the branch-free term read the high word of a double as a float (two FSETP each);
    they must select exactly the sets the integer exponent tests define."""
    rng = np.random.default_rng(10)
    n = 2_000_000
    bits = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    edge = np.array([0x2800000000000000, 0x27ffffffffffffff, 0x5800000000000000, 0x57ffffffffffffff,
                     0x3c90000000000000, 0x3c8fffffffffffff, 0x4080000000000000, 0x407fffffffffffff,
                     0x3ff0000000000000, 0x3ff00000ffffffff, 0x3ff0000100000000, 0x8000000000000000, 0,
                     0x7ff0000000000000, 0x7ff8000000000000, 0xfff8000000000000, 0x0000000000000001], dtype=np.uint64)
    edge = np.concatenate([edge, edge | np.uint64(1 << 63)])
    x = np.concatenate([bits, edge]).view(np.float64)
    hi = (x.view(np.uint64) >> np.uint64(32)).astype(np.uint64)
    e = (hi >> np.uint64(20)) & np.uint64(0x7ff)
    assert np.array_equal(inverter.kat_math(9, x) == 1.0, (e >= 0x280) & (e < 0x580))
    assert np.array_equal(inverter.kat_math(10, x) == 1.0, (e >= 0x3c9) & (e < 0x408))
    assert np.array_equal(inverter.kat_math(11, x) == 1.0, hi <= 0x3ff00000)


@pytest.mark.parametrize("mode", [0, 1])
def test_depth_sigma_matches_reference(inverter, mode):
    """phb_depth_sigma_host (samodel.c:1376-1477 on the GPU: host-side libc draws, trial chains on the device)
    against the golden run of the reference's own functions with the same seed: every trial depth, the sigma
    table and the sigma plane, bit for bit."""
    g = load_golden("depth_sigma_murion")
    desc = desc_from_golden(g)
    sig, table, trials, st = inverter.depth_sigma_host(desc, g["planes"], g["prior"], -g["depth"], int(g["seed"]),
                                                       int(g["n_samples"]), mode, int(g["max_intervals"]))
    assert st["n_valid"] == (g[f"trials{mode}"] != 0).sum() > 50
    assert bits_equal(trials, g[f"trials{mode}"]).all()
    assert bits_equal(table, g[f"table{mode}"]).all()
    assert np.array_equal(sig.view(np.int32), g[f"sigma{mode}"].view(np.int32))


def test_lee_kd_secchi_match_reference(inverter):
    """MODEL Lee_Kd_LS8 / Lee_Secchi_LS8 (secchi.c) on the GPU == the reference's rasters (golden), bit for bit,
    through the host entry and the device entry."""
    import torch
    g = load_golden("lee_ls8")
    for mode, key in ((0, "kd"), (1, "zsd")):
        got = inverter.lee_ls8_host(mode, g["coastal"], g["blue"], g["green"], g["red"], g["spv"], float(g["theta_s"]))
        same = (got.view(np.int32) == g[key].view(np.int32)) | (np.isnan(got) & np.isnan(g[key]))
        assert same.all(), key
        dev = [torch.from_numpy(g[k]).cuda() for k in ("coastal", "blue", "green", "red")]
        got2 = inverter.lee_ls8_device(mode, *dev, g["spv"], float(g["theta_s"])).cpu().numpy()
        assert np.array_equal(got2.view(np.int32), got.view(np.int32))


def test_device_log10_equals_host_libm(inverter):
    import math
    rng = np.random.default_rng(6)
    x = np.concatenate([np.exp(rng.uniform(-40, 40, 200_000)), rng.uniform(0.9, 1.1, 50_000), [1.0, 10.0, 1e-310, 0.5]])
    assert bits_equal(inverter.kat_math(12, x), np.array([math.log10(v) for v in x])).all()


def test_refine_matches_oracle(inverter, oracle_port):
    """REFINE (model/refine.c): every flag combination against the CPU restatement, bit exact."""
    from photic_b200 import capi
    rng = np.random.default_rng(11)
    grid = (-rng.uniform(0.2, 35.0, (57, 43))).astype(np.float32)
    grid[rng.uniform(size=grid.shape) < 0.2] = -9999.0
    land = np.where(rng.uniform(size=grid.shape) < 0.3, -9999.0, 1.0).astype(np.float32)
    shallow = np.where(rng.uniform(size=grid.shape) < 0.2, -9999.0, 1.0).astype(np.float32)
    args = np.array([-30.0, -0.5, -40.0, 0.0, 1.3, 0.9, -0.25, -32.0, -1.0, 1.1, 0.95], dtype=np.float32)
    for flags in (0, 1, 2, 3, 4, 8, 16, 1 | 2 | 4 | 8 | 16, 2 | 16, 1 | 8):
        for masks in ((None, None), (land, shallow)):
            a2 = args.copy()
            if flags == 3:
                a2[4] = 1.0  # SHAPE 1.0: the linear rescale branch
            got = inverter.refine_host(grid, -9999.0, flags, a2, land=masks[0], shallow=masks[1])
            exp = oracle_port.refine(grid, -9999.0, masks[0], -9999.0, masks[1], -9999.0, flags, a2)
            assert np.array_equal(got.view(np.int32), exp.view(np.int32)), (flags, masks[0] is not None)


def test_refine_matches_reference_golden(inverter):
    """REFINE kernels against the outputs of the reference's own run_refine() (tests/golden/refine.npz, made by
    oracle/_ref with refine.c compiled in): 72 cases incl. one-mask-only (whole grid blanked), no-CLIP min/max, SHAPE 1."""
    from conftest import refine_cases
    g = load_golden("refine")
    n = 0
    for flags, ld, ldn, sh, shn, args, exp in refine_cases(g):
        got = inverter.refine_host(g["grid"], -9999.0, flags, args, land=ld, land_nodata=ldn, shallow=sh, shallow_nodata=shn)
        assert np.array_equal(got.view(np.int32), exp.view(np.int32)), (flags, ld is not None, sh is not None)
        n += 1
    assert n == 72


def test_python_samodel_surface(inverter):
    """The reference-named call (scene / geogrid / samodel) fills the caller's grids in place."""
    from photic_b200 import scene as sc
    from photic_b200.samodel import geogrid, samodel, scene
    spec = sc.CONFIGS["murion"].scaled(16, 12)
    planes, prior = sc.generate(spec)
    grids = [geogrid(planes[g].numpy().copy()) for g in range(spec.n_planes)] + [geogrid(prior.numpy().copy())]
    scenes = [scene(f"d{s}", [4 * s + b for b in range(4)], list(spec.wavelengths), spec.theta_view, spec.theta_sun(s),
                    spec.h_tide(s)) for s in range(spec.n_dates)]
    outs = [np.zeros((16, 12), dtype=np.float32) for _ in range(10)]
    for s in scenes:
        s.R_sigma = [spec.r_sigma] * 4
    st = samodel(scenes, grids, list(range(spec.n_dates)), spec.n_dates, True, grids[-1], 1, 2, 3, *outs, inverter=inverter,
                 sigma_seed=5)
    assert st["n_valid"] > 0 and (outs[0] <= 0).all() and (outs[0] < 0).sum() == st["n_valid"]
    assert st["sigma_trials"] > 0 and (outs[1] >= 0).all() and (outs[1][outs[0] == 0] == 0).all()
