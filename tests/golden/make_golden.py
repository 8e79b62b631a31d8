#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libphotic_ref.so).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The reference ships no tests or fixtures for this path, so these known
answers -- produced by the reference's own compiled functions on seeded inputs -- are the pin for
oracle/photic_oracle.c (tests/test_oracle_vs_golden.py) and, through it, for the CUDA path.
"""
import os
import sys
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import Oracle, SceneCfg, build  # noqa: E402
from photic_b200 import scene  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (file, base config, rows, cols, overrides, use prior)
SCENES = [
    ("scene_murion", "murion", 24, 20, {}, True),
    ("scene_exmouth", "exmouth", 20, 16, {}, True),
    ("scene_qatar", "qatar", 16, 16, {}, True),
    ("scene_noprior", "murion", 8, 8, {}, False),
    ("scene_nspatial1", "murion", 12, 10, {"n_spatial": 1}, True),
    ("scene_nsmooth2_nb2", "murion", 12, 10, {"n_smoothing_radius": 2, "n_bottoms": 2}, True),
    ("scene_nspatial3", "murion", 8, 8, {"n_spatial": 3, "n_dates": 2}, True),
]


def main():
    build()
    ref = Oracle("reference")
    rng = np.random.default_rng(20261017)
    for fname, base, R, Cc, over, use_prior in SCENES:
        spec = replace(scene.CONFIGS[base].scaled(R, Cc), **over)
        planes, prior = scene.generate(spec)
        planes, prior = planes.numpy(), prior.numpy()
        if fname == "scene_murion":  # edge cases: negative neighbour, prior nodata on a valid pixel, very shallow prior
            v = np.argwhere(scene.valid_mask(scene.generate(spec)[0]).numpy())
            planes[1, v[5][0], v[5][1]] = -1.0e-4
            prior[v[9][0], v[9][1]] = scene.NODATA
            prior[v[11][0], v[11][1]] = -0.4
        cfg = SceneCfg.from_spec(spec)
        ii, jj = np.meshgrid(np.arange(R), np.arange(Cc), indexing="ij")
        ii, jj = ii.ravel(), jj.ravel()
        out = ref.invert_pixels(cfg, planes, scene.NODATA, prior if use_prior else None, scene.NODATA, ii, jj, nthreads=8)
        np.savez_compressed(
            os.path.join(HERE, fname + ".npz"), planes=planes, prior=prior, use_prior=use_prior,
            wavelengths=np.array(spec.wavelengths), theta_view=spec.theta_view,
            theta_sun=np.array([spec.theta_sun(s) for s in range(spec.n_dates)]),
            h_tide=np.array([spec.h_tide(s) for s in range(spec.n_dates)]), n_smooth=spec.n_smoothing_radius,
            n_spatial=spec.n_spatial, n_bottoms=spec.n_bottoms, nodata=scene.NODATA, rec=out["rec"],
            status=out["status"], converged=out["converged"], n_evals=out["n_evals"])
        print(fname, "valid", int(out["status"].sum()), "of", R * Cc, "conv", int(out["converged"].sum()),
              "evals", out["n_evals"][out["status"] == 1].mean())

    # known answers of the objective / forward model on random parameter vectors
    kat = {}
    for tag, ns, nb, nr in (("a", 4, 3, 9), ("b", 6, 1, 9), ("c", 8, 3, 6), ("d", 2, 2, 4)):
        spec = replace(scene.CONFIGS["murion"], n_dates=ns)
        cfg = SceneCfg.from_spec(spec)
        meas = rng.uniform(0.002, 0.02, (nr, ns, 4))
        n = nr + 2 * nr * nb + 3 * ns
        params = np.concatenate([rng.uniform(0.3, 45, (64, nr)), rng.uniform(5, 60, (64, nr * nb)),
                                 rng.uniform(0.1, 2, (64, nr * nb)), rng.uniform(0.5, 12, (64, 3 * ns))], axis=1)
        params[::7] *= -1.0  # sign is ignored through fabs
        params[3, :nr] = 0.7  # shallow: K penalties
        params[4, :nr] = 1.7
        params[5, :nr] = 2.7
        params[6, :nr] = 3.7
        params[8, :nr] = 4.7
        assert params.shape[1] == n
        origin = nr // 2
        out, rrs, K = ref.error_kat(cfg, nb, nr, origin, meas, params)
        for k, v in (("meas", meas), ("params", params), ("out", out), ("rrs", rrs), ("K", K),
                     ("meta", np.array([ns, nb, nr, origin]))):
            kat[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kat_objective.npz"), **kat)

    # nelmin known answers on analytic functions (incl. ties, restarts, kcount exhaustion)
    nm = {}
    idx = 0
    for fn_id in (0, 1, 2):
        for n in (2, 5, 12, 30):
            for kcount in (5000, 300):
                start = rng.uniform(-2, 3, n)
                step = rng.uniform(0.3, 2.0, n)
                xmin, y, ic, nr_, ifl = ref.nelmin_kat(fn_id, start, step, 1e-2, 100 if n > 5 else 10, kcount)
                nm[f"{idx}_in"] = np.concatenate([[fn_id, n, kcount, 100 if n > 5 else 10], start, step])
                nm[f"{idx}_out"] = np.concatenate([[y, ic, nr_, ifl], xmin])
                idx += 1
    np.savez_compressed(os.path.join(HERE, "kat_nelmin.npz"), **nm)

    # interp_1d / approx_equal / band tables
    X = np.array([443.0, 482.0, 561.0, 655.0])
    Y = rng.uniform(0.001, 0.02, 4)
    xs = np.array([440.0, 443.0, 443.001, 482.0, 490.0, 550.0, 561.0, 640.0, 655.0, 654.999, 750.0, 400.0])
    vals = np.array([ref.interp_1d(X, Y, x) for x in xs])
    ae = np.array([[a, b, e, ref.approx_equal(a, b, e)] for a, b, e in
                   [(-9999.0, -9999.0, 1e-6), (-9999.0, -9998.99, 1e-6), (-9999.0, -9998.0, 1e-4), (0.0, 0.0, 1e-6),
                    (1e-7, 0.0, 1e-6), (0.01, 0.0100000001, 1e-6), (443.0, 443.004, 1e-5), (443.0, 443.005, 1e-5)]])
    cfg = SceneCfg.from_spec(scene.CONFIGS["abudhabi"])
    t0, t1 = ref.tables(cfg)
    np.savez_compressed(os.path.join(HERE, "kat_misc.npz"), X=X, Y=Y, xs=xs, vals=vals, approx=ae, tables=t0, aux=t1)
    print("done")


def make_depth_sigma():
    """Depth-error phase (samodel.c:1376-1477) through the reference's own functions with srand(seed):
    one hot-start chain (the reference) and the chain restarted at every interval."""
    build()
    ref = Oracle("reference")
    spec = scene.CONFIGS["murion"].scaled(40, 32)
    planes, prior = scene.generate(spec)
    planes, prior = planes.numpy(), prior.numpy()
    cfg = SceneCfg.from_spec(spec)
    import torch
    ii, jj = np.nonzero(scene.valid_mask(torch.from_numpy(planes)).numpy())
    out = ref.invert_pixels(cfg, planes, scene.NODATA, prior, scene.NODATA, ii, jj, nthreads=8)
    depth = np.zeros((spec.nrows, spec.ncols), dtype=np.float32)
    depth[ii, jj] = out["rec"][:, 0].astype(np.float32)  # (float) md->depth, samodel.c:1120
    prior[ii[3], jj[3]] = scene.NODATA                    # a valid pixel without prior: its trials score 0
    seed, n_samples, max_int = 20261017, 16, 40
    res = {}
    for mode in (0, 1):
        table, trials, sig = ref.depth_sigma(cfg, planes, scene.NODATA, prior, scene.NODATA, depth, seed, n_samples, mode,
                                             max_int)
        res[f"table{mode}"], res[f"trials{mode}"], res[f"sigma{mode}"] = table, trials, sig
        print("depth_sigma mode", mode, "intervals", len(table), "trials run", int((trials != 0).sum()), "sigma>0 cells",
              int((sig > 0).sum()))
    np.savez_compressed(
        os.path.join(HERE, "depth_sigma_murion.npz"), planes=planes, prior=prior, depth=depth, seed=seed,
        n_samples=n_samples, max_intervals=max_int, wavelengths=np.array(spec.wavelengths), theta_view=spec.theta_view,
        theta_sun=np.array([spec.theta_sun(s) for s in range(spec.n_dates)]),
        h_tide=np.array([spec.h_tide(s) for s in range(spec.n_dates)]), n_smooth=spec.n_smoothing_radius,
        n_spatial=spec.n_spatial, n_bottoms=spec.n_bottoms, nodata=scene.NODATA, use_prior=True, **res)


def make_lee_ls8():
    """MODEL Lee_Kd_LS8 / Lee_Secchi_LS8 through the reference's Kd_LS8 / secchi_disk_depth (secchi.c)."""
    build()
    ref = Oracle("reference")
    spec = scene.CONFIGS["murion"].scaled(96, 80)
    planes, _ = scene.generate(spec)
    c, b, g, r = [planes[k].numpy().copy() for k in range(4)]
    rng = np.random.default_rng(20261018)
    sl = slice(0, 24)  # free-ranging positive reflectances, incl. values the model turns into negative / huge Kd
    for a, hi in ((c, 0.05), (b, 0.05), (g, 0.05), (r, 0.03)):
        a[sl] = np.exp(rng.uniform(np.log(1e-6), np.log(hi), a[sl].shape)).astype(np.float32)
    g[30, 5] = -9999.0  # nodata in one band only
    spv = np.array([-9999.0, -9999.0, -9999.0, -9999.0], dtype=np.float32)
    theta = 28.5
    kd = ref.lee_ls8(0, c, b, g, r, spv, theta)
    zsd = ref.lee_ls8(1, c, b, g, r, spv, theta)
    print("lee_ls8: cells", kd.size, "valid", int((kd != -9999.0).sum()), "nan", int(np.isnan(kd).sum()), int(np.isnan(zsd).sum()))
    np.savez_compressed(os.path.join(HERE, "lee_ls8.npz"), coastal=c, blue=b, green=g, red=r, spv=spv, theta_s=theta,
                        kd=kd, zsd=zsd)


def extreme_params(rng, ns, nb, nr):
    """Parameter vectors that leave the guaranteed range of the device's branch-free fast paths (exponent of exp
    beyond +-512 or below 2^-54, huge / tiny / zero / non-finite operands) in some or all regions, so that the
    out-of-line fallbacks run next to fast-path lanes. Layout: SURVEY.md appendix A.1."""
    n = nr + 2 * nr * nb + 3 * ns
    base = np.concatenate([rng.uniform(0.5, 30, nr), rng.uniform(5, 60, nr * nb), rng.uniform(0.1, 2, nr * nb),
                           rng.uniform(0.5, 12, 3 * ns)])
    off = nr + 2 * nr * nb
    rows = []

    def add(mod):
        v = base.copy()
        mod(v)
        rows.append(v)

    for h in (0.0, 1e-18, 1e-300, 95.0, 150.0, 300.0, 1e4, 1e300):
        add(lambda v, h=h: v.__setitem__(slice(0, nr), h))          # every region
        add(lambda v, h=h: v.__setitem__(slice(0, nr, 2), h))       # every other region: mixed lanes
    for x in (1e8, 1e-30, 0.0, 1e300):
        add(lambda v, x=x: v.__setitem__(slice(off + 2, n, 3), x))  # X (backscatter) of every scene
        add(lambda v, x=x: v.__setitem__(off + 2, x))               # of the first scene only
    for pv in (1e-10, 1e-300, 0.0, 1e6):
        add(lambda v, pv=pv: v.__setitem__(slice(off, n, 3), pv))   # P
        add(lambda v, pv=pv: v.__setitem__(slice(off + 1, n, 3), pv))  # G
    for b in (1e5, 1e-300, 0.0, 1e300):
        add(lambda v, b=b: v.__setitem__(slice(nr, nr + nr * nb), b))  # bottom albedo
    add(lambda v: v.__setitem__(slice(nr + nr * nb, nr + nr * nb + nb), 0.0))  # q all zero in region 0: 0/0
    add(lambda v: v.__setitem__(slice(nr + nr * nb, off), 1e-320))            # denormal mixing weights
    add(lambda v: v.__setitem__(slice(0, n), 1e200))
    add(lambda v: v.__setitem__(slice(0, n), -1e-200))
    add(lambda v: v.__setitem__(0, float("inf")))
    add(lambda v: v.__setitem__(1, float("nan")))
    return np.array(rows)


def make_kat_extreme():
    """samodel_error of the UNMODIFIED reference on extreme parameter vectors (fallback paths of the device code)."""
    build()
    ref = Oracle("reference")
    rng = np.random.default_rng(20261020)
    kat = {}
    for tag, ns, nb, nr in (("a", 6, 3, 9), ("b", 8, 1, 9), ("c", 4, 2, 4)):
        spec = replace(scene.CONFIGS["murion"], n_dates=ns)
        cfg = SceneCfg.from_spec(spec)
        meas = rng.uniform(0.002, 0.02, (nr, ns, 4))
        params = extreme_params(rng, ns, nb, nr)
        origin = nr // 2
        out, rrs, K = ref.error_kat(cfg, nb, nr, origin, meas, params)
        print("kat_extreme", tag, params.shape, "finite objective in", int(np.isfinite(out[:, 0]).sum()), "of", len(out))
        for k, v in (("meas", meas), ("params", params), ("out", out), ("meta", np.array([ns, nb, nr, origin]))):
            kat[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "kat_objective_extreme.npz"), **kat)


def make_jerlov():
    """COMPUTE K through the reference's own jerlov.c (jerlov, compute_k_from_jerlov, compute_k_from_ratio)."""
    import ctypes as C
    from oracle.binding import REF_SO
    build()
    L = C.CDLL(REF_SO)
    fp = C.POINTER(C.c_float)
    L.ref_jerlov_fit.argtypes = [C.c_float] * 4 + [fp, fp, C.c_int, C.c_float, fp]
    L.ref_jerlov_k.argtypes = [C.c_float, fp, C.c_int, fp]
    L.ref_jerlov_k_from_ratio.argtypes = [C.c_float, C.c_float, C.c_float, fp, C.c_int, fp, fp]
    P = lambda a: a.ctypes.data_as(fp)  # noqa: E731
    rng = np.random.default_rng(20261019)
    bands = np.array([443.0, 482.0, 561.0, 655.0], dtype=np.float32)
    # the reference printf()s its progress: keep the generator's stdout readable
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        fits = []
        for case in range(96):
            bi, bj = rng.choice(4, 2, replace=False)
            wi, wj = float(bands[bi]), float(bands[bj])
            if case % 8 == 7:  # wavelengths off the band centres, incl. the table's ends
                wi, wj = [float(v) for v in rng.choice([400.0, 425.0, 512.5, 699.0, 700.0, 450.0], 2, replace=False)]
            npts = int(rng.integers(3, 400))
            z = np.sort(rng.uniform(0.5, 25.0, npts))
            ki_true, slope = rng.uniform(0.02, 0.5), rng.uniform(0.2, 6.0)
            lsmi, lsmj = float(rng.uniform(5.0, 60.0)), float(rng.uniform(5.0, 60.0))
            Li = (lsmi + rng.uniform(200, 4000) * np.exp(-2.0 * ki_true * z) * (1 + 0.03 * rng.standard_normal(npts))).astype(np.float32)
            Lj = (lsmj + rng.uniform(200, 4000) * np.exp(-2.0 * ki_true * slope * z) * (1 + 0.03 * rng.standard_normal(npts))).astype(np.float32)
            manual = 0.0
            if case % 6 == 5:
                manual = float(rng.uniform(0.3, 5.0))
            if case == 10:  # every point below the deep-water signal: no shallow points, singular fit
                Li[:] = lsmi
            if case == 11:  # all points identical: singular fit
                Li[:] = Li[0]
                Lj[:] = Lj[0]
            if case == 12:  # wavelength outside the table
                wi = 380.0
            out = np.zeros(6, dtype=np.float32)
            ok = L.ref_jerlov_fit(wi, wj, lsmi, lsmj, P(Li), P(Lj), npts, manual, P(out))
            fits.append((wi, wj, lsmi, lsmj, manual, Li, Lj, ok, out))
        wts = np.concatenate([np.linspace(0.0, 8.999, 61), rng.uniform(0.0, 9.0, 67)]).astype(np.float32)
        wls = np.concatenate([bands, [400.0, 700.0, 399.9, 700.1, 425.0, 512.3, 675.0, 0.0],
                              rng.uniform(400.0, 700.0, 20)]).astype(np.float32)
        kk = np.zeros((len(wts), len(wls)), dtype=np.float32)
        for a, wt in enumerate(wts):
            L.ref_jerlov_k(float(wt), P(wls), len(wls), P(kk[a]))
        ratios = []
        for case in range(64):
            bi, bj = rng.choice(4, 2, replace=False)
            ratio = float(np.float32(rng.uniform(0.05, 12.0)))
            wt = C.c_float(0.0)
            k = np.zeros(len(wls), dtype=np.float32)
            ok = L.ref_jerlov_k_from_ratio(ratio, float(bands[bi]), float(bands[bj]), P(wls), len(wls), C.byref(wt), P(k))
            ratios.append((ratio, float(bands[bi]), float(bands[bj]), ok, np.float32(wt.value), k))
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    print("jerlov: fits ok", sum(f[7] for f in fits), "of", len(fits), "; ratio cases ok", sum(r[3] for r in ratios), "of", len(ratios))
    lens = np.array([len(f[5]) for f in fits], dtype=np.int32)
    np.savez_compressed(
        os.path.join(HERE, "jerlov.npz"),
        fit_args=np.array([f[:5] for f in fits], dtype=np.float32), fit_len=lens,
        fit_Li=np.concatenate([f[5] for f in fits]), fit_Lj=np.concatenate([f[6] for f in fits]),
        fit_ok=np.array([f[7] for f in fits], dtype=np.int32), fit_out=np.stack([f[8] for f in fits]),
        k_wt=wts, k_wl=wls, k_out=kk,
        ratio_args=np.array([r[:3] for r in ratios], dtype=np.float32), ratio_ok=np.array([r[3] for r in ratios], dtype=np.int32),
        ratio_wt=np.array([r[4] for r in ratios], dtype=np.float32), ratio_k=np.stack([r[5] for r in ratios]))


# (config, non-converged pixels kept, high-evaluation-count converged pixels kept)
MINED = (("exmouth", 40, 10), ("qatar", 40, 10), ("abudhabi", 10, 6), ("pilbara", 17, 6))


def make_mined():
    """Reference goldens for the pixels the small scenes never contain: nelmin's `ifault = 2` (kcount exhausted,
    asa047.c:217, 411-453, 481-493 -> md->converged = false, samodel.c:2396-2402), restarts after a failed factorial
    test (numres >= 1) and evaluation counts at the kcount edge. The pixels were found on the GPU in full-size scenes
    (tests/manual/mine_nonconverged.py -> gpurun_out/mined_<config>.npz: 3x3 input neighbourhoods side by side, patch
    k = columns [3k, 3k+3), centre (1, 3k+1), global coordinates kept). Here the UNMODIFIED reference inverts the
    centres of a subset; the device's own evaluation counts / flags from the mining run are kept beside them only as
    a record (the tests compare fresh device results with the reference's)."""
    build()
    ref = Oracle("reference")
    for name, n_nonconv, n_high in MINED:
        src = os.path.join(ROOT, "gpurun_out", f"mined_{name}.npz")
        if not os.path.exists(src):
            print("skip", name, "(no mining output)")
            continue
        m = np.load(src)
        spec = scene.CONFIGS[name]
        conv, ev = m["dev_converged"], m["dev_evals"]
        bad = np.nonzero(conv == 0)[0]
        bad = bad[np.linspace(0, len(bad) - 1, min(n_nonconv, len(bad))).round().astype(int)]
        good = np.nonzero(conv == 1)[0]
        _, first = np.unique(ev[good], return_index=True)      # one pixel per distinct evaluation count, highest first
        good = good[np.sort(first)][:n_high]
        keep = np.concatenate([bad, good])
        K = len(keep)
        planes = np.concatenate([m["planes"][:, :, 3 * k:3 * k + 3] for k in keep], axis=2)
        prior = np.concatenate([m["prior"][:, 3 * k:3 * k + 3] for k in keep], axis=1)
        ci, cj = np.ones(K, dtype=np.int32), (3 * np.arange(K) + 1).astype(np.int32)
        cfg = SceneCfg.from_spec(spec)
        out = ref.invert_pixels(cfg, planes, scene.NODATA, prior, scene.NODATA, ci, cj, nthreads=8)
        assert (out["status"] == 1).all()
        agree = int((out["n_evals"] == ev[keep]).sum())
        np.savez_compressed(
            os.path.join(HERE, f"mined_{name}.npz"), planes=planes, prior=prior, use_prior=True,
            wavelengths=np.array(spec.wavelengths), theta_view=spec.theta_view,
            theta_sun=np.array([spec.theta_sun(s) for s in range(spec.n_dates)]),
            h_tide=np.array([spec.h_tide(s) for s in range(spec.n_dates)]), n_smooth=spec.n_smoothing_radius,
            n_spatial=spec.n_spatial, n_bottoms=spec.n_bottoms, nodata=scene.NODATA, centre_i=ci, centre_j=cj,
            global_i=m["gi"][keep], global_j=m["gj"][keep], rec=out["rec"], status=out["status"],
            converged=out["converged"], n_evals=out["n_evals"], n_restarts=out["n_restarts"],
            mining_run_evals=ev[keep], mining_run_converged=conv[keep])
        print(f"mined_{name}: {K} centres, reference says non-converged {int((out['converged'] == 0).sum())}, "
              f"restarts>=1 {int((out['n_restarts'] >= 1).sum())}, max evals {int(out['n_evals'].max())}, "
              f"mining-run evaluation counts equal to the reference's: {agree}/{K}")


# pixels of seeded small scenes whose factorial test fails once and restarts nelmin (numres = 1, asa047.c:445-493),
# found with the CPU restatement (n_restarts of oracle.binding.Oracle.invert_pixels): (config, rows, cols, pixels)
RESTARTS = (("exmouth", 160, 140, ((55, 39), (60, 135))), ("qatar", 120, 120, ((90, 86),)))


def make_restarts():
    """Same patch layout as make_mined() for pixels on which nelmin restarts (numres >= 1), each with its right-hand
    neighbour as an ordinary pixel beside it; the reference's restart count comes from the wrapped nelmin of
    oracle/ref_harness.c (the reference keeps numres in a local and drops it, samodel.c:2146, 2373)."""
    build()
    ref = Oracle("reference")
    for name, R, Cc, pixels in RESTARTS:
        spec = scene.CONFIGS[name].scaled(R, Cc)
        planes, prior = scene.generate(spec)
        planes, prior = planes.numpy(), prior.numpy()
        pts = [q for (i, j) in pixels for q in ((i, j), (i, j + 1))]
        pp = np.concatenate([planes[:, i - 1:i + 2, j - 1:j + 2] for i, j in pts], axis=2)
        pr = np.concatenate([prior[i - 1:i + 2, j - 1:j + 2] for i, j in pts], axis=1)
        K = len(pts)
        ci, cj = np.ones(K, dtype=np.int32), (3 * np.arange(K) + 1).astype(np.int32)
        out = ref.invert_pixels(SceneCfg.from_spec(spec), pp, scene.NODATA, pr, scene.NODATA, ci, cj, nthreads=4)
        assert (out["n_restarts"][::2] >= 1).all(), out["n_restarts"]
        np.savez_compressed(
            os.path.join(HERE, f"mined_restart_{name}.npz"), planes=pp, prior=pr, use_prior=True,
            wavelengths=np.array(spec.wavelengths), theta_view=spec.theta_view,
            theta_sun=np.array([spec.theta_sun(s) for s in range(spec.n_dates)]),
            h_tide=np.array([spec.h_tide(s) for s in range(spec.n_dates)]), n_smooth=spec.n_smoothing_radius,
            n_spatial=spec.n_spatial, n_bottoms=spec.n_bottoms, nodata=scene.NODATA, centre_i=ci, centre_j=cj,
            global_i=np.array([p[0] for p in pts], dtype=np.int32), global_j=np.array([p[1] for p in pts], dtype=np.int32),
            rec=out["rec"], status=out["status"], converged=out["converged"], n_evals=out["n_evals"],
            n_restarts=out["n_restarts"])
        print(f"mined_restart_{name}: restarts", out["n_restarts"].tolist(), "evals", out["n_evals"].tolist(), "status",
              out["status"].tolist())


REFINE_FLAGS = (0, 1, 2, 3, 4, 8, 16, 1 | 2 | 4 | 8 | 16, 2 | 16, 1 | 8, 2 | 4, 1 | 2 | 16)  # include/photic_b200.h PHB_REFINE_*


def refine_inputs():
    """Inputs of the REFINE goldens (also regenerated by the tests: only the outputs need storing)."""
    rng = np.random.default_rng(11)
    grid = (-rng.uniform(0.2, 35.0, (57, 43))).astype(np.float32)
    grid[rng.uniform(size=grid.shape) < 0.2] = -9999.0
    grid[3, 4], grid[5, 6], grid[7, 8] = 2.5, 0.0, -9998.5   # positive depth, zero, a value approx_equal() to nodata at 1e-4
    land = np.where(rng.uniform(size=grid.shape) < 0.3, -9999.0, 1.0).astype(np.float32)
    shallow = np.where(rng.uniform(size=grid.shape) < 0.2, -7777.0, 1.0).astype(np.float32)
    args = np.array([-30.0, -0.5, -40.0, 0.0, 1.3, 0.9, -0.25, -32.0, -1.0, 1.1, 0.95], dtype=np.float32)
    return grid, land, shallow, args


def make_refine():
    """REFINE through the reference's own run_refine() (refine.c:12-302, driven by ref_refine in oracle/ref_harness.c):
    every flag set of the GPU test, with both masks, one mask only (the reference then blanks the whole grid,
    refine.c:242-244) and none; SHAPE 1.0 (linear rescale branch); min/max taken from the grid when CLIP is absent."""
    build()
    ref = Oracle("reference")
    grid, land, shallow, args = refine_inputs()
    res = {"grid": grid, "land": land, "shallow": shallow, "args": args, "flags": np.array(REFINE_FLAGS, dtype=np.int32)}
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)  # run_refine printf()s its POWER arguments
    try:
        for flags in REFINE_FLAGS:
            for mk, (ld, sh) in enumerate(((None, None), (land, shallow), (land, None), (None, shallow))):
                for shape1 in (0, 1):
                    if shape1 and not (flags & 2):
                        continue
                    a2 = args.copy()
                    if shape1:
                        a2[4] = 1.0
                    res[f"out_{flags}_{mk}_{shape1}"] = ref.refine(grid, -9999.0, ld, -9999.0, sh, -7777.0, flags, a2)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
    np.savez_compressed(os.path.join(HERE, "refine.npz"), **res)
    print("refine:", sum(k.startswith("out_") for k in res), "cases;", "open cells in the full-flag case",
          int((res["out_31_1_0"] != -9999.0).sum()))


def nc_pack_cases():
    """Grids for the int16 packing goldens (small: stored with their outputs). The order-dependent range of
    compress_2d (nc.c:286-298) is what the cases are about; a 300 x 333 grid spans many 4096-cell chunks of the
    device passes, with record minima placed deep inside it."""
    rng = np.random.default_rng(5)
    g = rng.uniform(-40, 0, (37, 29)).astype(np.float32)
    g[rng.uniform(size=g.shape) < 0.3] = -9999.0
    yield "random", g
    d = np.sort(rng.uniform(-40, 5, (11, 13)).astype(np.float32).ravel())[::-1].reshape(11, 13).copy()
    yield "decreasing", d          # every value lowers the minimum: the maximum stays FLT_MIN and the shorts wrap
    yield "increasing", d.ravel()[::-1].reshape(11, 13).copy()
    f = g.copy()
    f[0, 0] = 100.0                # the first value is the largest: it never reaches the maximum
    yield "first_is_max", f
    yield "flat", np.full((5, 7), 3.25, dtype=np.float32)
    yield "all_spval", np.full((4, 4), -9999.0, dtype=np.float32)
    h = g.copy()
    h[2, 3], h[4, 5], h[6, 7] = np.nan, np.inf, -np.inf
    yield "nan_inf", h
    p = rng.uniform(1e-3, 50, (9, 9)).astype(np.float32)
    p[0, 0] = -9999.0
    yield "positive", p
    big = rng.uniform(-30, -1, (300, 333)).astype(np.float32)
    big[rng.uniform(size=big.shape) < 0.4] = -9999.0
    flat = big.ravel()
    for k, v in ((5000, 7.0), (5001, -35.0), (40000, 9.0), (40001, -36.0), (40002, 8.5), (90000, -37.0), (99000, 12.0)):
        flat[k] = v                # a maximum right before a new minimum, one that IS a new minimum's neighbour, a late record
    flat[99899] = -50.0            # a record minimum in the last chunk that would otherwise be the ...
    yield "many_chunks", big
    dec = np.linspace(20.0, -20.0, 3 * 4096 + 77, dtype=np.float32).reshape(1, -1).copy()
    dec[0, 6000] = 19.5            # the only value that does not lower the running minimum, two chunks in
    yield "decreasing_many_chunks", dec


def make_nc_pack():
    """compress_2d / decompress_2d of the unmodified nc.c (oracle/ref_nc.c includes it where it lies)."""
    build()
    ref = Oracle("reference")
    res = {}
    for name, g in nc_pack_cases():
        packed, off, sc, miss = ref.nc_pack(g, -9999.0)
        res[f"{name}_grid"] = g
        res[f"{name}_packed"] = packed
        res[f"{name}_meta"] = np.array([off, sc], dtype=np.float32)
        res[f"{name}_missing"] = np.array([miss], dtype=np.int16)
        res[f"{name}_unpacked"] = ref.nc_unpack(packed, off, sc, miss, -9999.0)
    np.savez_compressed(os.path.join(HERE, "nc_pack.npz"), **res)
    print("nc_pack:", len(res) // 5, "cases;", {n: (float(res[f"{n}_meta"][0]), float(res[f"{n}_meta"][1])) for n, _ in nc_pack_cases()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nc_pack":
        make_nc_pack()
    elif len(sys.argv) > 1 and sys.argv[1] == "refine":
        make_refine()
    elif len(sys.argv) > 1 and sys.argv[1] == "mined":
        make_mined()
    elif len(sys.argv) > 1 and sys.argv[1] == "restarts":
        make_restarts()
    elif len(sys.argv) > 1 and sys.argv[1] == "depth_sigma":
        make_depth_sigma()
    elif len(sys.argv) > 1 and sys.argv[1] == "lee_ls8":
        make_lee_ls8()
    elif len(sys.argv) > 1 and sys.argv[1] == "jerlov":
        make_jerlov()
    elif len(sys.argv) > 1 and sys.argv[1] == "kat_extreme":
        make_kat_extreme()
    else:
        main()
        make_depth_sigma()
        make_lee_ls8()
        make_jerlov()
        make_kat_extreme()
        make_refine()
        make_mined()
        make_restarts()
        make_nc_pack()
