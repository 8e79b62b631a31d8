"""One process, several GPUs: phb_plan_row_bands / phb_invert_host_multi (include/photic_b200.h).
The planner is host code and is tested here on the CPU against the Python planner of photic_b200/sharded.py; the
multi-context inversion is a GPU test and must equal the one-context result bit for bit."""
import numpy as np
import pytest

from photic_b200 import capi, scene, sharded
from photic_b200.samodel import Inverter


def _scene(name, R, C):
    spec = scene.CONFIGS[name].scaled(R, C)
    planes, prior = scene.generate(spec)
    return spec, planes, prior


@pytest.mark.parametrize("name,R,C", [("exmouth", 240, 160), ("qatar", 150, 90)])
def test_host_planner_matches_python_planner(product_lib, name, R, C):
    spec, planes, prior = _scene(name, R, C)
    desc = capi.desc_from_spec(spec)
    valid = ((planes >= 0) & (planes != scene.NODATA)).all(dim=0)
    want_cost = sharded.row_cost_from_prior(valid, prior).numpy().astype(np.float64)
    for parts in (1, 2, 3, 8):
        edges, cost = Inverter.plan_row_bands_host(desc, planes.numpy(), prior.numpy(), parts)
        assert np.allclose(cost, want_cost, rtol=1e-6, atol=1e-3)
        plan = sharded.plan_row_bands(cost, parts)
        assert edges.tolist() == [a for a, _ in plan] + [plan[-1][1]]
        assert edges[0] == 0 and edges[-1] == R and np.all(np.diff(edges) >= 0)
        if parts > 1 and cost.sum() > 0:  # balance: no band carries more than its share plus one row
            band = np.add.reduceat(cost, edges[:-1][np.diff(edges) > 0])
            assert band.max() <= cost.sum() / parts + cost.max() + 1e-9


def test_host_planner_degenerate_inputs(product_lib):
    spec, planes, prior = _scene("murion", 6, 5)
    desc = capi.desc_from_spec(spec)
    land = np.full_like(planes.numpy(), scene.NODATA)
    edges, cost = Inverter.plan_row_bands_host(desc, land, prior.numpy(), 4)  # nothing valid: equal split of rows
    assert cost.sum() == 0 and edges.tolist() == [0, 2, 3, 5, 6]  # llround(6 k / 4)
    edges, _ = Inverter.plan_row_bands_host(desc, planes.numpy(), prior.numpy(), 16)  # more bands than rows
    assert edges[0] == 0 and edges[-1] == 6 and np.all(np.diff(edges) >= 0)
    d2 = capi.desc_from_spec(spec, prior_present=False)  # no DEPTHS grid: every valid pixel costs the same
    _, c2 = Inverter.plan_row_bands_host(d2, planes.numpy(), None, 2)
    valid = ((planes >= 0) & (planes != scene.NODATA)).all(dim=0).numpy()
    assert np.allclose(c2, valid.sum(axis=1) * 8 * 2.4, rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name,R,C,parts", [("abudhabi", 36, 24, 2), ("murion", 40, 28, 3), ("qatar", 30, 26, 5)])
def test_invert_host_multi_equals_single_context(inverter, name, R, C, parts):
    import torch
    spec, planes, prior = _scene(name, R, C)
    desc = capi.desc_from_spec(spec)
    pl, pr = planes.numpy(), prior.numpy()
    one, st1 = inverter.invert_host(desc, pl, pr)
    ndev = torch.cuda.device_count()
    ivs = [inverter] + [Inverter(k % ndev) for k in range(1, parts)]  # other devices when there are any, else device 0 again
    try:
        many, stn = Inverter.invert_host_multi(ivs, desc, pl, pr)
    finally:
        for iv in ivs[1:]:
            iv.close()
    assert stn["edges"][0] == 0 and stn["edges"][-1] == R
    assert sum(p["n_valid"] for p in stn["per_ctx"]) == stn["n_valid"] == st1["n_valid"] > 50
    assert stn["n_evals"] == st1["n_evals"] and stn["n_converged"] == st1["n_converged"]
    for key in one:
        assert np.array_equal(one[key].view(np.uint8), many[key].view(np.uint8)), key


# ---- row bands that share their work (phb_shard_*): the device-resident multi-GPU path --------------------------

def _whole(inverter, spec, planes, prior):
    return inverter.invert_host(capi.desc_from_spec(spec), planes.numpy(), prior.numpy(), scene_planes=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name,R,C,parts", [("exmouth", 30, 22, 2), ("qatar", 33, 20, 3)])
def test_bands_take_work_from_each_other(inverter, name, R, C, parts):
    """Bands of one scene, each with its halo rows, all on this device: the first solve finds its own queue, then
    empties its peers' queues through their handles (the path a second GPU takes over NVLink); every band's result
    planes must equal the unsharded result bit for bit, whoever computed the pixel."""
    import torch
    from photic_b200.samodel import Band
    spec, planes, prior = _scene(name, R, C)
    one, st1 = _whole(inverter, spec, planes, prior)
    halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
    plan = sharded.equal_row_bands(R, parts)
    bands, wins = [], []
    for r0, r1 in plan:
        w0, w1, lb, le = sharded.window(r0, r1, halo, R)
        b = Band(inverter, capi.desc_from_spec(spec, nrows=w1 - w0), lb, le)
        b.planes.copy_(planes[:, w0:w1].cuda())
        b.prior.copy_(prior[w0:w1].cuda())
        bands.append(b)
        wins.append((w0, w1, lb, le, r0, r1))
    hs = [b.export() for b in bands]
    for b in bands:
        b.prepare()
    torch.cuda.synchronize()
    assert sum(b.valid() for b in bands) == st1["n_valid"]
    stats = [b.solve([hs[(k + q) % parts] for q in range(1, parts)]) for k, b in enumerate(bands)]
    torch.cuda.synchronize()
    assert stats[0]["n_valid"] == st1["n_valid"] and all(s["n_valid"] == 0 for s in stats[1:])  # the first took it all
    assert sum(s["n_evals"] for s in stats) == st1["n_evals"]
    for b, (w0, w1, lb, le, r0, r1) in zip(bands, wins):
        for key in capi.SCALAR_PLANES + ("converged", "n_evals"):
            got = b.outputs[key][lb:le].cpu().numpy()
            assert np.array_equal(got.view(np.uint8), one[key][r0:r1].view(np.uint8)), (key, r0, r1)
    for b in bands:
        b.close()


def _ipc_worker(rank, world, port, name, R, C, out_path):
    import os, sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from photic_b200 import capi, scene, sharded
    from photic_b200.samodel import Inverter
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    inv = Inverter(dev)
    spec = scene.CONFIGS[name].scaled(R, C)
    halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
    r0, r1 = sharded.equal_row_bands(R, world)[rank]
    w0, w1, lb, le = sharded.window(r0, r1, halo, R)
    planes, prior = scene.generate(spec, w0, w1, device=torch.device("cuda", dev))
    grp = sharded.BandGroup(inv, capi.desc_from_spec(spec, nrows=w1 - w0), lb, le, rank, world)
    grp.band.planes.copy_(planes)
    grp.band.prior.copy_(prior)
    st = grp.step()
    out = {k: grp.band.outputs[k][lb:le].cpu().numpy() for k in capi.SCALAR_PLANES + ("converged", "n_evals")}
    np.savez(out_path % rank, r0=r0, r1=r1, n_valid=st["n_valid"], shared=int(st["shared"]), **out)
    dist.barrier()
    grp.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_bands_across_processes_map_each_other(inverter, tmp_path):
    """One process per band (as under torch.distributed.run): the handles cross the process boundary as bytes and are
    opened with CUDA IPC; a band's pixels may be inverted by the other process. Gathered result == unsharded."""
    import socket
    import torch.multiprocessing as mp
    name, R, C, world = "murion", 26, 18, 2
    spec, planes, prior = _scene(name, R, C)
    one, st1 = _whole(inverter, spec, planes, prior)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    pat = str(tmp_path / "band%d.npz")
    mp.spawn(_ipc_worker, args=(world, port, name, R, C, pat), nprocs=world, join=True)
    total = 0
    for rank in range(world):
        g = np.load(pat % rank)
        assert int(g["shared"]) == 1
        total += int(g["n_valid"])
        for key in capi.SCALAR_PLANES + ("converged", "n_evals"):
            assert np.array_equal(g[key].view(np.uint8), one[key][int(g["r0"]):int(g["r1"])].view(np.uint8)), (key, rank)
    assert total == st1["n_valid"]


@pytest.mark.gpu
def test_row_pointer_entry_equals_plane_entry(inverter):
    """phb_invert_rows: rasters as the reference holds them (float **, one allocation per row) in and out, through the
    pinned staging ring; two contexts. Must equal phb_invert_host on contiguous planes, K / P / G / X included."""
    import ctypes as C
    spec, planes, prior = _scene("murion", 23, 17)
    desc = capi.desc_from_spec(spec)
    pl, pr = planes.numpy(), prior.numpy()
    one, st1 = inverter.invert_host(desc, pl, pr)
    R, Cc, ns, mb = spec.nrows, spec.ncols, spec.n_dates, 4
    keep = []
    fp = C.POINTER(C.c_float)

    def rows_of(a2d):  # every row its own allocation, like allocate_float_array_2d (common.c:562-572)
        rws = [np.array(a2d[r], dtype=np.float32, copy=True) for r in range(a2d.shape[0])]
        arr = (fp * len(rws))(*[r.ctypes.data_as(fp) for r in rws])
        keep.extend([rws, arr])
        return arr, rws

    plane_rows = (C.POINTER(fp) * pl.shape[0])()
    for g in range(pl.shape[0]):
        arr, _ = rows_of(pl[g])
        plane_rows[g] = C.cast(arr, C.POINTER(fp))
    prior_rows, _ = rows_of(pr)
    out = capi.RowOutputs()
    got = {}
    for name in capi.SCALAR_PLANES:
        arr, rws = rows_of(np.full((R, Cc), 7.0, dtype=np.float32))
        setattr(out, name, C.cast(arr, C.POINTER(fp)))
        got[name] = rws
    for name, count in (("K", ns * mb), ("P", ns), ("G", ns), ("X", ns)):
        stack = (C.POINTER(fp) * count)()
        got[name] = []
        for q in range(count):
            arr, rws = rows_of(np.full((R, Cc), 7.0, dtype=np.float32))
            stack[q] = C.cast(arr, C.POINTER(fp))
            got[name].append(rws)
        keep.append(stack)
        setattr(out, name, C.cast(stack, C.POINTER(C.POINTER(fp))))
    iv2 = Inverter(0)
    try:
        ctxs = (C.c_void_p * 2)(inverter.ctx.value, iv2.ctx.value)
        st = capi.Stats()
        capi.check(capi.lib().phb_invert_rows(ctxs, 2, C.byref(desc), C.cast(plane_rows, C.c_void_p),
                                              C.cast(prior_rows, C.c_void_p), C.byref(out), C.byref(st), None, None))
    finally:
        iv2.close()
    assert st.n_valid == st1["n_valid"] and st.n_evals == st1["n_evals"]
    for name in capi.SCALAR_PLANES:
        assert np.array_equal(np.stack(got[name]).view(np.int32), one[name].view(np.int32)), name
    assert np.array_equal(np.stack([np.stack(r) for r in got["K"]]).reshape(ns, mb, R, Cc).view(np.int32), one["K"].view(np.int32))
    for name in ("P", "G", "X"):
        assert np.array_equal(np.stack([np.stack(r) for r in got[name]]).view(np.int32), one[name].view(np.int32)), name


@pytest.mark.gpu
def test_every_grid_is_tested_against_its_own_nodata(inverter):
    """samodel.c:683, 941, 2999-3003 compare band k with gridded_data[k].nodata_value. Scene B = scene A with the nodata
    cells of three grids re-coded to 0, -1 and 12345, and those values declared per grid: same result as scene A."""
    spec, planes, prior = _scene("exmouth", 26, 21)
    pl, pr = planes.numpy().copy(), prior.numpy()
    one, st1 = inverter.invert_host(capi.desc_from_spec(spec), pl, pr, scene_planes=False)
    nd = [[float(scene.NODATA)] * 4 for _ in range(spec.n_dates)]
    for (s, b), v in {(0, 1): 0.0, (2, 3): -1.0, (5, 0): 12345.0}.items():
        g = 4 * s + b
        assert not (pl[g] == v).any()
        pl[g][pl[g] == scene.NODATA] = v
        nd[s][b] = v
    desc = capi.make_desc(spec.wavelengths, spec.theta_view, [spec.theta_sun(s) for s in range(spec.n_dates)],
                          [spec.h_tide(s) for s in range(spec.n_dates)], spec.nrows, spec.ncols, nodata_band=nd,
                          r_sigma=spec.r_sigma)
    two, st2 = inverter.invert_host(desc, pl, pr, scene_planes=False)
    assert st2["n_valid"] == st1["n_valid"] > 50 and st2["n_evals"] == st1["n_evals"]
    for key in one:
        assert np.array_equal(one[key].view(np.uint8), two[key].view(np.uint8)), key
