"""Host-side multi-GPU logic on CPU: gloo, world_size 2 and 3. The per-shard compute is the CPU oracle
(test-only stand-in for the CUDA library), so what is checked is the sharding itself: cost-balanced
row bands, halo exchange, edge clamping at the GLOBAL image edge, and the final gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, R, C, out_path):
    import sys
    sys.path.insert(0, ROOT)
    from oracle.binding import Oracle, SceneCfg
    from photic_b200 import scene, sharded
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = scene.CONFIGS[name].scaled(R, C)
    halo = sharded.halo_rows(spec.n_spatial, spec.n_smoothing_radius)
    # phase 1: every rank estimates the cost of an equal split of rows, all-gather, plan balanced bands
    eq = sharded.plan_row_bands(np.ones(spec.nrows), world)
    e0, e1 = eq[rank]
    planes_eq, prior_eq = scene.generate(spec, e0, e1)
    cost = torch.zeros(spec.nrows)
    cost[e0:e1] = sharded.row_cost(scene.valid_mask(planes_eq), prior_eq.abs() <= 8.0)
    dist.all_reduce(cost)
    plan = sharded.plan_row_bands(cost.numpy(), world)
    # phase 2: rows move from the equal split to the cost-balanced bands (point-to-point), then each rank
    # holds ONLY its own rows; halo rows come from the neighbours
    r0, r1 = plan[rank]
    planes = sharded.repartition_rows(planes_eq, eq, plan, rank, world)
    prior = sharded.repartition_rows(prior_eq[None], eq, plan, rank, world)[0]
    chk_p, chk_pr = scene.generate(spec, r0, r1)
    assert torch.equal(planes, chk_p) and torch.equal(prior, chk_pr)
    win = sharded.exchange_halo(planes, plan, halo, rank, world)
    w0, w1, lb, le = sharded.window(r0, r1, halo, spec.nrows)
    assert win.shape[1] == w1 - w0
    # the exchanged window must equal the same rows generated directly
    direct, _ = scene.generate(spec, w0, w1)
    assert torch.equal(win, direct)
    prior_win = sharded.exchange_halo(prior[None], plan, halo, rank, world)[0]
    # per-shard inversion of rows [lb, le) of the window (oracle stands in for the GPU here)
    port_o = Oracle("port")
    cfg = SceneCfg.from_spec(spec)
    ii, jj = np.meshgrid(np.arange(lb, le), np.arange(spec.ncols), indexing="ij")
    res = port_o.invert_pixels(cfg, win.numpy(), scene.NODATA, prior_win.numpy(), scene.NODATA, ii.ravel(), jj.ravel(), nthreads=2)
    depth = torch.from_numpy(res["rec"][:, 0].reshape(le - lb, spec.ncols).copy())
    evals = torch.from_numpy(res["n_evals"].reshape(le - lb, spec.ncols).astype(np.int32))
    full_d = sharded.gather_bands(depth, plan, rank, world)
    full_e = sharded.gather_bands(evals, plan, rank, world)
    lo, hi = sharded.allreduce_minmax(float(depth.min()), float(depth.max()), "cpu")
    if rank == 0:
        np.savez(out_path, depth=full_d.numpy(), evals=full_e.numpy(), plan=np.array(plan), minmax=np.array([lo, hi]),
                 cost=cost.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,R,C", [(2, "murion", 14, 12), (3, "exmouth", 13, 9)])
def test_sharded_equals_unsharded(tmp_path, oracle_port, world, name, R, C):
    from oracle.binding import SceneCfg
    from photic_b200 import scene
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), name, R, C, out), nprocs=world, join=True)
    g = np.load(out)
    spec = scene.CONFIGS[name].scaled(R, C)
    planes, prior = scene.generate(spec)
    ii, jj = np.meshgrid(np.arange(R), np.arange(C), indexing="ij")
    ref = oracle_port.invert_pixels(SceneCfg.from_spec(spec), planes.numpy(), scene.NODATA, prior.numpy(), scene.NODATA,
                                    ii.ravel(), jj.ravel())
    assert np.array_equal(g["evals"].ravel(), ref["n_evals"])
    assert np.array_equal(g["depth"].ravel().view(np.int64), ref["rec"][:, 0].view(np.int64))
    plan = g["plan"]
    assert plan[0][0] == 0 and plan[-1][1] == R and all(plan[k][1] == plan[k + 1][0] for k in range(world - 1))
    d = ref["rec"][:, 0]
    assert g["minmax"][0] == np.float32(d.min()) and g["minmax"][1] == np.float32(d.max())


def test_plan_row_bands_balances_cost():
    from photic_b200 import sharded
    rng = np.random.default_rng(0)
    cost = np.concatenate([np.zeros(300), rng.uniform(1, 3, 500), np.zeros(100), rng.uniform(5, 9, 124)])
    for world in (1, 2, 4, 8):
        plan = sharded.plan_row_bands(cost, world)
        assert plan[0][0] == 0 and plan[-1][1] == len(cost)
        sums = np.array([cost[a:b].sum() for a, b in plan])
        assert sums.max() <= cost.sum() / world + cost.max() + 1e-9     # within one row of perfect balance
    assert sharded.plan_row_bands(np.zeros(10), 4) == [(0, 2), (2, 5), (5, 8), (8, 10)]
    assert sharded.window(0, 5, 1, 20) == (0, 6, 0, 5) and sharded.window(15, 20, 1, 20) == (14, 20, 1, 6)
    assert sharded.halo_rows(2, 1) == 1 and sharded.halo_rows(0, 1) == 0 and sharded.halo_rows(3, 2) == 3


def test_scene_generator_is_window_invariant():
    """Any row window of a scene is bit-identical to the same rows of the full scene (rank independence)."""
    from photic_b200 import scene
    spec = scene.CONFIGS["qatar"].scaled(40, 30)
    full, prior = scene.generate(spec)
    part, pp = scene.generate(spec, 13, 29, chunk_rows=5)
    assert torch.equal(full[:, 13:29], part) and torch.equal(prior[13:29], pp)
    m = scene.valid_mask(full)
    assert 0.2 < m.float().mean() < 0.9
