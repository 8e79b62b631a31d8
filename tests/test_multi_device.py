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
