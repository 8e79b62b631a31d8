"""COMPUTE K (model/jerlov.c) -- SURVEY.md row N3: the product's host functions phb_jerlov_* against known answers of
the UNMODIFIED reference (tests/golden/jerlov.npz, made by tests/golden/make_golden.py from oracle/_ref).
Host code by design (scene-level scalars), so these parity tests need no GPU. Bar: bit equality of every float."""
import os

import numpy as np
import pytest

from photic_b200 import jerlov as pj

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jerlov.npz")


@pytest.fixture(scope="module")
def gold(product_lib):
    return np.load(GOLD)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def test_jerlov_fit_matches_reference(gold):
    ofs = np.concatenate([[0], np.cumsum(gold["fit_len"])])
    n_ok = 0
    for c, (wi, wj, lsmi, lsmj, manual) in enumerate(gold["fit_args"]):
        Li, Lj = gold["fit_Li"][ofs[c]:ofs[c + 1]], gold["fit_Lj"][ofs[c]:ofs[c + 1]]
        ok, res = pj.jerlov(wi, wj, lsmi, lsmj, Li, Lj, manual)
        assert ok == bool(gold["fit_ok"][c]), c
        got = np.array([res[k] for k in ("ki", "kj", "m", "c", "r", "water_type")], dtype=np.float32)
        assert np.array_equal(_bits(got), _bits(gold["fit_out"][c])), (c, got, gold["fit_out"][c])
        n_ok += ok
    assert 20 < n_ok < len(gold["fit_args"])  # both the success and the `return false` paths are exercised


def test_compute_k_matches_reference(gold):
    for a, wt in enumerate(gold["k_wt"]):
        k = pj.compute_k(wt, gold["k_wl"])
        assert np.array_equal(_bits(k), _bits(gold["k_out"][a])), (a, wt)
    alphas = np.zeros(8, dtype=np.float32)
    pj.compute_k_from_jerlov(2.5, alphas, [5, 1, 3, 0], [443.0, 482.0, 561.0, 655.0])
    assert np.array_equal(alphas[[5, 1, 3, 0]], pj.compute_k(2.5, [443.0, 482.0, 561.0, 655.0])) and alphas[2] == 0.0
    assert pj.compute_k(np.float32(3.0), 800.0) == 0.0  # outside the table: 0.0, as the reference returns
    with pytest.raises(Exception):
        pj.compute_k(9.0, 443.0)  # the reference reads past its table here; the product refuses


def test_compute_k_from_ratio_matches_reference(gold):
    for c, (ratio, wi, wj) in enumerate(gold["ratio_args"]):
        ok, wt, k = pj.compute_k_from_ratio(ratio, wi, wj, gold["k_wl"])
        assert ok == bool(gold["ratio_ok"][c]), c
        assert _bits(wt) == _bits(gold["ratio_wt"][c]), c
        assert np.array_equal(_bits(k), _bits(gold["ratio_k"][c])), c


def test_water_type_label():
    assert pj.jerlov_water_type_str(2.0) == "OIB + 0.00" and pj.jerlov_water_type_str(5.25) == "C1 + 0.25"
