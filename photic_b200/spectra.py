"""Spectral data tables, read from the single source of truth ``include/photic_spectra.h``.

Only the synthetic-scene generator uses these from Python (to *simulate* reflectance); the
inversion itself reads the same header from C/CUDA. Reference tables: model/samodel.c:103-286.
"""
from __future__ import annotations

import os
import re

import numpy as np

_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "photic_spectra.h")


def _parse(header: str) -> dict[str, np.ndarray]:
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name in ("PH_SPEC_AW", "PH_SPEC_BBW", "PH_SPEC_A0", "PH_SPEC_A1"):
        m = re.search(name + r"\[PH_SPEC_N\]\s*=\s*\{(.*?)\};", src, re.S)
        out[name] = np.array([float(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()])
        assert out[name].shape == (41,)
    m = re.search(r"PH_SPEC_BOTTOM\[PH_N_BOTTOM_TYPES\]\[PH_SPEC_N\]\s*=\s*\{(.*?)\};", src, re.S)
    rows = re.findall(r"\{(.*?)\}", m.group(1), re.S)
    out["PH_SPEC_BOTTOM"] = np.array(
        [[float(v) for v in r.replace("\n", " ").split(",") if v.strip()] for r in rows]
    )
    assert out["PH_SPEC_BOTTOM"].shape == (8, 41)
    return out


_T = _parse(_HEADER)
LAMBDA = 400.0 + 10.0 * np.arange(41)
AW, BBW, A0, A1, BOTTOM = _T["PH_SPEC_AW"], _T["PH_SPEC_BBW"], _T["PH_SPEC_A0"], _T["PH_SPEC_A1"], _T["PH_SPEC_BOTTOM"]


def at(table: np.ndarray, wavelength_nm: float) -> float:
    """Linear interpolation on the 10 nm grid (simulation use only)."""
    return float(np.interp(wavelength_nm, LAMBDA, table))
