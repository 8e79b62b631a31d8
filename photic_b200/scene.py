"""Deterministic synthetic Landsat-8-shaped multi-date scenes.

The reference ships no data (README.md:26-88 are Google-Drive links), so parity and throughput
are measured on synthetic scenes with the shapes BASELINE.json names. A scene is a stack of
float32 reflectance planes ``[n_dates*4][rows][cols]`` (coastal, blue, green, red per date),
simulated with the same Lee/HOPE forward model the inversion fits (model/samodel.c:2846-2949),
plus a DEPTHS prior plane (negative-down metres, model/samodel.c:960-967).

Everything is a pure function of the GLOBAL pixel coordinate and the seed (counter-based
splitmix64 hash, integer ops only), so any row window of a scene can be generated on any rank,
on CPU or GPU, and is the same scene. torch is used only as an array library here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace

import numpy as np
import torch

from . import spectra

NODATA = -9999.0


@dataclass(frozen=True)
class SceneSpec:
    name: str
    nrows: int
    ncols: int
    n_dates: int
    seed: int
    wavelengths: tuple = (443, 482, 561, 655)  # int nm, as scene.wavelengths (common.h:200)
    land_fraction: float = 0.45
    noise: float = 0.01           # multiplicative U(+-noise)
    outlier_fraction: float = 0.0  # share of samples hit by +-outlier_amp
    outlier_amp: float = 0.30
    depth_mode: str = "ramp"      # "ramp" | "mixed" (50/50 deep/shallow patches: divergence stress)
    n_smoothing_radius: int = 1   # SET NSMOOTH  (bam.c:5628)
    n_spatial: int = 2            # SET NSPATIAL (bam.c:5636)
    n_bottoms: int = 3            # SET NBOTTOMS (bam.c:5644)
    theta_view: float = 0.0
    r_sigma: float = 1.0e-4

    @property
    def n_bands(self) -> int:
        return len(self.wavelengths)

    @property
    def n_planes(self) -> int:
        return self.n_dates * self.n_bands

    def theta_sun(self, s: int) -> float:
        return 25.0 + 3.0 * s

    def h_tide(self, s: int) -> float:
        return 0.1 * s

    def scaled(self, nrows: int, ncols: int, name: str | None = None) -> "SceneSpec":
        """Same physics and seed on a smaller raster (used by parity tests)."""
        return replace(self, nrows=nrows, ncols=ncols, name=name or f"{self.name}-{nrows}x{ncols}")


# The five BASELINE.json configurations (shapes from the reference's README.md:26-88).
CONFIGS = {
    "murion": SceneSpec("murion", 1040, 305, 4, seed=0x5EED0001),
    "exmouth": SceneSpec("exmouth", 3930, 2858, 6, seed=0x5EED0002),
    "abudhabi": SceneSpec("abudhabi", 3370, 4700, 8, seed=0x5EED0003),
    "qatar": SceneSpec("qatar", 7361, 7344, 8, seed=0x5EED0004, noise=0.05, outlier_fraction=0.01,
                       depth_mode="mixed"),
    "pilbara": SceneSpec("pilbara", 12662, 16077, 8, seed=0x5EED0005),
}


def _i64(c: int) -> int:
    c &= (1 << 64) - 1
    return c - (1 << 64) if c >= (1 << 63) else c


_C1, _C2, _C3 = _i64(0x9E3779B97F4A7C15), _i64(0xBF58476D1CE4E5B9), _i64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    return (z >> k) & ((1 << (64 - k)) - 1)


def hash01(idx: torch.Tensor, stream: int) -> torch.Tensor:
    """splitmix64 of (idx, stream) -> float64 uniform in [0, 1). idx: int64 tensor."""
    z = idx + _i64((stream + 1) * 0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _C2
    z = (z ^ _lsr(z, 27)) * _C3
    z = z ^ _lsr(z, 31)
    return _lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def _scalar01(spec: SceneSpec, tag: int) -> float:
    return float(hash01(torch.tensor([tag], dtype=torch.int64), spec.seed ^ 0xABCDEF)[0])


def date_params(spec: SceneSpec, s: int) -> tuple[float, float, float]:
    """True P, G, X of acquisition date s (paper example magnitudes, photic.tex:816-818)."""
    return (0.03 + 0.05 * _scalar01(spec, 3 * s), 0.04 + 0.06 * _scalar01(spec, 3 * s + 1),
            0.005 + 0.02 * _scalar01(spec, 3 * s + 2))


def _land_field(spec: SceneSpec, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    ph = [2 * math.pi * _scalar01(spec, 100 + k) for k in range(6)]
    return (torch.sin(2 * math.pi * (1.3 * u + 0.4 * v) + ph[0]) + 0.8 * torch.sin(2 * math.pi * (0.6 * u - 1.7 * v) + ph[1])
            + 0.6 * torch.sin(2 * math.pi * (3.1 * u + 2.3 * v) + ph[2]) + 0.5 * torch.sin(2 * math.pi * (5.2 * v - 4.1 * u) + ph[3])
            + 0.3 * torch.sin(2 * math.pi * (9.0 * u + 7.0 * v) + ph[4]) + 0.25 * torch.sin(2 * math.pi * (13.0 * u - 11.0 * v) + ph[5]))


def _land_threshold(spec: SceneSpec) -> float:
    g = torch.linspace(0.0, 1.0, 257, dtype=torch.float64)[:-1] + 0.5 / 256
    f = _land_field(spec, g[:, None].expand(256, 256), g[None, :].expand(256, 256)).reshape(-1)
    k = max(1, min(f.numel(), int(round((1.0 - spec.land_fraction) * f.numel()))))
    return float(torch.kthvalue(f, k).values)


def truth_fields(spec: SceneSpec, row0: int, row1: int, device="cpu") -> dict:
    """True depth / bottom fields and land mask for global rows [row0, row1)."""
    i = torch.arange(row0, row1, dtype=torch.float64, device=device)[:, None]
    j = torch.arange(0, spec.ncols, dtype=torch.float64, device=device)[None, :]
    u = ((i + 0.5) / spec.nrows).expand(row1 - row0, spec.ncols)
    v = ((j + 0.5) / spec.ncols).expand(row1 - row0, spec.ncols)
    two_pi = 2 * math.pi
    if spec.depth_mode == "mixed":
        bi = torch.div(torch.arange(row0, row1, device=device), 48, rounding_mode="floor")[:, None]
        bj = torch.div(torch.arange(0, spec.ncols, device=device), 48, rounding_mode="floor")[None, :]
        deep = hash01((bi * 100003 + bj).to(torch.int64), spec.seed ^ 0x77) < 0.5
        w = 0.5 + 0.25 * torch.sin(two_pi * (11.0 * u + 3.0 * v)) + 0.25 * torch.cos(two_pi * (5.0 * v - 7.0 * u))
        H = torch.where(deep, 20.0 + 20.0 * w, 0.5 + 7.4 * w)
    else:
        base = (0.55 * v + 0.15 * u + 0.15 * torch.sin(two_pi * 3.0 * u) * torch.cos(two_pi * 2.0 * v)
                + 0.10 * torch.sin(two_pi * (7.0 * u + 5.0 * v)) + 0.15)
        H = 0.5 + 39.5 * base.clamp(0.0, 1.0) ** 1.5
    B = [0.35 + 0.10 * torch.sin(two_pi * (2.0 * u + 1.0 * v)),
         0.10 + 0.05 * torch.sin(two_pi * (3.0 * v - 1.0 * u) + 1.0),
         0.15 + 0.05 * torch.cos(two_pi * (4.0 * u + 2.0 * v) + 2.0)]
    w = [1.0 + 0.9 * torch.sin(two_pi * (1.5 * u + 2.5 * v) + 0.3),
         1.0 + 0.9 * torch.sin(two_pi * (2.5 * u - 1.5 * v) + 2.1),
         1.0 + 0.9 * torch.cos(two_pi * (3.5 * u + 0.5 * v) + 4.0)]
    wsum = w[0] + w[1] + w[2]
    q = [wk / wsum for wk in w]
    land = _land_field(spec, u, v) > _land_threshold(spec)
    return {"H": H, "B": B, "q": q, "land": land, "u": u, "v": v}


def forward_rrs(spec: SceneSpec, s: int, b: int, H, B, q, P, G, X):
    """Lee/HOPE above-surface Rrs for date s, band b (same equations as model/samodel.c:2873-2944,
    with the fixed Rrs440/Rrs490 = 0.0085/0.0098 of SURVEY 8d for the backscatter slope Y)."""
    lam = float(spec.wavelengths[b])
    a0, a1 = spectra.at(spectra.A0, lam), spectra.at(spectra.A1, lam)
    aw, bbw = spectra.at(spectra.AW, lam), spectra.at(spectra.BBW, lam)
    rho = sum(q[k] * B[k] * spectra.at(spectra.BOTTOM[k], lam) for k in range(3))
    a = aw + (a0 + a1 * torch.log(P)) * P + G * math.exp(-0.015 * (lam - 440.0))
    Y = min(2.5, max(0.0, 3.44 * (1.0 - 3.17 * math.exp(-2.01 * 0.0085 / 0.0098))))
    bb = bbw + X * (440.0 / lam) ** Y
    uu = bb / (a + bb)
    K = (a + bb).clamp(0.0, 2.5)
    sec_v = 1.0 / math.cos(math.radians(spec.theta_view))
    sec_s = 1.0 / math.cos(math.radians(spec.theta_sun(s)))
    rrs_dp = (0.084 + 0.170 * uu) * uu
    DuC = 1.03 * torch.sqrt(1.0 + 2.4 * uu)
    DuB = 1.04 * torch.sqrt(1.0 + 5.4 * uu)
    Hs = H  # the reference adds the tide only to md->depth, never to the optics (samodel.c:2873)
    rrs = rrs_dp * (1.0 - torch.exp(-(sec_s + DuC * sec_v) * K * Hs)) + rho / math.pi * torch.exp(-(sec_s + DuB * sec_v) * K * Hs)
    return 0.5 * rrs / (1.0 - 1.5 * rrs)


def generate(spec: SceneSpec, row0: int = 0, row1: int | None = None, device="cpu", chunk_rows: int = 512):
    """Reflectance planes and DEPTHS prior for global rows [row0, row1).

    Returns (planes float32 [n_planes, rows, ncols], prior float32 [rows, ncols]) on ``device``.
    """
    row1 = spec.nrows if row1 is None else row1
    rows = row1 - row0
    planes = torch.empty((spec.n_planes, rows, spec.ncols), dtype=torch.float32, device=device)
    prior = torch.empty((rows, spec.ncols), dtype=torch.float32, device=device)
    npx = spec.nrows * spec.ncols
    for c0 in range(row0, row1, chunk_rows):
        c1 = min(row1, c0 + chunk_rows)
        t = truth_fields(spec, c0, c1, device)
        H, land, u, v = t["H"], t["land"], t["u"], t["v"]
        pix = (torch.arange(c0, c1, device=device, dtype=torch.int64)[:, None] * spec.ncols
               + torch.arange(0, spec.ncols, device=device, dtype=torch.int64)[None, :])
        for s in range(spec.n_dates):
            P0, G0, X0 = date_params(spec, s)
            mod = 1.0 + 0.1 * torch.sin(2 * math.pi * (1.0 * u + 2.0 * v) + s)
            P, G, X = P0 * mod, G0 * mod, X0 * mod
            for b in range(spec.n_bands):
                g = s * spec.n_bands + b
                R = forward_rrs(spec, s, b, H, t["B"], t["q"], P, G, X)
                R = R * (1.0 + spec.noise * (2.0 * hash01(pix + g * npx, spec.seed) - 1.0))
                if spec.outlier_fraction > 0.0:
                    hit = hash01(pix + g * npx, spec.seed ^ 0x0DD) < spec.outlier_fraction
                    sign = torch.where(hash01(pix + g * npx, spec.seed ^ 0x51C) < 0.5, -1.0, 1.0)
                    R = torch.where(hit, R * (1.0 + spec.outlier_amp * sign), R)
                R = torch.where(land, torch.full_like(R, NODATA), R)
                planes[g, c0 - row0:c1 - row0] = R.to(torch.float32)
        pr = -H * (1.0 + 0.2 * (hash01(pix, spec.seed ^ 0xD3) - 0.5))
        pr = torch.where(land, torch.full_like(pr, NODATA), pr)
        prior[c0 - row0:c1 - row0] = pr.to(torch.float32)
    return planes, prior


def valid_mask(planes: torch.Tensor, nodata: float = NODATA) -> torch.Tensor:
    """Pixels samodel would invert: no band of any date is nodata or negative (samodel.c:933-947)."""
    return ((planes != nodata) & (planes >= 0.0)).all(dim=0)
