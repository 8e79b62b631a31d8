"""Row-band sharding of one scene over the GPUs of a box (one process per GPU, torch.distributed).

Pixels are independent given a read-only input halo: a pixel reads
(n_spatial-1)+(n_smoothing_radius-1) rows either side (model/samodel.c:2971-2989,
model/common.c:240-258), clamped at the image edge. So the data path needs no collective while
inverting; the only exchanges are
  * the halo rows between neighbouring bands when every rank holds just its own rows
    (point-to-point NCCL send/recv over NVLink), and
  * the final gather of the output planes to rank 0.
Load balance (round 2): the bands are EQUAL row counts and the devices share the work at run time. Every rank's band
lives in a library allocation that the other ranks map (CUDA IPC over NVLink peer access, `BandGroup` below); a solve
kernel that runs out of its own pixels takes pixels from its neighbours' queues, reads their neighbourhoods from the
owner's planes and stores the results into the owner's planes -- pixel by pixel, no collective and no cost model on the
data path. The cost-balanced contiguous plan (`plan_row_bands`, `row_cost_from_prior`: a shallow-water pixel with all
NBOTTOMS substrates costs ~3x a sand-only one) remains for boxes without peer access.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class BandGroup:
    """This rank's Band (photic_b200.samodel.Band) plus the handles of every other rank's band, exchanged once.
    step() = prepare -> [all ranks prepared] -> solve with work sharing -> [all ranks done]; the two synchronisation
    points are stream-ordered one-element all-reduces under NCCL (no host round trip) and barriers under gloo."""

    def __init__(self, inverter, desc, row_begin: int, row_end: int, rank: int, world: int, group=None,
                 scene_planes: bool = False, share: bool = True):
        from .samodel import Band
        self.rank, self.world, self.group = rank, world, group
        self.band = Band(inverter, desc, row_begin, row_end, scene_planes)
        self.peers: list[bytes] = []
        self.shared = False
        dev = self.band.planes.device
        self._nccl = world > 1 and dist.get_backend(group) == "nccl"
        self._flag = torch.zeros(1, device=dev)
        if world > 1 and share:
            mine = torch.frombuffer(bytearray(self.band.export()), dtype=torch.uint8)
            if self._nccl:
                mine = mine.to(dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            hs = [bytes(p.cpu().numpy().tobytes()) for p in parts]
            self.peers = [hs[(rank + q) % world] for q in range(1, world)]  # take work from the next rank first
            self.shared = True

    def sync(self):
        if self.world == 1:
            return
        if self._nccl:  # stream-ordered: the kernels queued after it start once every rank has reached it
            dist.all_reduce(self._flag, group=self.group)
        else:
            torch.cuda.synchronize(self._flag.device)
            dist.barrier(group=self.group)

    def step(self, stream=None) -> dict:
        """Inverts the band (its rasters must have been written into band.planes / band.prior). Returns this
        DEVICE's stats: with work sharing `n_valid` counts the pixels it inverted, its own or its neighbours'."""
        from . import capi
        self.band.prepare(stream)
        self.sync()
        try:
            st = self.band.solve(self.peers, stream)
        except capi.PhoticError as e:  # no peer access on this box: every rank works on its own band only
            if getattr(e, "code", None) != capi.PHB_ENOPEER:
                raise
            self.peers, self.shared = [], False
            st = self.band.solve([], stream)
        self.sync()
        st["shared"] = self.shared
        return st

    def close(self):
        self.band.close()


def equal_row_bands(nrows: int, world: int) -> list[tuple[int, int]]:
    return [(nrows * k // world, nrows * (k + 1) // world) for k in range(world)]


def halo_rows(n_spatial: int, n_smoothing_radius: int) -> int:
    nsp = 1 if n_spatial == 0 else n_spatial  # samodel.c:2967
    return (nsp - 1) + (n_smoothing_radius - 1)


def plan_row_bands(row_cost, world: int) -> list[tuple[int, int]]:
    """Contiguous [r0, r1) per rank with near-equal cumulative cost; every rank gets >= 0 rows."""
    c = np.asarray(row_cost, dtype=np.float64)
    nrows = len(c)
    total = c.sum()
    if total <= 0:
        edges = np.linspace(0, nrows, world + 1).round().astype(int)
    else:
        cum = np.concatenate([[0.0], np.cumsum(c)])
        edges = [0]
        for k in range(1, world):
            edges.append(int(np.searchsorted(cum, total * k / world, side="left")))
        edges.append(nrows)
        edges = np.maximum.accumulate(np.clip(np.array(edges), 0, nrows))
    return [(int(edges[k]), int(edges[k + 1])) for k in range(world)]


def window(r0: int, r1: int, halo: int, nrows: int) -> tuple[int, int, int, int]:
    """Rows a rank must hold to invert [r0, r1): returns (w0, w1, local_begin, local_end)."""
    w0, w1 = max(0, r0 - halo), min(nrows, r1 + halo)
    return w0, w1, r0 - w0, r1 - w0


def row_cost(valid: torch.Tensor, shallow: torch.Tensor, shallow_weight: float = 3.0) -> torch.Tensor:
    """Estimated work per row from the validity and shallow-class masks ([rows, cols] bool)."""
    v = valid.to(torch.float32)
    return (v * (1.0 + (shallow_weight - 1.0) * shallow.to(torch.float32))).sum(dim=1)


# Relative cost of one inversion as a function of the DEPTHS prior |h| (metres): mean evaluation count per depth
# bin measured with the CPU oracle on Exmouth- and Pilbara-shaped rasters (1600 px each, tools notes in DESIGN.md
# section 7) times the per-evaluation cost of the class (all substrates below 8 m: 1.3x a sand-only evaluation,
# tools/class_cost.py). Normalised to the 8-12 m sand-only bin.
_COST_EDGES = (2.0, 4.0, 6.0, 8.0, 12.0, 16.0, 24.0, 32.0)
_COST_WEIGHT = (2.4, 2.9, 2.6, 2.2, 1.0, 1.1, 1.05, 1.35, 1.4)
_COST_NO_PRIOR = 8.0 * 2.4  # eight H starts, all substrates


def row_cost_from_prior(valid: torch.Tensor, prior: torch.Tensor, prior_nodata: float | None = None) -> torch.Tensor:
    """Estimated work per row from the validity mask and the DEPTHS prior plane ([rows, cols]), with the rules of the
    C planner (photic_b200.cu:row_costs): a prior above -1 m is inverted from 1 m (samodel.c:963-967); a pixel whose
    prior is nodata is inverted without one -- eight depth starts with all substrates (samodel.c:2222-2241)."""
    h = torch.where(prior > -1.0, torch.ones_like(prior), prior.abs()).to(torch.float32)
    edges = torch.tensor(_COST_EDGES, dtype=torch.float32, device=prior.device)
    wts = torch.tensor(_COST_WEIGHT, dtype=torch.float32, device=prior.device)
    c = wts[torch.bucketize(h, edges, right=True)]
    if prior_nodata is not None:
        # approx_equal(prior, nodata, 1e-6), common.c:392 (float difference, products in double)
        d = (prior - prior_nodata).abs().to(torch.float64)
        big = torch.maximum(prior.abs().to(torch.float64), torch.full_like(d, abs(float(prior_nodata))))
        c = torch.where(d <= big * 1.0e-6, torch.full_like(c, _COST_NO_PRIOR), c)
    return (c * valid.to(torch.float32)).sum(dim=1)


def exchange_halo(band: torch.Tensor, plan, halo: int, rank: int, world: int, group=None) -> torch.Tensor:
    """band: [planes, r1-r0, cols] rows owned by this rank. Returns the band with up to `halo` rows of the
    neighbouring bands attached above and below (fewer at the image edge). Empty bands are skipped over."""
    if halo == 0 or world == 1:
        return band
    nrows = plan[-1][1]
    r0, r1 = plan[rank]
    if r1 <= r0:
        return band
    w0, w1, _, _ = window(r0, r1, halo, nrows)
    ops, recv = [], {}
    # rows [w0, r0) come from lower ranks, rows [r1, w1) from higher ranks; symmetric sends.
    for other in range(world):
        if other == rank:
            continue
        o0, o1 = plan[other]
        if o1 <= o0:
            continue
        ow0, ow1, _, _ = window(o0, o1, halo, nrows)
        # what I need from `other`
        lo, hi = max(w0, o0), min(w1, o1)
        need = [(a, b) for a, b in ((lo, min(hi, r0)), (max(lo, r1), hi)) if b > a]
        for a, b in need:
            buf = torch.empty((band.shape[0], b - a, band.shape[2]), dtype=band.dtype, device=band.device)
            recv[(a, b)] = buf
            ops.append(dist.P2POp(dist.irecv, buf, other, group))
        # what `other` needs from me
        lo, hi = max(ow0, r0), min(ow1, r1)
        give = [(a, b) for a, b in ((lo, min(hi, o0)), (max(lo, o1), hi)) if b > a]
        for a, b in give:
            ops.append(dist.P2POp(dist.isend, band[:, a - r0:b - r0].contiguous(), other, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    top = [recv[k] for k in sorted(recv) if k[1] <= r0]
    bot = [recv[k] for k in sorted(recv) if k[0] >= r1]
    return torch.cat(top + [band] + bot, dim=1).contiguous()


def repartition_rows(band: torch.Tensor, plan_old, plan_new, rank: int, world: int, group=None) -> torch.Tensor:
    """Moves rows between ranks so that a raster held as row bands `plan_old` is held as `plan_new`
    (both lists of contiguous [r0, r1) covering the same rows). band: [planes, rows_old, cols].
    Point-to-point sends of exactly the rows that change owner (NCCL over NVLink on a B200 box)."""
    if world == 1 or list(plan_old) == list(plan_new):
        return band
    o0, o1 = plan_old[rank]
    n0, n1 = plan_new[rank]
    out = torch.empty((band.shape[0], n1 - n0, band.shape[2]), dtype=band.dtype, device=band.device)
    a, b = max(o0, n0), min(o1, n1)
    if b > a:  # rows that stay
        out[:, a - n0:b - n0] = band[:, a - o0:b - o0]
    ops = []
    for other in range(world):
        if other == rank:
            continue
        p0, p1 = plan_old[other]
        q0, q1 = plan_new[other]
        a, b = max(o0, q0), min(o1, q1)  # mine -> other's new band
        if b > a:
            ops.append(dist.P2POp(dist.isend, band[:, a - o0:b - o0].contiguous(), other, group))
        a, b = max(p0, n0), min(p1, n1)  # other's old rows -> my new band
        if b > a:
            buf = torch.empty((band.shape[0], b - a, band.shape[2]), dtype=band.dtype, device=band.device)
            ops.append(dist.P2POp(dist.irecv, buf, other, group))
            ops.append(("copy", buf, a - n0, b - n0))
    reqs = dist.batch_isend_irecv([o for o in ops if not isinstance(o, tuple)]) if any(not isinstance(o, tuple) for o in ops) else []
    for r in reqs:
        r.wait()
    for o in ops:
        if isinstance(o, tuple):
            _, buf, x0, x1 = o
            out[:, x0:x1] = buf
    return out


def gather_bands(local: torch.Tensor, plan, rank: int, world: int, dst: int = 0, group=None):
    """local: [..., r1-r0, cols] rows owned by this rank -> full [..., nrows, cols] on rank dst (None elsewhere).
    Bands differ in height, so they are padded to the tallest band for one all_gather."""
    if world == 1:
        return local
    hmax = max(b - a for a, b in plan)
    lead, cols = local.shape[:-2], local.shape[-1]
    pad = torch.zeros((*lead, hmax, cols), dtype=local.dtype, device=local.device)
    pad[..., : local.shape[-2], :] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    if rank != dst:
        return None
    return torch.cat([parts[k][..., : plan[k][1] - plan[k][0], :] for k in range(world)], dim=-2)


def allreduce_minmax(lo: float, hi: float, device, group=None) -> tuple[float, float]:
    """REFINE without CLIP needs the grid's global min/max (model/refine.c:215-225)."""
    t = torch.tensor([lo, -hi], dtype=torch.float32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t[0]), float(-t[1])
