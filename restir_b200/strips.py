"""Horizontal-strip decomposition of the image across GPUs.

The path shards by image rows: scene, BVH and light tables are replicated, pixel indices and RNG streams stay
global (restir.cu:126-127), so a strip computes exactly the pixels the full frame would.  A strip keeps ``halo``
rows of its neighbours' planes on each side; ONE neighbour exchange per frame, between phase A and phase B, fills them:

* current G-buffer rows (``geom_cur``, ``matid_cur``) and post-temporal reservoirs (``resv_temp``): read by THIS
  frame's spatial pass within ceil(radius) + 1 rows (restir.cu:53-56);
* the history reservoirs phase A just wrote (``resv_out``; phase B does not touch them, restir.cu:211-212): read by
  the NEXT frame's temporal step at the reprojected row (restir.cu:23), together with the then-previous G-buffer rows.

``halo`` must therefore exceed both ceil(radius) and the largest vertical reprojection distance of any pixel
(``Frame.motion_rows``); a read outside the resident rows is counted (``Frame.halo_miss``) and must stay 0.

Strips are described by ``bounds``: N+1 increasing row indices, rank r owns rows [bounds[r], bounds[r+1]).
Equal-height strips balance badly (sky rows cost almost nothing, ground rows everything), so ``balanced_bounds``
places the cuts by a measured per-row cost (``Frame.row_cost``), refined in a closed loop by ``refine_row_cost``.
"""
from __future__ import annotations

import math

import numpy as np


def uniform_bounds(height: int, world: int) -> list[int]:
    """Equal-height strips; the first ``height % world`` strips get one extra row."""
    base, extra = divmod(height, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < extra else 0))
    return b


def balanced_bounds(row_cost, world: int, min_rows: int = 8) -> list[int]:
    """Cuts that give every rank (about) the same summed row cost; every strip keeps at least ``min_rows`` rows."""
    cost = np.asarray(row_cost, np.float64)
    H = int(cost.shape[0])
    if world * min_rows > H:
        return uniform_bounds(H, world)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    total = cum[-1]
    b = [0]
    for r in range(1, world):
        target = total * r / world
        cut = int(np.searchsorted(cum, target))
        cut = max(cut, b[-1] + min_rows)
        cut = min(cut, H - (world - r) * min_rows)
        b.append(cut)
    b.append(H)
    return b


def refine_row_cost(row_cost, bounds, measured, damping: float = 1.0):
    """One step of the closed loop that places the strip cuts: the row costs inside strip r are rescaled so that they sum
    to the time rank r was MEASURED to take with the current cuts (the model keeps its shape inside a strip, its level
    comes from the measurement).  ``damping`` < 1 applies only that power of the correction: a strip whose cost sits in a
    few of its rows (600 rows of sky above 40 rows at the horizon) otherwise overshoots from one side of the balance to
    the other.  Returns a new array; cut again with :func:`balanced_bounds`."""
    cost = np.array(row_cost, np.float64)
    cost *= float(np.sum(measured)) / max(float(cost.sum()), 1e-300)          # same units as the measurement
    for r in range(len(bounds) - 1):
        seg = slice(bounds[r], bounds[r + 1])
        total = cost[seg].sum()
        if total > 0:
            cost[seg] *= (float(measured[r]) / total) ** damping
        else:
            cost[seg] = float(measured[r]) / max(bounds[r + 1] - bounds[r], 1)
    return cost


def row_cost_from_matid(matid, width: int, shaded_weight: float = 1.0, base_weight: float = 0.03):
    """Per-row cost estimate from a G-buffer material-id plane: shaded pixels (id >= 0) run the candidate loop and two
    more rays, the others only a primary ray that leaves the scene box."""
    m = np.asarray(matid).reshape(-1, width)
    return (m >= 0).sum(1) * shaded_weight + width * base_weight


def strip_rows(height: int, world: int, rank: int, bounds=None) -> tuple[int, int]:
    b = bounds or uniform_bounds(height, world)
    return b[rank], b[rank + 1]


def halo_rows(height: int, world: int, rank: int, halo: int, bounds=None) -> tuple[int, int]:
    """Rows resident on ``rank`` (own strip + halo, clipped to the image)."""
    r0, r1 = strip_rows(height, world, rank, bounds)
    return max(0, r0 - halo), min(height, r1 + halo)


def exchange_plan(height: int, world: int, halo: int, bounds=None) -> list[tuple[int, int, int, int]]:
    """All (src, dst, row0, row1) messages of one halo exchange: rows owned by ``src`` that ``dst`` keeps as halo.
    With halo <= strip height only adjacent ranks talk; taller halos reach further ranks."""
    plan = []
    for dst in range(world):
        lo, hi = halo_rows(height, world, dst, halo, bounds)
        for src in range(world):
            if src == dst:
                continue
            s0, s1 = strip_rows(height, world, src, bounds)
            a, b = max(lo, s0), min(hi, s1)
            if a < b:
                plan.append((src, dst, a, b))
    return plan


def default_halo(spatial_radius: float, motion_rows: int = 31) -> int:
    """Halo rows: ceil(radius)+1 for the spatial disk (restir.cu:53-55) and motion_rows+1 for the reprojection
    (restir.cu:23; ``motion_rows`` = the measured per-frame bound ``Frame.motion_rows`` over the camera path)."""
    return max(int(math.ceil(spatial_radius)) + 1, int(motion_rows) + 1)
