"""Horizontal-strip decomposition of the image across GPUs.

The path shards by image rows: scene, BVH and light tables are replicated, pixel indices and RNG streams stay
global (restir.cu:126-127), so a strip computes exactly the pixels the full frame would.  Two neighbour
exchanges per frame carry reservoir rows only (the G-buffer of the halo rows is re-rendered locally):

* E2, between phase A and phase B: post-temporal reservoirs (plane ``resv_temp``) for the spatial radius;
* E1, after phase B: history reservoirs (plane ``resv_history``) for next frame's temporal reprojection.
"""
from __future__ import annotations


def strip_rows(height: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [r0, r1) owned by ``rank``; the first ``height % world`` strips get one extra row."""
    base, extra = divmod(height, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def halo_rows(height: int, world: int, rank: int, halo: int) -> tuple[int, int]:
    """Rows resident on ``rank`` (own strip + halo, clipped to the image)."""
    r0, r1 = strip_rows(height, world, rank)
    return max(0, r0 - halo), min(height, r1 + halo)


def exchange_plan(height: int, world: int, halo: int) -> list[tuple[int, int, int, int]]:
    """All (src, dst, row0, row1) messages of one halo exchange: rows owned by ``src`` that ``dst`` keeps as halo.
    With halo <= strip height only adjacent ranks talk; taller halos reach further ranks."""
    plan = []
    for dst in range(world):
        lo, hi = halo_rows(height, world, dst, halo)
        for src in range(world):
            if src == dst:
                continue
            s0, s1 = strip_rows(height, world, src)
            a, b = max(lo, s0), min(hi, s1)
            if a < b:
                plan.append((src, dst, a, b))
    return plan


def default_halo(spatial_radius: float, temporal_margin: int = 32) -> int:
    """Halo rows: ceil(radius)+1 for the spatial disk (restir.cu:53-55), >= temporal_margin for reprojection."""
    import math

    return max(int(math.ceil(spatial_radius)) + 1, int(temporal_margin))
