"""Screen-space partition of one frame over N GPUs (north_star: tile strips + sharded shadow maps).

The reference has no multi-GPU code; the parity oracle of this module is "same image
as one GPU". Pure host logic, no CUDA: tested on CPU with gloo (tests/test_partition_gloo.py).
"""
from __future__ import annotations

TILE = 16


def strips(height: int, world: int) -> tuple[int, list[tuple[int, int]]]:
    """Equal strips of `chunk` IMAGE rows, rank k owning image rows [k*chunk, (k+1)*chunk) (the last may be short),
    so that ONE in-place all-gather over the image buffer assembles the frame on every rank (rank 0 included).
    Returns (chunk, [(screen row0, screen row1)] per rank); image row r = screen y = H-1-r (buffer.go:225)."""
    chunk = (height + world - 1) // world
    out = []
    for k in range(world):
        i0, i1 = min(height, k * chunk), min(height, (k + 1) * chunk)
        out.append((height - i1, height - i0))
    return chunk, out


def shadow_chunks(height: int, world: int, n_cast: int) -> tuple[int, list[list[tuple[int, int, int]]]]:
    """The Ls shadow maps are one stacked [Ls*H][W] array cut into `world` equal chunks of `chunk` stacked rows
    (the last may be short): rank k rasterises exactly its chunk and ONE in-place all-gather completes every
    rank's copy. Returns (chunk, per-rank list of (casting index, row0, row1)). C3 on 8 GPUs: half a light each."""
    total = n_cast * height
    chunk = (total + world - 1) // world if world else 0
    per_rank = []
    for k in range(world):
        a, b = min(total, k * chunk), min(total, (k + 1) * chunk)
        units = []
        while a < b:
            li, r0 = divmod(a, height)
            r1 = min(height, r0 + (b - a))
            units.append((li, r0, r1))
            a += r1 - r0
        per_rank.append(units)
    return chunk, per_rank


def shadow_units(height: int, world: int, casting: list[int]) -> list[tuple[int, int, int, int]]:
    """Flat list of (light id, row0, row1, owner rank) over all ranks (see shadow_chunks)."""
    _, per_rank = shadow_chunks(height, world, len(casting))
    return [(casting[ci], a, b, k) for k, units in enumerate(per_rank) for ci, a, b in units]


def image_rows(height: int, row0: int, row1: int) -> tuple[int, int]:
    """Screen rows [row0,row1) -> image rows [H-row1, H-row0) (image row r = screen y = H-1-r, buffer.go:225)."""
    return height - row1, height - row0


# ---- load balancing (peer-memory frames only: the NCCL exchange needs equal chunks, prc_render_peer takes any rows) ----

def equal_bounds(total: int, parts: int) -> list[int]:
    """parts+1 boundaries of `parts` contiguous ranges of ceil(total/parts) items (the last may be short or empty)."""
    chunk = (total + parts - 1) // parts if parts else 0
    return [min(total, k * chunk) for k in range(parts + 1)]


def balanced_bounds(bounds: list[int], cost: list[float], damping: float = 1.0, min_size: int = 1) -> list[int]:
    """New boundaries of the contiguous ranges [bounds[k], bounds[k+1]) such that every range gets the same share of the
    measured cost, assuming the cost of a range was spread evenly over its items (piecewise-constant density). `cost[k]` is
    what range k took last time (any unit); `damping` < 1 moves only part of the way (the density model ignores per-range
    fixed costs, so repeated application converges from one side instead of overshooting). The result covers the same
    [bounds[0], bounds[-1]), is non-decreasing and keeps at least `min_size` items per range when there is room."""
    n = len(cost)
    assert len(bounds) == n + 1 and n >= 1
    lo, hi = bounds[0], bounds[-1]
    total_cost = float(sum(max(0.0, c) for c in cost))
    if n == 1 or hi - lo <= 0 or total_cost <= 0.0:
        return list(bounds)
    target = total_cost / n
    new = [lo]
    k, done = 0, 0.0  # walking range k; `done` = cost of the ranges before it
    for j in range(1, n):
        want = j * target
        while k < n - 1 and done + max(0.0, cost[k]) < want:
            done += max(0.0, cost[k])
            k += 1
        size, c = bounds[k + 1] - bounds[k], max(0.0, cost[k])
        x = bounds[k] + (size * (want - done) / c if c > 0.0 and size > 0 else 0.0)
        x = bounds[j] + damping * (x - bounds[j])
        new.append(int(round(x)))
    new.append(hi)
    # monotone, at least min_size per range where the total allows it
    m = min_size if (hi - lo) >= n * min_size else 0
    for j in range(1, n):
        new[j] = max(new[j], new[j - 1] + m)
    for j in range(n - 1, 0, -1):
        new[j] = min(new[j], new[j + 1] - m)
    return new


def strips_from_bounds(height: int, bounds: list[int]) -> list[tuple[int, int]]:
    """Image-row boundaries (rank k owns image rows [bounds[k], bounds[k+1]), rank 0 the top) -> per-rank SCREEN rows
    (row0, row1) as prc_frame wants them (image row r = screen y = H-1-r, buffer.go:225)."""
    return [(height - bounds[k + 1], height - bounds[k]) for k in range(len(bounds) - 1)]


def shadow_units_from_bounds(height: int, casting: list[int], bounds: list[int]) -> list[tuple[int, int, int, int]]:
    """Boundaries over the STACKED shadow rows [0, len(casting)*height) -> (light id, row0, row1, owner rank) units; a range
    that spans several lights yields one unit per light (like shadow_chunks)."""
    out = []
    for k in range(len(bounds) - 1):
        a, b = bounds[k], bounds[k + 1]
        while a < b:
            ci, r0 = divmod(a, height)
            r1 = min(height, r0 + (b - a))
            out.append((casting[ci], r0, r1, k))
            a += r1 - r0
    return out
