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
