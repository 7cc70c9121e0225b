"""Screen-space partition of one frame over N GPUs (north_star: tile strips + sharded shadow maps).

The reference has no multi-GPU code; the parity oracle of this module is "same image
as one GPU". Pure host logic, no CUDA: tested on CPU with gloo (tests/test_partition_gloo.py).
"""
from __future__ import annotations

TILE = 16


def strips(height: int, world: int) -> list[int]:
    """world+1 row cuts; strip k = screen rows [cuts[k], cuts[k+1]), 16-row aligned except the last."""
    tiles = (height + TILE - 1) // TILE
    cuts = [min(height, ((tiles * k) // world) * TILE) for k in range(world + 1)]
    cuts[-1] = height
    return cuts


def shadow_units(height: int, world: int, casting: list[int]) -> list[tuple[int, int, int, int]]:
    """Work units (light, row0, row1, owner_rank): each casting light's map is cut into
    max(1, world // Ls) row ranges and dealt round-robin, so every rank rasterises the same number
    of shadow texels when Ls divides world (C3: 4 lights x 2 halves on 8 GPUs)."""
    parts = max(1, world // max(1, len(casting)))
    units = []
    for li in casting:
        for p in range(parts):
            units.append((li, (height * p) // parts, (height * (p + 1)) // parts))
    return [(li, a, b, k % world) for k, (li, a, b) in enumerate(units)]


def image_rows(height: int, row0: int, row1: int) -> tuple[int, int]:
    """Screen rows [row0,row1) -> image rows [H-row1, H-row0) (image row r = screen y = H-1-r, buffer.go:225)."""
    return height - row1, height - row0
