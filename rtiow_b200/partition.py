"""Frame partition for multi-GPU renders: interleaved row tiles.

Rank r of G renders the top-down rows {y : (y // tile_rows) % G == r}; its tile buffer holds those rows in
increasing y ("rank-local row order").  This is host-side index math only, mirroring rows_of_rank /
max_rows_per_rank / local_to_global_row / deinterleave_kernel in rtiow_b200/csrc; the CUDA side does the work.
"""
from __future__ import annotations

import numpy as np


def rows_of_rank(height: int, tile_rows: int, world: int, rank: int) -> list[int]:
    tile_rows = min(tile_rows, height)            # a tile taller than the frame is the whole frame (capi.cu: tile_rows_of)
    n_tiles = -(-height // tile_rows)
    rows = []
    for tg in range(rank, n_tiles, world):
        rows.extend(range(tg * tile_rows, min((tg + 1) * tile_rows, height)))
    return rows


def max_rows_per_rank(height: int, tile_rows: int, world: int) -> int:
    tile_rows = min(tile_rows, height)
    n_tiles = -(-height // tile_rows)
    return -(-n_tiles // world) * tile_rows


def tile_buffer_bytes(width: int, height: int, tile_rows: int, world: int) -> int:
    return max_rows_per_rank(height, tile_rows, world) * width * 4


def owner_and_local_row(y: int, tile_rows: int, world: int) -> tuple[int, int]:
    """tile_rows: already clamped to the frame height by the caller"""
    tg, within = divmod(y, tile_rows)
    return tg % world, (tg // world) * tile_rows + within


def gather_index(width: int, height: int, tile_rows: int, world: int) -> np.ndarray:
    """frame_row[y] = gathered_rows[index[y]] where gathered = the G tile buffers concatenated in rank order,
    viewed as rows of `width` pixels (what an all-gather of equal-size tile buffers leaves on every rank)."""
    tile_rows = min(tile_rows, height)
    per = max_rows_per_rank(height, tile_rows, world)
    idx = np.empty(height, np.int64)
    for y in range(height):
        r, lr = owner_and_local_row(y, tile_rows, world)
        idx[y] = r * per + lr
    return idx
