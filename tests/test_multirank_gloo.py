"""CPU suite, part 3: the N > 1 host path with world_size-2/3 `gloo` process groups.

Each rank owns the interleaved row tiles the CUDA kernels would render, fills its (padded, equal-size) tile buffer with a
recognisable pattern, all-gathers, and de-interleaves with the same index math the device kernel uses
(rtiow_b200/partition.py mirrors csrc).  The frame must come back complete and in top-down order on every rank, for
tile counts that do not divide evenly (H=675, T=4, G=2/3 as at BASELINE's 1200x675).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rtiow_b200 import partition as pr


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, W, H, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = pr.rows_of_rank(H, T, world, rank)
        per = pr.max_rows_per_rank(H, T, world)
        tiles = torch.full((per, W), 0xFFFFFFFF, dtype=torch.int64)                 # padding rows stay 0xFFFFFFFF
        for lr, y in enumerate(rows):                                             # rank-local top-down order
            tiles[lr] = torch.arange(W, dtype=torch.int64) + y * W                  # "pixel" = its global linear index
        gathered = torch.empty((world * per, W), dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, tiles)                              # the one collective of the path
        frame = gathered[torch.from_numpy(pr.gather_index(W, H, T, world))]
        ok = bool((frame.flatten() == torch.arange(W * H)).all())
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, float(t.item()), len(rows)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,T", [(2, 40, 675, 4), (3, 16, 225, 7), (2, 8, 2, 64)])
def test_tiles_allgather_deinterleave(world, W, H, T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, T, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _, _ in res) and all(mx == world for _, _, mx, _ in res)
    assert sum(n for *_, n in res) == H


def test_partition_properties():
    for (H, T, G) in [(675, 4, 8), (2160, 8, 8), (225, 1, 3), (133, 7, 4), (2, 64, 8), (675, 4, 1)]:
        rows = [pr.rows_of_rank(H, T, G, r) for r in range(G)]
        assert sorted(sum(rows, [])) == list(range(H))
        assert all(r == sorted(r) for r in rows)
        assert max(len(r) for r in rows) <= pr.max_rows_per_rank(H, T, G)
        for y in range(H):
            r, lr = pr.owner_and_local_row(y, T, G)
            assert rows[r][lr] == y
        # interleaving balances cost: every rank gets within one tile of H/G rows
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= T
