"""World-size-2 gloo test (CPU) of the N>1 host logic: IC sharding with no data-path
collective, ragged blocks, and the result gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_plasma_flux_b200.ensemble import gather_states, shard_range


def test_shard_range_partitions():
    for n, world in [(4096, 8), (65536, 8), (10, 4), (3, 8), (7, 2)]:
        blocks = [shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_ics, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_ics * 3 * 8, dtype=torch.float32).reshape(n_ics, 3, 8)
        a, b = shard_range(n_ics, rank, world)
        local = full[a:b] * 2.0 + 1.0                 # stand-in for "advance my block"; no communication
        got = gather_states(local, n_ics)
        ok = torch.equal(got, full * 2.0 + 1.0)
        t = torch.tensor([float(b - a)])
        dist.all_reduce(t)                            # every IC owned exactly once
        out[rank] = bool(ok and int(t.item()) == n_ics)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for n_ics in (8, 7):
        out = mp.get_context("spawn").Manager().dict()
        mp.spawn(_worker, args=(2, port, n_ics, out), nprocs=2, join=True)
        assert out[0] and out[1]
        port += 1
