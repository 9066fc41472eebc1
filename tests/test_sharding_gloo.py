"""CPU, world_size 2 over gloo: the N > 1 host logic — stream partitioning, bench-style max-over-ranks timing reduction and
the gather of steered-response maps (the path's only collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from beamform_b200.shard import gather_maps, shard_streams


def test_shard_streams_partitions_exactly():
    for n in (0, 1, 7, 8, 1184, 1025):
        for w in (1, 2, 4, 8):
            blocks = [shard_streams(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in blocks]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        shard_streams(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_streams(n_streams, world, rank)
    full = torch.arange(n_streams * 3 * 5, dtype=torch.float32).reshape(n_streams, 3, 5)   # maps[s][t][d]
    got = gather_maps(full[b:e].clone(), n_streams)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)      # bench.py: elapsed time = max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret[rank] = (bool(torch.equal(got, full)), float(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [5, 8])
def test_gather_maps_world2_gloo(n_streams):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_streams, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), "every rank must hold the full, correctly ordered map tensor"
    assert all(ret[r][1] == 11.0 for r in range(world))
