"""World-size-2 `gloo` tests (CPU) of the multi-GPU sharding logic in selfpose3d_b200/dist.py: the view / cube
partitions, the variable-length all-gather and the view broadcast.  The kernels themselves need a GPU; the
2-GPU equality test against the single-GPU result is tests/multi_gpu_check.py (run with torchrun on the box)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from selfpose3d_b200 import dist as sd


def test_partitions_cover_everything_once():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 7, 80):
            covered = []
            for r in range(world):
                b, e = sd.shard_slice(n, r, world)
                covered += list(range(b, e))
            assert covered == list(range(n))
        views = []
        for r in range(world):
            b, e = sd.view_range(r, world, 5)
            views += list(range(b, e))
            for v in range(b, e):
                assert sd.view_owner(v, world, 5) == r
        assert views == list(range(5))
    assert sd.view_range(0, 2, 5) == (0, 3) and sd.view_range(1, 2, 5) == (3, 5)
    assert sd.view_range(7, 8, 5) == (5, 5)      # more ranks than views: idle rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # variable-length all-gather of per-rank result rows (the [n, J, 3] joints exchange)
        n = 7
        counts = [sd.shard_slice(n, r, world)[1] - sd.shard_slice(n, r, world)[0] for r in range(world)]
        b, e = sd.shard_slice(n, rank, world)
        full = torch.arange(n * 2 * 3, dtype=torch.float32).reshape(n, 2, 3)
        got = sd.all_gather_rows(full[b:e].clone(), counts)
        ok1 = torch.equal(got, full)
        # view broadcast: every rank ends with all views, each from its owner
        V, shape = 5, (2, 3, 4, 5)
        v0, v1 = sd.view_range(rank, world, V)
        local = {v: torch.full(shape, float(v + 1)) for v in range(v0, v1)}
        views = sd.broadcast_views(local, V, shape, torch.device("cpu"))
        ok2 = all(torch.equal(views[v], torch.full(shape, float(v + 1))) for v in range(V))
        # the voxel-grid exchange is a plain sum all-reduce of numerators and counts
        part = torch.full((2, 2, 6), float(rank + 1))
        dist.all_reduce(part)
        ok3 = bool((part == sum(range(1, world + 1))).all())
        results[rank] = (ok1, ok2, ok3)
    finally:
        dist.destroy_process_group()


def test_collective_helpers_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: (True, True, True), 1: (True, True, True)}
