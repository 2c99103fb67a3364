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


def test_image_shard_is_balanced_and_covers_every_view_sample_once():
    V, B = 5, 8
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            (b, e), per_view = sd.image_shard(r, world, V, B)
            assert e - b == V * B // world
            rows = []
            for v, (s0, s1) in sorted(per_view.items()):
                assert 0 <= s0 < s1 <= B
                rows += [v * B + i for i in range(s0, s1)]
            assert rows == list(range(b, e))
            seen += rows
        assert seen == list(range(V * B))
    assert sd.image_shard(1, 8, 5, 8) == ((5, 10), {0: (5, 8), 1: (0, 2)})


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


def _toy_cfg():
    from selfpose3d_b200.config import default_config
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 3
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [64, 96], [16, 24]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE, cfg.MULTI_PERSON.MAX_PEOPLE_NUM, cfg.MULTI_PERSON.THRESHOLD = [8, 8, 4], 2, -1e9
    cfg.PICT_STRUCT.CUBE_SIZE = [8, 8, 8]
    return cfg


def _train_worker(rank, world, port, results):
    """Data-parallel training step: every rank runs the supervised step on its own frame (kernels emulated on CPU),
    then the gradients are averaged with sd.allreduce_gradients."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import bench
    import test_training_cpu
    test_training_cpu.apply_emulation_in_this_process()
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, step = bench.build_training_step(_toy_cfg(), "cpu", [64, 96], 3, seed=5 + rank)
        step()
        sd.allreduce_gradients(model.parameters(), bucket_bytes=1 << 20)      # several buckets
        results[rank] = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_average_world2_gloo():
    import bench
    import test_training_cpu
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_train_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    got0, got1 = results[0], results[1]
    assert set(got0) == set(got1) and all(torch.equal(got0[n], got1[n]) for n in got0)     # replicas agree
    # single-process truth: the mean of the two ranks' own gradients
    import subprocess
    import sys
    import pickle
    code = ("import sys, pickle, torch; sys.path.insert(0, %r); sys.path.insert(0, %r); import bench, test_training_cpu, "
            "test_dist_gloo; test_training_cpu.apply_emulation_in_this_process(); out = []\n"
            "for seed in (5, 6):\n"
            "    m, step = bench.build_training_step(test_dist_gloo._toy_cfg(), 'cpu', [64, 96], 3, seed=seed); step()\n"
            "    out.append({n: p.grad for n, p in m.named_parameters() if p.grad is not None})\n"
            "pickle.dump(out, sys.stdout.buffer)" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                      os.path.dirname(os.path.abspath(__file__))))
    single = pickle.loads(subprocess.run([sys.executable, "-c", code], check=True, capture_output=True).stdout)
    trained = [n for n in got0 if n.split(".")[0] in ("root_net", "pose_net")]
    assert len(trained) >= 100
    top = max(float(((single[0][n] + single[1][n]) / 2).norm()) for n in trained)
    for n in trained:
        want = (single[0][n] + single[1][n]) / 2
        # (thread counts differ between the workers and the single process: summation order, not semantics)
        assert float((got0[n] - want).norm()) <= 1e-3 * max(float(want.norm()), 1e-6 * top), n
