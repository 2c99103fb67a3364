#!/usr/bin/env python
"""BASELINE configs[3]: ONE batch of 8 frames strong-scaled over N GPUs -- backbone sharded by view, root grid
exchanged with one NCCL all-reduce of numerators + counts, person cubes sharded by (sample, proposal), heat-maps
broadcast, joints all-gathered (selfpose3d_b200/dist.py).  Run under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
      profiles/config4_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import dist as sd, ops, synthetic  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet_ssv  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ops.set_volume_dtype(torch.bfloat16)
cfg = bench.make_cfg(bench.BATCH)
model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
model = model.to(dev).eval()
meta = synthetic.make_meta(synthetic.ring_cameras(bench.VIEWS, seed=0), bench.BATCH, bench.IMAGE_SIZE)
images = synthetic.random_images(bench.BATCH, bench.VIEWS, bench.IMAGE_SIZE, seed=0)
v0, v1 = sd.view_range(rank, world, bench.VIEWS)
mine = {v: images[v].to(dev) for v in range(v0, v1)}


def step():
    return sd.infer_view_sharded(model, mine, meta)[0]


for _ in range(3):
    step()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    step()
e1.record()
dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("config 4 (one batch of %d frames over %d GPUs, views + cubes sharded): %.2f ms per batch = %.1f frames/s"
          % (bench.BATCH, world, float(t), bench.BATCH / float(t) * 1e3))
dist.destroy_process_group()
