#!/usr/bin/env python
"""One forward step of the bench workload between cudaProfilerStart/Stop, for ncu captures:

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python profiles/run_step.py [--batch B] [--proposals P]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import synthetic  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet_ssv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=bench.BATCH)
ap.add_argument("--proposals", type=int, default=bench.PROPOSALS)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--volume-dtype", default="f32", choices=["bf16", "f32"], help="f32 = the default float32-faithful tensor-core mode (bf16 term pairs)")
a = ap.parse_args()

from selfpose3d_b200 import ops  # noqa: E402
ops.set_volume_dtype(torch.bfloat16 if a.volume_dtype == "bf16" else torch.float32)
bench.PROPOSALS = a.proposals
cfg = bench.make_cfg(a.batch)
model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
model = model.cuda().eval()
meta = synthetic.make_meta(synthetic.ring_cameras(bench.VIEWS, seed=0), a.batch, bench.IMAGE_SIZE)
images = [im.cuda() for im in synthetic.random_images(a.batch, bench.VIEWS, bench.IMAGE_SIZE, seed=0)]
for _ in range(a.warmup):
    model(views1=images, meta1=meta, inference=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(views1=images, meta1=meta, inference=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
