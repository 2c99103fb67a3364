#!/usr/bin/env python
"""Per-layer device time of one bench step (CUDA events around every C-ABI launch):
  python profiles/layer_times.py [--volume-dtype bf16|f32|f32x3|f32x6] [--batch B] [--proposals P]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import ops, profiler, synthetic  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet_ssv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=bench.BATCH)
ap.add_argument("--proposals", type=int, default=bench.PROPOSALS)
ap.add_argument("--volume-dtype", default="bf16")
a = ap.parse_args()
ops.set_volume_dtype(torch.bfloat16 if a.volume_dtype == "bf16" else torch.float32)
if a.volume_dtype in ("f32x3", "f32x6"):    # float32 activations on the tcgen05 kernel through bf16 operand splitting
    ops.set_float32_conv("bf16x3" if a.volume_dtype == "f32x3" else "bf16x6")
bench.PROPOSALS = a.proposals
cfg = bench.make_cfg(a.batch)
model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
model = model.cuda().eval()
meta = synthetic.make_meta(synthetic.ring_cameras(bench.VIEWS, seed=0), a.batch, bench.IMAGE_SIZE)
images = [im.cuda() for im in synthetic.random_images(a.batch, bench.VIEWS, bench.IMAGE_SIZE, seed=0)]
for _ in range(2):
    model(views1=images, meta1=meta, inference=True)
torch.cuda.synchronize()
profiler.enable()
model(views1=images, meta1=meta, inference=True)
profiler.disable()
torch.cuda.synchronize()
rows = sorted(profiler.detail_summary().items(), key=lambda kv: -kv[1]["ms"])
total = sum(v["ms"] for v in profiler.summary().values())
print("total kernel ms per step: %.2f" % total)
for k, v in profiler.summary().items():
    print("  %-20s %9.3f ms  %5d launches" % (k, v["ms"], v["launches"]))
print("%-48s %9s %8s %9s" % ("layer", "ms", "launches", "TFLOP/s"))
for k, v in rows[:60]:
    print("%-48s %9.3f %8d %9.1f" % (k, v["ms"], v["launches"], v["work"] / (v["ms"] * 1e-3) / 1e12))
