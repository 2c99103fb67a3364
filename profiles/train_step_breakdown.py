#!/usr/bin/env python
"""Where one supervised training step (bench.py's `train_step` case: 1 frame x 5 views, frozen backbone, root net + pose
net forward and backward) spends its device time, by kernel family and by convolution layer:
  python profiles/train_step_breakdown.py [--mode bf16x3|simt]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import ops, profiler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="bf16x3")
ap.add_argument("--cprofile", action="store_true", help="host-side profile of one step (top functions by own time)")
a = ap.parse_args()
ops.set_volume_dtype(torch.float32)
ops.set_float32_conv(a.mode)
dev = torch.device("cuda", 0)
_, step = bench.build_training_step(bench.make_cfg(1), dev, bench.IMAGE_SIZE, bench.VIEWS)
step()
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
if a.cprofile:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
profiler.enable()
step()
profiler.disable()
torch.cuda.synchronize()
fam = profiler.summary()
total = sum(v["ms"] for v in fam.values())
print("mode %s: wall %.1f ms per step, kernel time %.1f ms in %d launches" % (a.mode, wall, total, sum(v["launches"] for v in fam.values())))
HBM_KINDS = ("bn", "layout", "maxpool", "maxpool_bwd", "elementwise", "softargmax", "softargmax_bwd", "unproject", "unproject_bwd")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    # (work = algorithmic bytes for the streaming kernels, FLOPs for the convolution families)
    rate = ("%8.0f GB/s" % (v["work"] / max(v["ms"], 1e-9) * 1e-6)) if k in HBM_KINDS else ("%8.1f TFLOP/s" % (v["work"] / max(v["ms"], 1e-9) * 1e-9))
    print("  %-22s %9.3f ms %6d launches %s" % (k, v["ms"], v["launches"], rate))
rows = sorted(profiler.detail_summary().items(), key=lambda kv: -kv[1]["ms"])
for k, v in rows[:25]:
    print("  %-52s %9.3f ms %5d launches %8.1f TFLOP/s" % (k, v["ms"], v["launches"], v["work"] / max(v["ms"], 1e-9) * 1e-9))
