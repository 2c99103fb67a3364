#!/usr/bin/env python
"""Device-time breakdown of the self-supervised training step of profiles/config5_train.py (one GPU):
  python profiles/ssl_step_breakdown.py [--batch-per-gpu 2] [--proposals 4]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import ops, profiler, synthetic  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet_ssv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch-per-gpu", type=int, default=2)
ap.add_argument("--proposals", type=int, default=4)
ap.add_argument("--cprofile", action="store_true", help="host-side profile of one step (top functions by own time)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = bench.make_cfg(a.batch_per_gpu)
cfg.MULTI_PERSON.MAX_PEOPLE_NUM = a.proposals
cfg.WITH_ATTN, cfg.USE_L1, cfg.TRAIN.L1_EPOCH = True, True, 0
cfg.NETWORK.FREEZE_ROOTNET = True
model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
model = model.to(dev).train()
model.root_net.eval()
sets = synthetic.ssl_training_case(bench.IMAGE_SIZE, bench.HEATMAP_SIZE, cfg.NETWORK.NUM_JOINTS, a.batch_per_gpu, bench.VIEWS,
                                   a.proposals, seed=77, image_seed=40)
(v1, m1, t1), (v2, m2, t2), (v3, m3, t3) = [([v.to(dev) for v in vs], m, [t.to(dev) for t in ts]) for vs, m, ts in sets]


def step():
    model.zero_grad(set_to_none=True)
    _, _, _, losses = model(views1=v1, meta1=m1, targets_2d1=t1, views2=v2, meta2=m2, targets_2d2=t2, views3=v3, meta3=m3,
                            targets_2d3=t3, inference=False, epoch=1)
    sum(losses.values()).backward()


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
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
profiler.enable()
step()
profiler.disable()
torch.cuda.synchronize()
fam = profiler.summary()
total = sum(v["ms"] for v in fam.values())
print("SSL step, %d frames, %d proposals: wall %.1f ms, kernel time %.1f ms in %d launches"
      % (a.batch_per_gpu, a.proposals, wall, total, sum(v["launches"] for v in fam.values())))
HBM_KINDS = ("bn", "layout", "maxpool", "maxpool_bwd", "elementwise", "softargmax", "softargmax_bwd", "unproject", "unproject_bwd")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
    # (work = algorithmic bytes for the streaming kernels, FLOPs for the convolution families)
    rate = ("%8.0f GB/s" % (v["work"] / max(v["ms"], 1e-9) * 1e-6)) if k in HBM_KINDS else ("%8.1f TFLOP/s" % (v["work"] / max(v["ms"], 1e-9) * 1e-9))
    print("  %-22s %9.3f ms %6d launches %s" % (k, v["ms"], v["launches"], rate))
