#!/usr/bin/env python
"""K1 alone at the bench size: 80 person cubes of 64^3 x 15 channels from 8 frames x 5 views of 96x72 heat-maps.
  python profiles/unproject_bench.py          # prints us per launch and GB/s on the algorithmic bytes"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import ops, synthetic  # noqa: E402
from selfpose3d_b200.models import project_layer  # noqa: E402

dev = "cuda:0"
cfg = bench.make_cfg(8)
B, V, J, P = 8, 5, 15, 10
meta = synthetic.make_meta(synthetic.ring_cameras(V, seed=0), B, bench.IMAGE_SIZE)
people = synthetic.synthetic_people(B, seed=1, num_joints=J)
hms = [h.to(dev) for h in synthetic.render_heatmaps(people, meta, bench.IMAGE_SIZE, bench.HEATMAP_SIZE, num_joints=J)]
cams = ops.pack_cameras(meta, bench.IMAGE_SIZE).to(dev)
g = torch.Generator().manual_seed(0)
centers = torch.cat([(torch.rand(B * P, 2, generator=g) - 0.5) * 6000, 800 + torch.rand(B * P, 1, generator=g) * 400,
                     torch.zeros(B * P, 1), torch.ones(B * P, 1)], 1).to(dev)
centers[:, 1] -= 500
sample = torch.arange(B * P, device=dev, dtype=torch.int32) // P
layer = project_layer.ProjectLayer(cfg)
alg_bytes = B * P * J * 64 ** 3 * 4 + B * V * J * 96 * 72 * 4    # float32 cubes written once + maps read once


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


hms_cl, st = project_layer._common_strides(hms)
f16 = ops.heatmaps_to_f16(hms_cl, st, J)
for name, fn in [
    ("float32 form, f32 cubes (channel-last)", lambda: layer.project_cl(hms, cams, centers, False, [2000.0] * 3, [64] * 3,
                                                                        cube_sample=sample)),
    ("float32 form, bf16 cubes pitch 20->bf16", lambda: layer.project_cl(hms, cams, centers, False, [2000.0] * 3, [64] * 3,
                                                                          cube_sample=sample, dtype=torch.bfloat16, c_pitch=32)),
    ("throughput form (fp16 maps, bf16 cubes)", lambda: layer.project_cl(hms, cams, centers, False, [2000.0] * 3, [64] * 3,
                                                                           cube_sample=sample, dtype=torch.bfloat16, c_pitch=16,
                                                                           hms_f16=f16)),
    ("fp16 map conversion", lambda: ops.heatmaps_to_f16(hms_cl, st, J)),
]:
    us = timed(fn)
    print("%-44s %9.1f us   %7.1f GB/s on %d algorithmic bytes" % (name, us, alg_bytes / us * 1e-3, alg_bytes))
