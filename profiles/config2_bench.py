#!/usr/bin/env python
"""BASELINE configs[1]: 5-view Panoptic-shaped synthetic batch = 4 (network input 960x512, heat-maps 128x240),
80x80x20 grid, CuboidProposalNet only (un-projection of all 15 joints -> V2VNet(15,1) -> NMS/top-10), bf16 mode.
  python profiles/config2_bench.py [--batch 4] [--root-channel]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from selfpose3d_b200 import ops, profiler, synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import cuboid_proposal_net  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--root-channel", action="store_true", help="ROOTNET_ROOTHM: only the root joint's map (C = 1)")
ap.add_argument("--mode", default="f32x3", choices=["f32x3", "bf16"], help="default (float32-faithful) or bf16 throughput mode")
a = ap.parse_args()
dev = "cuda:0"
ops.set_volume_dtype(torch.bfloat16 if a.mode == "bf16" else torch.float32)
cfg = default_config()
cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [960, 512], [240, 128]
cfg.NETWORK.ROOTNET_ROOTHM = bool(a.root_channel)
net = cuboid_proposal_net.CuboidProposalNet(cfg).to(dev).eval()
B, V, J = a.batch, 5, cfg.NETWORK.NUM_JOINTS
meta = synthetic.make_meta(synthetic.ring_cameras(V, seed=0), B, cfg.NETWORK.IMAGE_SIZE)
people = synthetic.synthetic_people(B, seed=1, num_joints=J)
hms = [h.to(dev) for h in synthetic.render_heatmaps(people, meta, cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE,
                                                    num_joints=J)]
for _ in range(3):
    net(hms, meta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    root, gc = net(hms, meta)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
# the same forward as ONE CUDA-graph launch (selfpose3d_b200/graphs.py): the host side of the ~40 launches disappears
from selfpose3d_b200 import graphs  # noqa: E402
graphed = graphs.graphed_proposal_net(net, hms, meta)
for _ in range(3):
    graphed(*hms)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    graphed(*hms)
e1.record()
torch.cuda.synchronize()
ms_graph = e0.elapsed_time(e1) / n
profiler.enable()
net(hms, meta)
torch.cuda.synchronize()
profiler.disable()
k = profiler.summary()
C = 1 if a.root_channel else J
alg = B * (V * C * 128 * 240 * 4 + C * 80 * 80 * 20 * 4)
print("config 2 (B=%d, C=%d): %.3f ms per batch = %.0f frames/s eager, %.3f ms = %.0f frames/s as one CUDA graph; kernels: %s" %
      (B, C, ms, B / ms * 1e3, ms_graph, B / ms_graph * 1e3, {kk: round(v["ms"], 3) for kk, v in k.items()}))
unp = k.get("unproject")
if unp:
    print("un-projection: %.1f us for %d algorithmic bytes = %.0f GB/s" % (unp["ms"] * 1e3, alg, alg / unp["ms"] * 1e-6))
