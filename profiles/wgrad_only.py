#!/usr/bin/env python
"""A few calls of the tcgen05 weight gradient on pose-net layer shapes (for ncu launch timing):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/wgrad_only.py"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from selfpose3d_b200 import grad_ops, ops  # noqa: E402

ops.set_float32_conv("bf16x3")
dev = "cuda:0"
for k, cin, cout, sp in [(3, 32, 32, 64), (7, 15, 16, 64), (3, 64, 64, 32), (3, 128, 128, 16)]:
    conv = nn.Conv3d(cin, cout, k, 1, k // 2).to(dev)
    pc = ops.PackedConv(conv.weight, conv.bias, None, 1, k // 2, relu=0)
    x = torch.randn(1, sp, sp, sp, ops.round_up(cin, 4), device=dev)
    gy = torch.randn(1, sp, sp, sp, ops.round_up(cout, 4), device=dev)
    for _ in range(2):
        grad_ops.conv_wgrad(pc, x, gy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        grad_ops.conv_wgrad(pc, x, gy)
    e1.record()
    torch.cuda.synchronize()
    print("wgrad k%d %d->%d @%d^3: %.3f ms per call" % (k, cin, cout, sp, e0.elapsed_time(e1) / 5))
