#!/usr/bin/env python
"""Where the tcgen05 convolution pipeline waits, per layer shape of the bench step:
per-CTA clock64 counters written by the kernel itself (sp3d_debug_conv_profile) plus the CUDA-event
time of the launch.   python profiles/conv_stalls.py [--cubes 16]"""
import argparse
import ctypes as C
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from selfpose3d_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cubes", type=int, default=16)
a = ap.parse_args()
dev = "cuda:0"
lib = _lib.load()
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)

# (kind, k, cin, cout, spatial, residual, out f32)
CASES = [("conv", 7, 15, 16, 64, False, False), ("conv", 3, 16, 32, 64, False, False),
         ("conv", 3, 32, 32, 64, True, False), ("conv", 3, 32, 32, 64, False, False),
         ("conv", 1, 16, 32, 64, False, False), ("conv", 1, 32, 15, 64, False, True),
         ("conv", 3, 32, 64, 32, False, False), ("conv", 3, 64, 64, 32, True, False),
         ("conv", 1, 32, 64, 32, False, False),
         ("conv", 3, 64, 128, 16, False, False), ("conv", 3, 128, 128, 16, True, False),
         ("convT", 2, 64, 32, 32, True, False), ("convT", 2, 128, 64, 16, True, False)]

print("%-34s %8s %8s | per-CTA mean kclk: %8s %8s %8s %8s | %8s %8s" %
      ("layer", "us", "TFLOP/s", "mma_tot", "w_halo", "w_wgt", "w_acc", "epi_tot", "epi_wait"))
for kind, k, cin, cout, sp, with_res, f32 in CASES:
    torch.manual_seed(0)
    if kind == "conv":
        m = nn.Conv3d(cin, cout, k, 1, k // 2).to(dev)
        pc = ops.PackedConv(m.weight, m.bias, nn.BatchNorm3d(cout).to(dev).eval(), 1, k // 2, relu=1)
        osp = sp
    else:
        m = nn.ConvTranspose3d(cin, cout, 2, 2, 0).to(dev)
        pc = ops.PackedConv(m.weight, m.bias, nn.BatchNorm3d(cout).to(dev).eval(), 2, 0, transposed=True, relu=2)
        osp = sp * 2
    x = torch.randn(a.cubes, sp, sp, sp, ops.round_up(cin, 16), device=dev).to(torch.bfloat16)
    pitch = ops.round_up(cout, 16)
    odt = torch.float32 if f32 else torch.bfloat16
    res = torch.randn(a.cubes, osp, osp, osp, pitch, device=dev).to(odt) if with_res else None
    for _ in range(2):
        pc(x, residual=res, out_pitch=pitch, out_dtype=odt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pc(x, residual=res, out_pitch=pitch, out_dtype=odt)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 3
    flops = 2.0 * a.cubes * sp ** 3 * cin * cout * (k ** 3 if kind == "conv" else 1)
    lib.sp3d_debug_conv_profile(C.c_void_p(buf.data_ptr()))
    buf.zero_()
    pc(x, residual=res, out_pitch=pitch, out_dtype=odt)
    torch.cuda.synchronize()
    lib.sp3d_debug_conv_profile(C.c_void_p(0))
    t = buf.view(148, 16).double().cpu()
    t = t[t[:, 0] > 0]
    mean = (t.mean(0) / 1e3).tolist()
    print("%-34s %8.1f %8.1f | %35s %8.1f %8.1f %8.1f %8.1f | %8.1f %8.1f   items/CTA %.1f" %
          ("%s k%d %d->%d @%dx%d^3%s%s" % (kind, k, cin, cout, a.cubes, sp, " +res" if with_res else "", " f32" if f32 else ""),
           us, flops / us * 1e-6, "", mean[0], mean[1], mean[2], mean[3], mean[4], mean[5], mean[6] * 1e3),
          " epi: ld %.0f res %.0f math %.0f fence+bar %.0f store+ring %.0f" % tuple(mean[8:13]))
