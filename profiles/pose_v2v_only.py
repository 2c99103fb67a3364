#!/usr/bin/env python
"""The pose net's V2VNet on N random person cubes in the float32-faithful tensor-core mode (term-pair activations),
between cudaProfilerStart/Stop -- a short command for `ncu --set full` captures of the convolution kernels:

  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -c 40 \
      -o gpurun_out/r02_ncu_pose_v2v python profiles/pose_v2v_only.py [--cubes 80]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from selfpose3d_b200 import ops, synthetic  # noqa: E402
from selfpose3d_b200.models import v2v_net  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cubes", type=int, default=80)
ap.add_argument("--layers", default="all", help="'stem' = the 7^3 stem only")
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--stalls", action="store_true", help="print the kernel's own per-CTA wait counters (sp3d_debug_conv_profile)")
ap.add_argument("--pair", type=int, default=0, help="0: single CTAs only (sp3d_debug_conv_pair), 1: CTA pairs share the weight stream")
a = ap.parse_args()
from selfpose3d_b200 import _lib as _l  # noqa: E402
_l.load().sp3d_debug_conv_pair(a.pair)
ops.set_volume_dtype(torch.float32)
ops.set_float32_conv("bf16x3")
net = v2v_net.V2VNet(15, 15)
net.load_state_dict(synthetic.trained_like_state_dict(net, seed=1), strict=True)
net = net.cuda().eval()
x = torch.rand(a.cubes, 64, 64, 64, 16, device="cuda")
x[..., 15] = 0
xs = ops.split_act(x, 15)
del x


def run():
    if a.layers == "stem":
        return net.front_layers[0].forward_cl(xs)
    return net.forward_cl(xs)


with torch.no_grad():
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    for _ in range(a.repeat):
        run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.profiler.stop()
print("%s x %d cubes: %.2f ms per pass" % (a.layers, a.cubes, dt / a.repeat * 1e3))
if a.stalls:
    import ctypes as C
    from selfpose3d_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")

    def counters(name, fn):
        with torch.no_grad():
            fn()
            torch.cuda.synchronize()
            lib.sp3d_debug_conv_profile(C.c_void_p(buf.data_ptr()))
            buf.zero_()
            fn()
            torch.cuda.synchronize()
            lib.sp3d_debug_conv_profile(C.c_void_p(0))
        t = buf.view(148, 16).double().cpu()
        t = t[t[:, 0] > 0]
        m = (t.mean(0) / 1e3).tolist()
        print("%-22s per-CTA mean kclk: mma total %.1f | wait halo %.1f (%.1f%%), wait weights %.1f (%.1f%%), wait acc_empty %.1f "
              "(%.1f%%) | epilogue busy %.1f (%.1f%%) | items %.1f | epi: ld %.1f res %.1f math %.1f fence+bar %.1f store+ring %.1f"
              % (name, m[0], m[1], 100 * m[1] / m[0], m[2], 100 * m[2] / m[0], m[3], 100 * m[3] / m[0], m[4] - m[5],
                 100 * (m[4] - m[5]) / max(m[4], 1e-9), m[6] * 1e3, m[8], m[9], m[10], m[11], m[12]))

    with torch.no_grad():
        h1 = net.front_layers[0].forward_cl(xs)                      # [2, n, 64^3, 16]
        c16, c32b, _ = net.front_layers[1]._packed()
        h2 = c16(h1)                                                  # 16 -> 32
        pool = net.encoder_decoder.encoder_pool1.forward_cl(h2, 32)
        e_a, e_b, e_s = net.encoder_decoder.encoder_res1._packed()    # 32 -> 64, 64 -> 64 at 32^3
        h3 = e_a(pool)
        h4 = e_b(h3, residual=h3)
        pool2 = net.encoder_decoder.encoder_pool2.forward_cl(h4, 64)
        f_a, f_b, f_s = net.encoder_decoder.encoder_res2._packed()    # 64 -> 128, 128 -> 128 at 16^3
        h5 = f_a(pool2)
    counters("7^3 15->16 stem", lambda: net.front_layers[0].forward_cl(xs))
    counters("3^3 16->32", lambda: c16(h1))
    counters("3^3 32->32 +res", lambda: c32b(h2, residual=h2))
    counters("3^3 32->64 @32^3", lambda: e_a(pool))
    counters("3^3 64->64 +res @32^3", lambda: e_b(h3, residual=h3))
    counters("3^3 64->128 @16^3", lambda: f_a(pool2))
    counters("3^3 128->128 +res @16^3", lambda: f_b(h5, residual=h5))
