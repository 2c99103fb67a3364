"""GPU: the tcgen05 weight gradient (``sp3d_conv_wgrad_tc``, csrc/conv_wgrad_tc.cu) of the V2VNet ``nn.Conv3d`` layers
against float64 autograd of the reference's layer and against the float32 FMA kernel (``sp3d_conv_wgrad``): 3 bf16
term pairs, float32 accumulation -> 5e-5 of the gradient's range."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

from selfpose3d_b200 import grad_ops, ops  # noqa: E402

DEV = "cuda:0"
CASES = [(3, 32, 32, (4, 6, 64), 2), (3, 16, 32, (3, 5, 64), 1), (7, 15, 16, (8, 7, 64), 1), (3, 64, 64, (5, 4, 32), 2),
         (3, 32, 64, (4, 4, 32), 1), (3, 128, 128, (4, 5, 16), 2), (3, 64, 128, (3, 3, 16), 1), (1, 16, 32, (3, 4, 64), 1),
         (1, 32, 15, (2, 3, 128), 2)]


def cl(x):
    N, C = x.shape[:2]
    out = torch.zeros((N,) + tuple(x.shape[2:]) + (ops.round_up(C, 4),), device=DEV)
    out[..., :C] = x.to(DEV).permute(0, 2, 3, 4, 1)
    return out.contiguous()


@pytest.mark.parametrize("k,cin,cout,sp,n", CASES)
def test_wgrad_tc_vs_float64_autograd(k, cin, cout, sp, n):
    torch.manual_seed(k * 100 + cin + cout)
    conv = nn.Conv3d(cin, cout, k, 1, k // 2).double()
    x = torch.randn(n, cin, *sp, dtype=torch.float64)
    y = conv(x)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    pc = ops.PackedConv(conv.weight.detach().float().to(DEV), conv.bias.detach().float().to(DEV), None, 1, k // 2, relu=0)
    xcl, gycl = cl(x.float()), cl(gy.float())
    ops.set_float32_conv("bf16x3")
    assert grad_ops._wgrad_tc_ok(pc, xcl, gycl)
    before = ops._lib.launch_count
    gw, gb = grad_ops.conv_wgrad(pc, xcl, gycl)
    assert ops._lib.launch_count - before == 3
    want_w, want_b = conv.weight.grad.numpy(), conv.bias.grad.numpy()
    err_w = float(np.abs(gw.cpu().numpy() - want_w).max() / np.abs(want_w).max())
    err_b = float(np.abs(gb.cpu().numpy() - want_b).max() / np.abs(want_b).max())
    print("wgrad tc k%d %d->%d %s x%d: dW %.3g  db %.3g" % (k, cin, cout, sp, n, err_w, err_b))
    assert gw.shape == conv.weight.shape and err_w <= 5e-5 and err_b <= 2e-5, (err_w, err_b)
    ops.set_float32_conv("simt")
    gw_s, gb_s = grad_ops.conv_wgrad(pc, xcl, gycl)
    assert float((gw - gw_s).abs().max() / gw_s.abs().max()) <= 5e-5


# (what, module, input shape [N, C, *spatial]): the generalised form -- rows that are not multiples of 16 (padded lines),
# 2-D layers of PoseResNet (1x1 / 3x3, up to 2048 channels: ci tiles and co blocks), transposed convolutions by phase
GENERAL = [("3d k3 z=20 (root grid)", lambda: nn.Conv3d(32, 32, 3, 1, 1), (2, 32, 6, 5, 20)),
           ("3d k7 1->16 z=20", lambda: nn.Conv3d(1, 16, 7, 1, 3), (1, 1, 8, 8, 20)),
           ("3d k3 z=10", lambda: nn.Conv3d(64, 64, 3, 1, 1), (1, 64, 5, 4, 10)),
           ("3d k3 z=5", lambda: nn.Conv3d(128, 128, 3, 1, 1), (2, 128, 4, 4, 5)),
           ("3d convT k2 s2", lambda: nn.ConvTranspose3d(64, 32, 2, 2), (2, 64, 4, 3, 8)),
           ("3d convT k2 s2 128->64 z=5", lambda: nn.ConvTranspose3d(128, 64, 2, 2), (1, 128, 3, 3, 5)),
           ("2d 1x1 64->256", lambda: nn.Conv2d(64, 256, 1, bias=False), (2, 64, 9, 24)),
           ("2d 1x1 1024->256 w=12", lambda: nn.Conv2d(1024, 256, 1, bias=False), (2, 1024, 5, 12)),
           ("2d 1x1 512->2048 w=12", lambda: nn.Conv2d(512, 2048, 1, bias=False), (1, 512, 4, 12)),
           ("2d 3x3 64->64 w=96", lambda: nn.Conv2d(64, 64, 3, 1, 1, bias=False), (2, 64, 7, 96)),
           ("2d 3x3 256->256 w=24", lambda: nn.Conv2d(256, 256, 3, 1, 1, bias=False), (2, 256, 6, 24)),
           ("2d 3x3 512->512 w=12", lambda: nn.Conv2d(512, 512, 3, 1, 1, bias=False), (1, 512, 9, 12)),
           ("2d convT k4 s2 p1 256->256", lambda: nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False), (2, 256, 5, 12)),
           ("2d convT k4 s2 p1 2048->256", lambda: nn.ConvTranspose2d(2048, 256, 4, 2, 1, bias=False), (1, 2048, 3, 12)),
           ("2d 1x1 256->15 head", lambda: nn.Conv2d(256, 15, 1), (2, 256, 8, 96))]


@pytest.mark.parametrize("what,make,shape", GENERAL, ids=[g[0] for g in GENERAL])
def test_general_wgrad_tc_vs_float64_autograd(what, make, shape):
    torch.manual_seed(len(what) + shape[1])
    mod = make().double()
    x = torch.randn(*shape, dtype=torch.float64)
    y = mod(x)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    transposed = isinstance(mod, (nn.ConvTranspose2d, nn.ConvTranspose3d))
    bias = None if mod.bias is None else mod.bias.detach().float().to(DEV)
    pc = ops.PackedConv(mod.weight.detach().float().to(DEV), bias, None, int(mod.stride[0]), int(mod.padding[0]),
                        transposed=transposed, relu=0)
    x5, g5 = (x, gy) if x.dim() == 5 else (x.unsqueeze(2), gy.unsqueeze(2))
    xcl, gycl = cl(x5.float()), cl(g5.float())
    ops.set_float32_conv("bf16x3")
    assert grad_ops._wgrad_tc_ok(pc, xcl, gycl)
    gw, gb = grad_ops.conv_wgrad(pc, xcl, gycl, with_bias=bias is not None)
    want_w = mod.weight.grad.numpy()
    err_w = float(np.abs(gw.cpu().numpy() - want_w).max() / np.abs(want_w).max())
    print("wgrad tc %s: dW %.3g" % (what, err_w))
    assert gw.shape == mod.weight.shape and err_w <= 5e-5, err_w
    if bias is not None:
        want_b = mod.bias.grad.numpy()
        assert float(np.abs(gb.cpu().numpy() - want_b).max() / np.abs(want_b).max()) <= 2e-5
    ops.set_float32_conv("simt")
    gw_s, _ = grad_ops.conv_wgrad(pc, xcl, gycl, with_bias=bias is not None)
    assert float((gw - gw_s).abs().max() / gw_s.abs().max()) <= 5e-5


def test_shapes_outside_the_tensor_core_form_take_the_fma_kernel():
    ops.set_float32_conv("bf16x3")
    conv = nn.Conv2d(64, 128, 3, 2, 1).to(DEV)            # strided convolutions: taps step by two input positions
    pc = ops.PackedConv(conv.weight, conv.bias, None, 2, 1, relu=0)
    x = torch.randn(1, 1, 8, 16, 64, device=DEV)
    assert not grad_ops._wgrad_tc_ok(pc, x, x)
    conv = nn.Conv2d(3, 64, 3, 1, 1).to(DEV)              # rows longer than 128 positions
    assert not grad_ops._wgrad_tc_ok(ops.PackedConv(conv.weight, conv.bias, None, 1, 1, relu=0),
                                     torch.randn(1, 1, 4, 192, 4, device=DEV), x)
    ops.set_float32_conv("simt")
    conv = nn.Conv3d(32, 32, 3, 1, 1).to(DEV)
    assert not grad_ops._wgrad_tc_ok(ops.PackedConv(conv.weight, conv.bias, None, 1, 1, relu=0),
                                     torch.randn(1, 8, 8, 16, 32, device=DEV), x)
    ops.set_float32_conv("bf16x3")
