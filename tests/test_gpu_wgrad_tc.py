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


def test_shapes_outside_the_tensor_core_form_take_the_fma_kernel():
    ops.set_float32_conv("bf16x3")
    conv = nn.Conv3d(32, 32, 3, 1, 1).to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, None, 1, 1, relu=0)
    x = torch.randn(1, 8, 8, 20, 32, device=DEV)          # z extent 20: the root grid
    assert not grad_ops._wgrad_tc_ok(pc, x, x)
    ct = nn.ConvTranspose3d(32, 16, 2, 2).to(DEV)
    assert not grad_ops._wgrad_tc_ok(ops.PackedConv(ct.weight, ct.bias, None, 2, 0, transposed=True, relu=0), x, x)
