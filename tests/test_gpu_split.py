"""GPU tests of the float32-faithful tensor-core mode (``SP3D_CONV_TC_BF16X3``): float32 activations and weights
enter the tcgen05 kernel as sums of bf16 terms (``sp3d_split_bf16`` + split weight packing, 3 or 6 term pairs as
extra K blocks), float32 accumulation.  Checked against float64 CPU convolutions on the UN-rounded operands, against
the reference's golden vectors and against the float32 SIMT path.

Every test appends its measured error to ``gpurun_out/split_mode_errors.txt`` (when that directory exists)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

from oracle import nets  # noqa: E402
from selfpose3d_b200 import ops, synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import pose_resnet, v2v_net  # noqa: E402
from test_gpu_parity import _small_model, meta_from_golden  # noqa: E402
from test_gpu_tensorcore import rand_bn  # noqa: E402

DEV = "cuda:0"
DEFAULT_F32_CONV = ops.float32_conv()   # conftest: "simt" for the modules written against the float32 FMA kernels
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# per-convolution bound relative to the output range.  3 term pairs drop ~3 * 2^-18 per product; beyond that both
# variants sit on the tensor core's float32 accumulator, which loses ~2^-24 of its magnitude per MMA (measured
# 2e-6 .. 2e-5 per layer, growing with the number of accumulated MMAs: profiles/r01_split_mode_errors.txt)
TOL = {"bf16x3": 5e-5, "bf16x6": 5e-5}


def note(line):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "split_mode_errors.txt"), "a") as f:
            f.write(line + "\n")
    print(line)


@pytest.fixture(params=["bf16x3", "bf16x6"])
def mode(request):
    ops.set_float32_conv(request.param)
    yield request.param
    ops.set_float32_conv(DEFAULT_F32_CONV)


@pytest.mark.parametrize("blocks", [2, 3])
@pytest.mark.parametrize("C,pitch,c_block", [(15, 16, 16), (32, 32, 32), (1, 4, 16), (3, 5, 16), (64, 64, 64)])
def test_split_kernel_bit_exact(C, pitch, c_block, blocks):
    g = torch.Generator().manual_seed(C * 7 + blocks)
    x = torch.randn(3, 5, 7, pitch, generator=g) * torch.exp(4 * torch.randn(3, 5, 7, pitch, generator=g))
    y = ops.split_bf16(x.to(DEV), C, c_block, blocks).cpu().float()
    assert y.shape == (blocks, 3, 5, 7, c_block)
    for s, t in enumerate(ops.bf16_terms(x[..., :C], blocks)):
        assert torch.equal(y[s, ..., :C], t)
        assert not y[s, ..., C:].any()


CONV_CASES = [(7, 15, 16), (7, 1, 16), (3, 16, 32), (3, 32, 32), (3, 32, 64), (3, 64, 64), (3, 64, 128),
              (3, 128, 128), (1, 16, 32), (1, 32, 64), (1, 64, 128), (1, 32, 15), (1, 32, 1)]


@pytest.mark.parametrize("k,cin,cout", CONV_CASES)
def test_split_conv3d_matches_float64_reference(mode, k, cin, cout):
    torch.manual_seed(k * 1000 + cin * 10 + cout)
    conv = nn.Conv3d(cin, cout, k, 1, k // 2)
    bn = rand_bn(nn.BatchNorm3d(cout), cin + cout)
    x = torch.randn(2, cin, 6, 20, 12)
    res = torch.randn(2, cout, 6, 20, 12)
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())) + res.double())
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, k // 2, relu=1)
    pitch = ops.round_up(cout, 4)
    y = pc(ops.to_channel_last(x.to(DEV)), residual=ops.to_channel_last(res.to(DEV), c_pitch=pitch))
    assert y.dtype == torch.float32 and y.shape[-1] == pitch
    got = ops.to_channel_first(y, cout).cpu().double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("conv3d k%d %d->%d %s: %.3g" % (k, cin, cout, mode, err))
    assert err <= TOL[mode], err
    if pitch > cout:
        assert not y[..., cout:].any()
    # the same layer on the float32 SIMT kernel agrees within the same bound
    ysimt = ops.to_channel_first(pc(ops.to_channel_last(x.to(DEV)), residual=ops.to_channel_last(res.to(DEV), c_pitch=pitch),
                                    algo=0), cout).cpu().double()
    assert float((got - ysimt).abs().max()) / float(want.abs().max()) <= TOL[mode]


@pytest.mark.parametrize("k,cin,cout,shape,with_res", [(3, 16, 32, (5, 17, 32), True), (3, 32, 32, (6, 20, 16), True),
                                                       (3, 32, 32, (5, 17, 32), False), (7, 1, 16, (6, 20, 12), False),
                                                       (7, 15, 16, (5, 17, 34), False), (7, 1, 16, (3, 9, 7), False)])
def test_split_conv3d_folded_forms(mode, k, cin, cout, shape, with_res):
    """The z-folded 3^3 / 7^3 kernels and the tap-stacked 1-channel stem on split operands (odd W: plain kernel)."""
    torch.manual_seed(k * 1000 + cin * 10 + shape[2])
    conv = nn.Conv3d(cin, cout, k, 1, k // 2)
    bn = rand_bn(nn.BatchNorm3d(cout), cin + cout)
    x = torch.randn(2, cin, *shape)
    res = torch.randn(2, cout, *shape) if with_res else None
    with torch.no_grad():
        want = bn.double()(conv.double()(x.double()))
        want = F.relu(want + res.double() if with_res else want)
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, k // 2, relu=1)
    y = pc(ops.to_channel_last(x.to(DEV)), residual=ops.to_channel_last(res.to(DEV)) if with_res else None)
    got = ops.to_channel_first(y, cout).cpu().double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("conv3d folded k%d %d->%d %s %s: %.3g" % (k, cin, cout, shape, mode, err))
    assert err <= TOL[mode], err


@pytest.mark.parametrize("shape", [(4, 16, 8), (3, 20, 12)])
@pytest.mark.parametrize("cin,cout", [(128, 64), (64, 32)])
def test_split_transposed_conv3d(mode, cin, cout, shape):
    torch.manual_seed(cin + shape[0])
    ct, bn = nn.ConvTranspose3d(cin, cout, 2, 2), rand_bn(nn.BatchNorm3d(cout), cin)
    x = torch.randn(2, cin, *shape)
    skip = torch.randn(2, cout, *[2 * s for s in shape])
    with torch.no_grad():
        want = F.relu(bn.double()(ct.double()(x.double()))) + skip.double()
    ct, bn = ct.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(ct.weight, ct.bias, bn, 2, 0, transposed=True, relu=2)
    y = pc(ops.to_channel_last(x.to(DEV)), residual=ops.to_channel_last(skip.to(DEV)))
    got = ops.to_channel_first(y, cout).cpu().double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("convT3d %d->%d %s %s: %.3g" % (cin, cout, shape, mode, err))
    assert err <= TOL[mode], err


CONV2D_CASES = [(1, 1, 0, 64, 256), (1, 1, 0, 1024, 512), (1, 2, 0, 256, 512), (3, 1, 1, 128, 128), (1, 1, 0, 256, 15)]


@pytest.mark.parametrize("k,s,p,cin,cout", CONV2D_CASES)
def test_split_conv2d(mode, k, s, p, cin, cout):
    torch.manual_seed(k * 100 + s * 10 + cin + cout)
    conv = nn.Conv2d(cin, cout, k, s, p, bias=(cout == 15))
    bn = rand_bn(nn.BatchNorm2d(cout), cin + cout)
    x = torch.randn(3, cin, 17, 11)
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())))
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, s, p, relu=1)
    assert pc.tc_supported()
    y = pc(ops.to_channel_last(x.unsqueeze(2).to(DEV)))
    got = ops.to_channel_first(y, cout)[:, :, 0].cpu().double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("conv2d k%d s%d %d->%d %s: %.3g" % (k, s, cin, cout, mode, err))
    assert err <= TOL[mode], err


def test_split_deconv2d_k4s2(mode):
    torch.manual_seed(9)
    ct, bn = nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False), rand_bn(nn.BatchNorm2d(256), 3)
    x = torch.randn(2, 256, 9, 7)
    with torch.no_grad():
        want = F.relu(bn.double()(ct.double()(x.double())))
    ct, bn = ct.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(ct.weight, None, bn, 2, 1, transposed=True, relu=1)
    y = pc(ops.to_channel_last(x.unsqueeze(2).to(DEV)))
    got = ops.to_channel_first(y, 256)[:, :, 0].cpu().double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("deconv2d k4s2 256->256 %s: %.3g" % (mode, err))
    assert err <= TOL[mode], err


def test_split_v2v_net_vs_float64_oracle(mode):
    """Whole V2VNet(15,15) on a 32^3 cube and V2VNet(1,1) on a 40x40x12 grid against the north star's heat-map bar
    (1e-4 of the range); the float32 oracle's own distance to float64 is printed beside it."""
    for cin, shape, seed in ((15, (1, 15, 32, 32, 32), 41), (1, (2, 1, 40, 40, 12), 42)):
        net = v2v_net.V2VNet(cin, cin)
        sd = synthetic.trained_like_state_dict(net, seed=seed)
        net.load_state_dict(sd, strict=True)
        x = torch.from_numpy(np.random.RandomState(seed).rand(*shape).astype(np.float32))
        y64 = nets.v2v_forward(x, sd, dtype=torch.float64)
        y32 = nets.v2v_forward(x, sd)
        y = net.to(DEV).eval()(x.to(DEV)).cpu().double()
        scale = float(y64.abs().max())
        err = float((y - y64).abs().max()) / scale
        ref = float((y32.double() - y64).abs().max()) / scale
        note("V2VNet(%d) %s: max error / range = %.3g (float32 CPU oracle: %.3g)" % (cin, mode, err, ref))
        assert err <= 1e-4, (err, ref)


def test_split_pose_resnet_vs_reference_golden(mode, golden):
    g = golden("pose_resnet50")
    net = pose_resnet.get_pose_net(default_config(), is_train=False)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(g["seed"])), strict=True)
    y = net.to(DEV).eval()(torch.from_numpy(g["x"]).to(DEV)).cpu().numpy()
    err = float(np.abs(y - g["y"]).max()) / float(np.abs(g["y"]).max())
    note("PoseResNet-50 %s vs reference golden: max error / range = %.3g" % (mode, err))
    assert err <= 3e-4, err


def test_split_inference_vs_reference_golden(mode, golden):
    """The reference's golden inference case (heat-maps -> proposals -> joints) with every covered convolution on
    the tensor cores: same proposals, joints within 0.15 mm of the reference's float32 CPU result (measured 0.02 -
    0.04 mm; the float32 SIMT path is at 1e-2 mm; beta = 100 amplifies logit noise)."""
    g = golden("inference_small")
    model, _ = _small_model(g, float(g["threshold"]))
    hms = [torch.from_numpy(h).to(DEV) for h in g["heatmaps"]]
    pred, _, gc = model(views1=None, meta1=meta_from_golden(g), input_heatmaps1=hms, inference=True)
    pred, gc = pred.cpu().numpy(), gc.cpu().numpy()
    np.testing.assert_allclose(gc[..., :3], g["grid_centers"][..., :3], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gc[..., 4], g["grid_centers"][..., 4], rtol=0, atol=2e-4)
    valid = g["pred"][:, :, 0, 3] >= 0
    assert np.array_equal(pred[:, :, 0, 3] >= 0, valid)
    err = float(np.abs(pred[valid][..., :3] - g["pred"][valid][..., :3]).max())
    note("inference_small %s: max joint error = %.3g mm" % (mode, err))
    assert err <= 0.15, err
