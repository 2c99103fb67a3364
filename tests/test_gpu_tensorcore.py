"""GPU tests of the tcgen05 (bf16 operands, float32 accumulation) convolution path against the
float64 CPU reference evaluated on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

from oracle import nets  # noqa: E402
from selfpose3d_b200 import ops, synthetic  # noqa: E402
from selfpose3d_b200.models import v2v_net  # noqa: E402

DEV = "cuda:0"


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def rand_bn(bn, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(bn.bias.shape, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.2)
        bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
    return bn.eval()


def to_cl_bf16(x):   # [N,C,X,Y,Z] float (bf16-representable) -> channel-last bf16, pitch = C rounded up to 16
    return ops.to_channel_last(x.to(DEV), c_pitch=ops.round_up(x.shape[1], 16), dtype=torch.bfloat16)


# every (kernel, cin, cout) the V2V nets use; spatial extent with partial bricks in x, y and z
CONV_CASES = [(7, 15, 16), (7, 1, 16), (3, 16, 32), (3, 32, 32), (3, 32, 64), (3, 64, 64), (3, 64, 128),
              (3, 128, 128), (1, 16, 32), (1, 32, 64), (1, 64, 128), (1, 32, 15), (1, 32, 1)]


@pytest.mark.parametrize("k,cin,cout", CONV_CASES)
def test_tc_conv_matches_float64_reference(k, cin, cout):
    torch.manual_seed(k * 1000 + cin * 10 + cout)
    conv = nn.Conv3d(cin, cout, k, 1, k // 2)
    bn = rand_bn(nn.BatchNorm3d(cout), cin + cout)
    with torch.no_grad():
        conv.weight.copy_(bf16_round(conv.weight))
    x = bf16_round(torch.randn(2, cin, 6, 20, 12))
    res = bf16_round(torch.randn(2, cout, 6, 20, 12))
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())) + res.double())
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, k // 2, relu=1)
    pitch = ops.round_up(cout, 16)
    res_cl = ops.to_channel_last(res.to(DEV), c_pitch=pitch, dtype=torch.float32)
    y = pc(to_cl_bf16(x), residual=res_cl, out_pitch=pitch, out_dtype=torch.float32)
    got = ops.to_channel_first(y, cout).cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale, float((got - want).abs().max()) / scale
    if pitch > cout:
        assert not y[..., cout:].any()
    # bf16 output + bf16 residual: one bf16 rounding of the result on top
    yb = pc(to_cl_bf16(x), residual=res_cl.to(torch.bfloat16), out_pitch=pitch)
    gotb = ops.to_channel_first(yb, cout, dtype=torch.float32).cpu().double()
    assert float((gotb - want).abs().max()) <= 6e-3 * scale


@pytest.mark.parametrize("shape", [(4, 16, 8), (3, 20, 12)])     # whole bricks / ragged in x, y and z
@pytest.mark.parametrize("with_skip", [True, False])
@pytest.mark.parametrize("cin,cout", [(128, 64), (64, 32)])
def test_tc_transposed_conv_matches_float64_reference(cin, cout, with_skip, shape):
    """k2/s2 transposed convolution (one fused launch: the 8 output phases are GEMM columns) + BN + ReLU (+ skip)."""
    torch.manual_seed(cin)
    ct = nn.ConvTranspose3d(cin, cout, 2, 2)
    bn = rand_bn(nn.BatchNorm3d(cout), cin)
    with torch.no_grad():
        ct.weight.copy_(bf16_round(ct.weight))
    x = bf16_round(torch.randn(2, cin, *shape))
    skip = bf16_round(torch.randn(2, cout, *[2 * v for v in shape]))
    with torch.no_grad():
        want = F.relu(bn.double()(ct.double()(x.double())))
        if with_skip:
            want = want + skip.double()
    ct, bn = ct.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(ct.weight, ct.bias, bn, 2, 0, transposed=True, relu=2)
    assert pc._tc_fused_ok(cout, torch.float32) and pc._tc_fused_ok(cout, torch.bfloat16)
    skip_cl = ops.to_channel_last(skip.to(DEV), c_pitch=cout, dtype=torch.float32)
    y = pc(to_cl_bf16(x), residual=skip_cl if with_skip else None, out_pitch=cout, out_dtype=torch.float32)
    got = ops.to_channel_first(y, cout).cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale
    yb = pc(to_cl_bf16(x), residual=skip_cl.to(torch.bfloat16) if with_skip else None, out_pitch=cout)
    gotb = ops.to_channel_first(yb, cout, dtype=torch.float32).cpu().double()
    assert float((gotb - want).abs().max()) <= 6e-3 * scale


@pytest.mark.parametrize("cin", [16, 32])
@pytest.mark.parametrize("with_res", [True, False])
@pytest.mark.parametrize("shape", [(6, 20, 16), (5, 17, 32)])
def test_tc_conv3_zfold_matches_float64_reference(cin, with_res, shape):
    """3^3 cin -> 32 layers run z-folded (N = 2 x 32) when W is even; with and without the fused residual."""
    torch.manual_seed(cin + 3)
    conv = nn.Conv3d(cin, 32, 3, 1, 1)
    bn = rand_bn(nn.BatchNorm3d(32), cin)
    with torch.no_grad():
        conv.weight.copy_(bf16_round(conv.weight))
    x = bf16_round(torch.randn(2, cin, *shape))
    res = bf16_round(torch.randn(2, 32, *shape))
    with torch.no_grad():
        want = bn.double()(conv.double()(x.double()))
        want = F.relu(want + res.double()) if with_res else F.relu(want)
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, 1, relu=1)
    assert pc._tc_zfold_ok(shape[2], 32)
    res_cl = ops.to_channel_last(res.to(DEV), c_pitch=32, dtype=torch.float32) if with_res else None
    y = pc(to_cl_bf16(x), residual=res_cl, out_pitch=32, out_dtype=torch.float32)
    got = ops.to_channel_first(y, 32).cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale, float((got - want).abs().max()) / scale
    yb = pc(to_cl_bf16(x), residual=res_cl.to(torch.bfloat16) if with_res else None, out_pitch=32)
    gotb = ops.to_channel_first(yb, 32, dtype=torch.float32).cpu().double()
    assert float((gotb - want).abs().max()) <= 6e-3 * scale


@pytest.mark.parametrize("cin", [15, 1])
@pytest.mark.parametrize("shape", [(6, 20, 12), (5, 17, 34), (3, 9, 7)])    # even W: z-folded kernel; odd W: plain kernel
def test_tc_stem_conv7_zfold_matches_float64_reference(cin, shape):
    """The 7^3 stem: two output positions along W share one GEMM row (N = 2 x 16) when W is even."""
    torch.manual_seed(cin)
    conv = nn.Conv3d(cin, 16, 7, 1, 3)
    bn = rand_bn(nn.BatchNorm3d(16), cin)
    with torch.no_grad():
        conv.weight.copy_(bf16_round(conv.weight))
    x = bf16_round(torch.randn(2, cin, *shape))
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())))
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, 3, relu=1)
    assert pc._tc_zfold_ok(shape[2], 16) == (shape[2] % 2 == 0)
    y = pc(to_cl_bf16(x), out_pitch=16, out_dtype=torch.float32)
    got = ops.to_channel_first(y, 16).cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale, float((got - want).abs().max()) / scale
    yb = pc(to_cl_bf16(x), out_pitch=16)
    gotb = ops.to_channel_first(yb, 16, dtype=torch.float32).cpu().double()
    assert float((gotb - want).abs().max()) <= 6e-3 * scale


@pytest.mark.parametrize("k,p,cin,cout,hw", [(3, 1, 128, 128, (24, 18)), (3, 1, 256, 256, (12, 34)), (7, 3, 3, 64, (40, 36))])
def test_tc_stride2_conv_space_to_depth_matches_float64_reference(k, p, cin, cout, hw):
    """Stride-2 3x3 / 7x7 convolutions as stride-1 tensor-core convolutions over the 2x2 space-to-depth tensor."""
    torch.manual_seed(k * 100 + cin)
    conv = nn.Conv2d(cin, cout, k, 2, p, bias=False)
    bn = rand_bn(nn.BatchNorm2d(cout), cin)
    with torch.no_grad():
        conv.weight.copy_(bf16_round(conv.weight))
    x = bf16_round(torch.randn(3, cin, *hw))
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())))
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    sc = ops.S2DConv(conv.weight, bn, p, relu=1)
    xd = x.to(DEV)
    if cin == 3:      # the stem reads the channel-first float32 image directly
        y = sc(xd, xd.stride(), 3, hw[0], hw[1])
    else:             # trunk convolutions read channel-last bf16 activations
        xcl = ops.to_channel_last(xd.unsqueeze(2), c_pitch=cin, dtype=torch.bfloat16)
        y = sc(xcl, (hw[0] * hw[1] * cin, 1, hw[1] * cin, cin), 3, hw[0], hw[1])
    got = ops.to_channel_first(y, cout, dtype=torch.float32)[:, :, 0].cpu().double()
    scale = float(want.abs().max())
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 6e-3 * scale, float((got - want).abs().max()) / scale


@pytest.mark.parametrize("n,shape,J", [(5, (32, 32, 32), 15), (3, (20, 24, 12), 15), (2, (16, 16, 16), 4)])
def test_tc_output_layer_with_fused_softargmax_head(n, shape, J):
    """V2VNet's 1x1x1 output layer + soft-argmax in one kernel (the heat-map volume never reaches HBM) against the
    two-kernel path (float32 volume + sp3d_softargmax3d_fwd) on the same inputs, and against a float64 evaluation
    of the soft-argmax on that float32 volume.  CTAs span cube boundaries in the first case (320 items / 148 CTAs)."""
    torch.manual_seed(n * 10 + J)
    conv = nn.Conv3d(32, J, 1).to(DEV)
    with torch.no_grad():
        conv.weight.mul_(0.05)
    pc = ops.PackedConv(conv.weight, conv.bias, None, 1, 0, relu=0)
    X, Y, Z = shape
    x = torch.randn(n, X, Y, Z, 32, device=DEV).to(torch.bfloat16)
    centers = torch.tensor([[100.0 * i, -500.0 + 37.0 * i, 800.0, 0.0, 1.0] for i in range(n)], device=DEV)
    beta, grid = 100.0, [2000.0] * 3
    vol = pc(x, out_pitch=16, out_dtype=torch.float32)
    two = ops.softargmax(vol, (X * Y * Z * 16, 1, 16), n, J, shape, centers, grid, beta)
    head = ops.SoftargmaxHead(n, J, shape, centers, grid, beta)
    before = ops._lib.launch_count
    fused = pc(x, head=head)
    assert ops._lib.launch_count - before == 2 and fused.shape == (n, J, 3)
    assert float((fused - two).abs().max()) <= 2e-3, float((fused - two).abs().max())
    # float64 soft-argmax of the same float32 logits
    lin = [torch.linspace(-grid[a] / 2, grid[a] / 2, shape[a]) for a in range(3)]
    v = vol[..., :J].double().cpu().reshape(n, -1, J)
    w = torch.softmax(beta * v, dim=1)
    for i in range(n):
        g = torch.stack(torch.meshgrid(*[(lin[a] + centers[i, a].cpu()).float() for a in range(3)], indexing="ij"), -1)
        want = (w[i].T @ g.reshape(-1, 3).double())
        assert float((fused[i].cpu().double() - want).abs().max()) <= 2e-3


@pytest.mark.parametrize("shape", [(8, 12, 16), (6, 10, 4)])
def test_maxpool_bf16_k2s2_exact(shape):
    """V2VNet's 2x2x2 pools on bf16 channel-last activations: exact (max of bf16 values is a bf16 value)."""
    torch.manual_seed(1)
    x = bf16_round(torch.randn(3, 32, *shape))
    want = F.max_pool3d(x, 2, 2)
    y = ops.maxpool(to_cl_bf16(x), 32, [2, 2, 2], [2, 2, 2], [0, 0, 0])
    got = ops.to_channel_first(y, 32, dtype=torch.float32).cpu()
    assert torch.equal(got, want)


def test_v2v_net_bf16_mode_vs_float64_oracle():
    """Whole V2VNet(15,15) on a 32^3 cube and V2VNet(1,1) on a 40x40x12 grid in bf16 tensor-core mode.
    bf16 activations carry ~3 significant digits; the result is compared with the float64 oracle relative
    to the output range (this is the throughput mode, not the parity mode)."""
    for cin, shape, seed in ((15, (1, 15, 32, 32, 32), 41), (1, (2, 1, 40, 40, 12), 42)):
        net = v2v_net.V2VNet(cin, cin)
        sd = synthetic.trained_like_state_dict(net, seed=seed)
        net.load_state_dict(sd, strict=True)
        x = torch.from_numpy(np.random.RandomState(seed).rand(*shape).astype(np.float32))
        y64 = nets.v2v_forward(x, sd, dtype=torch.float64)
        ops.set_volume_dtype(torch.bfloat16)
        try:
            y = net.to(DEV).eval()(x.to(DEV)).cpu().double()
            if cin == 1:   # the root net's pitch-1 float32 score volume (what sp3d_nms_topk3d reads)
                xcl = ops.to_channel_last(x.to(DEV), c_pitch=16, dtype=torch.bfloat16)
                y1 = net.forward_cl(xcl, out_pitch=1)
                assert y1.shape == (2, 40, 40, 12, 1) and y1.dtype == torch.float32
                assert torch.equal(y1[..., 0].cpu().double(), y[:, 0])
        finally:
            ops.set_volume_dtype(torch.float32)
        err = float((y - y64).abs().max()) / float(y64.abs().max())
        print("V2VNet(%d) bf16 mode: max error / output range = %.3g" % (cin, err))
        assert err < 5e-2, err


# ------------------------------------------------------------------------------------------ 2-D (PoseResNet) shapes
CONV2D_CASES = [(1, 1, 0, 64, 64), (1, 1, 0, 64, 256), (1, 1, 0, 256, 64), (1, 1, 0, 1024, 512), (1, 2, 0, 256, 512),
                (3, 1, 1, 64, 64), (3, 1, 1, 128, 128), (3, 1, 1, 256, 256), (1, 1, 0, 256, 15)]


@pytest.mark.parametrize("k,s,p,cin,cout", CONV2D_CASES)
def test_tc_conv2d_matches_float64_reference(k, s, p, cin, cout):
    torch.manual_seed(k * 100 + s * 10 + cin + cout)
    conv = nn.Conv2d(cin, cout, k, s, p, bias=(cout == 15))
    bn = rand_bn(nn.BatchNorm2d(cout), cin + cout)
    with torch.no_grad():
        conv.weight.copy_(bf16_round(conv.weight * (1.0 / (cin ** 0.5))))
    x = bf16_round(torch.randn(5, cin, 21, 13))
    with torch.no_grad():
        y0 = bn.double()(conv.double()(x.double()))
    res = bf16_round(torch.randn(*y0.shape))
    want = F.relu(y0 + res.double())
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, s, p, relu=1)
    assert pc.tc_supported()
    pitch = ops.round_up(cout, 16)
    xcl = ops.to_channel_last(x.unsqueeze(2).to(DEV), c_pitch=cin, dtype=torch.bfloat16)
    res_cl = ops.to_channel_last(res.unsqueeze(2).to(DEV), c_pitch=pitch, dtype=torch.float32)
    y = pc(xcl, residual=res_cl, out_pitch=pitch, out_dtype=torch.float32)
    got = ops.to_channel_first(y, cout)[:, :, 0].cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale, float((got - want).abs().max()) / scale


@pytest.mark.parametrize("cin", [256, 2048])
def test_tc_deconv2d_k4s2_matches_float64_reference(cin):
    torch.manual_seed(cin)
    ct = nn.ConvTranspose2d(cin, 256, 4, 2, 1, bias=False)
    bn = rand_bn(nn.BatchNorm2d(256), cin)
    with torch.no_grad():
        ct.weight.copy_(bf16_round(ct.weight))
    x = bf16_round(torch.randn(3, cin, 12, 9))
    with torch.no_grad():
        want = F.relu(bn.double()(ct.double()(x.double())))
    ct, bn = ct.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(ct.weight, None, bn, 2, 1, transposed=True, relu=1)
    assert pc.tc_supported()
    xcl = ops.to_channel_last(x.unsqueeze(2).to(DEV), c_pitch=cin, dtype=torch.bfloat16)
    y = pc(xcl, out_pitch=256, out_dtype=torch.float32)
    got = ops.to_channel_first(y, 256)[:, :, 0].cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 2e-5 * scale, float((got - want).abs().max()) / scale


def test_simt_conv_bf16_storage_strided():
    """3x3 stride-2 and the 7x7 stride-2 stem stay on the float32-math SIMT kernel with bf16 storage."""
    torch.manual_seed(3)
    conv = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
    bn = rand_bn(nn.BatchNorm2d(64), 9)
    with torch.no_grad():
        conv.weight.copy_(bf16_round(conv.weight))
    x = bf16_round(torch.randn(2, 64, 17, 11))
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())))
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, None, bn, 2, 1, relu=1)
    assert not pc.tc_supported()
    xcl = ops.to_channel_last(x.unsqueeze(2).to(DEV), c_pitch=64, dtype=torch.bfloat16)
    y = pc(xcl)
    assert y.dtype == torch.bfloat16
    got = ops.to_channel_first(y, 64, dtype=torch.float32)[:, :, 0].cpu().double()
    assert float((got - want).abs().max()) <= 6e-3 * float(want.abs().max())


def test_pose_resnet_bf16_mode_vs_reference_golden(golden):
    from selfpose3d_b200.config import default_config
    from selfpose3d_b200.models import pose_resnet
    g = golden("pose_resnet50")
    net = pose_resnet.get_pose_net(default_config(), is_train=False)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(g["seed"])), strict=True)
    ops.set_volume_dtype(torch.bfloat16)
    try:
        y = net.to(DEV).eval()(torch.from_numpy(g["x"]).to(DEV)).cpu().numpy()
    finally:
        ops.set_volume_dtype(torch.float32)
    err = float(np.abs(y - g["y"]).max()) / float(np.abs(g["y"]).max())
    print("PoseResNet-50 bf16 mode: max error / output range = %.3g" % err)
    assert err < 5e-2, err
