"""GPU tests of the term-pair activation format (``SP3D_BF16X2`` / ``ops.SplitAct``) of the float32-faithful tensor-core
mode: the convolution epilogue, the max-pool, the space-to-depth pass and the un-projection write float32 results
directly as the two bf16 term planes the next convolution reads (no ``sp3d_split_bf16`` pass between layers).
Every kernel is compared with a float64 CPU evaluation of the reference's layer on the un-rounded operands."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

from selfpose3d_b200 import ops, synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import project_layer  # noqa: E402
from test_gpu_tensorcore import rand_bn  # noqa: E402

DEV = "cuda:0"
DEFAULT_F32_CONV = ops.float32_conv()   # conftest: "simt" for the modules written against the float32 FMA kernels
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 5e-5     # of the output range, as tests/test_gpu_split.py


def note(line):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "pair_mode_errors.txt"), "a") as f:
            f.write(line + "\n")
    print(line)


@pytest.fixture(autouse=True)
def split_mode():
    ops.set_volume_dtype(torch.float32)
    ops.set_float32_conv("bf16x3")
    yield
    ops.set_float32_conv(DEFAULT_F32_CONV)


def to_pair(x, channels=None):
    """[N,C,*sp] float32 CPU -> SplitAct on the device."""
    if x.dim() == 4:
        x = x.unsqueeze(2)
    C = int(x.shape[1]) if channels is None else channels
    return ops.split_act(ops.to_channel_last(x.to(DEV)), C)


def from_pair(y, channels, nd=3):
    out = ops.to_channel_first(ops.merge_act(y, channels), channels).cpu().double()
    return out[:, :, 0] if nd == 2 else out


@pytest.mark.parametrize("C", [15, 16, 32, 64, 256])
def test_split_merge_roundtrip(C):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, C, 3, 5, 6, generator=g) * torch.exp(3 * torch.randn(2, C, 3, 5, 6, generator=g))
    p = to_pair(x)
    assert p.planes.shape == (2, 2, 3, 5, 6, ops.split_pitch(C))
    hi, lo = ops.bf16_terms(x, 2)
    got = from_pair(p, C)
    assert torch.equal(got.float(), hi + lo)
    assert not p.planes[..., C:].any()


# (z extents that are multiples of 32 take the 4-fold z-folded stem kernel with its 5-slot halo slice ring)
CONV_CASES = [(7, 15, 16, (6, 20, 12), False), (7, 1, 16, (6, 20, 12), False), (7, 15, 16, (5, 17, 34), False),
              (7, 15, 16, (6, 20, 32), False), (7, 15, 16, (9, 18, 64), False), (7, 15, 16, (13, 35, 32), False),
              (3, 16, 32, (5, 17, 32), True), (3, 32, 32, (6, 20, 16), True), (3, 32, 32, (6, 20, 12), True),
              (3, 32, 64, (6, 20, 12), True), (3, 64, 64, (6, 20, 12), True), (3, 64, 128, (6, 20, 12), False),
              (3, 128, 128, (6, 20, 12), True), (1, 16, 32, (6, 20, 12), False), (1, 32, 64, (6, 20, 12), False),
              (1, 64, 128, (5, 17, 9), False)]


@pytest.mark.parametrize("k,cin,cout,shape,with_res", CONV_CASES)
def test_pair_conv3d(k, cin, cout, shape, with_res):
    torch.manual_seed(k * 1000 + cin * 10 + cout + shape[2])
    conv = nn.Conv3d(cin, cout, k, 1, k // 2)
    bn = rand_bn(nn.BatchNorm3d(cout), cin + cout)
    x = torch.randn(2, cin, *shape)
    res = torch.randn(2, cout, *shape) if with_res else None
    with torch.no_grad():
        want = bn.double()(conv.double()(x.double()))
        want = F.relu(want + res.double() if with_res else want)
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, k // 2, relu=1)
    xin, rin = to_pair(x), (to_pair(res) if with_res else None)
    before = ops._lib.launch_count
    y = pc(xin, residual=rin)
    launches = ops._lib.launch_count - before
    assert isinstance(y, ops.SplitAct) and y.planes.shape[-1] == ops.split_pitch(cout)
    got = from_pair(y, cout)
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("pair conv3d k%d %d->%d %s res=%d: %.3g (%d launches)" % (k, cin, cout, shape, with_res, err, launches))
    assert err <= TOL, err
    assert launches <= 2          # the convolution (+ the tap stacking of the 1-channel stem): no split / merge pass
    assert not y.planes[..., cout:].any()
    # float32 out of the same input (the nets' output layers)
    yf = pc(xin, residual=ops.merge_act(to_pair(res), cout) if with_res else None, out_dtype=torch.float32)
    gotf = ops.to_channel_first(yf, cout).cpu().double()
    assert float((gotf - want).abs().max()) / float(want.abs().max()) <= TOL


@pytest.mark.parametrize("shape", [(4, 16, 8), (3, 20, 12)])
@pytest.mark.parametrize("cin,cout", [(128, 64), (64, 32)])
def test_pair_transposed_conv3d(cin, cout, shape):
    torch.manual_seed(cin + shape[0])
    ct, bn = nn.ConvTranspose3d(cin, cout, 2, 2), rand_bn(nn.BatchNorm3d(cout), cin)
    x = torch.randn(2, cin, *shape)
    skip = torch.randn(2, cout, *[2 * s for s in shape])
    with torch.no_grad():
        want = F.relu(bn.double()(ct.double()(x.double()))) + skip.double()
    ct, bn = ct.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(ct.weight, ct.bias, bn, 2, 0, transposed=True, relu=2)
    y = pc(to_pair(x), residual=to_pair(skip))
    got = from_pair(y, cout)
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("pair convT3d %d->%d %s: %.3g" % (cin, cout, shape, err))
    assert err <= TOL, err


CONV2D_CASES = [(1, 1, 0, 64, 256, True), (1, 1, 0, 1024, 512, False), (1, 2, 0, 256, 512, False),
                (3, 1, 1, 128, 128, False), (3, 1, 1, 64, 64, False), (1, 1, 0, 2048, 512, True)]


@pytest.mark.parametrize("k,s,p,cin,cout,with_res", CONV2D_CASES)
def test_pair_conv2d(k, s, p, cin, cout, with_res):
    torch.manual_seed(k * 100 + s * 10 + cin + cout)
    conv = nn.Conv2d(cin, cout, k, s, p, bias=False)
    bn = rand_bn(nn.BatchNorm2d(cout), cin + cout)
    x = torch.randn(3, cin, 18, 12)
    with torch.no_grad():
        want = bn.double()(conv.double()(x.double()))
        res = torch.randn(*want.shape) if with_res else None
        want = F.relu(want + res.double() if with_res else want)
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(conv.weight, None, bn, s, p, relu=1)
    y = pc(to_pair(x), residual=to_pair(res) if with_res else None)
    assert isinstance(y, ops.SplitAct)
    got = from_pair(y, cout, nd=2)
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("pair conv2d k%d s%d %d->%d: %.3g" % (k, s, cin, cout, err))
    assert err <= TOL, err


def test_pair_deconv2d_k4s2():
    torch.manual_seed(9)
    ct, bn = nn.ConvTranspose2d(256, 256, 4, 2, 1, bias=False), rand_bn(nn.BatchNorm2d(256), 3)
    x = torch.randn(2, 256, 9, 7)
    with torch.no_grad():
        want = F.relu(bn.double()(ct.double()(x.double())))
    ct, bn = ct.float().to(DEV), bn.float().to(DEV)
    pc = ops.PackedConv(ct.weight, None, bn, 2, 1, transposed=True, relu=1)
    y = pc(to_pair(x))
    got = from_pair(y, 256, nd=2)
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("pair deconv2d k4s2 256->256: %.3g" % err)
    assert err <= TOL, err


@pytest.mark.parametrize("k,p,cin,cout,hw", [(3, 1, 128, 128, (20, 12)), (3, 1, 512, 512, (12, 10)), (7, 3, 3, 64, (36, 28))])
def test_pair_stride2_space_to_depth(k, p, cin, cout, hw):
    """The stride-2 convolutions of PoseResNet (3x3 in layers 2-4 on term pairs, the 7x7 stem straight from the float32
    NCHW image) through the space-to-depth form on split operands."""
    torch.manual_seed(k + cin)
    conv, bn = nn.Conv2d(cin, cout, k, 2, p, bias=False), rand_bn(nn.BatchNorm2d(cout), cin)
    x = torch.randn(2, cin, *hw)
    with torch.no_grad():
        want = F.relu(bn.double()(conv.double()(x.double())))
    conv, bn = conv.float().to(DEV), bn.float().to(DEV)
    sc = ops.S2DConv(conv.weight, bn, p, relu=1)
    if cin == 3:
        xd = x.to(DEV)
        y = sc.call_split(xd, xd.stride(), 2, hw[0], hw[1])
    else:
        y = sc.call_split(to_pair(x))
    got = from_pair(y, cout, nd=2)
    err = float((got - want).abs().max()) / float(want.abs().max())
    note("pair s2d conv k%d/s2 %d->%d: %.3g" % (k, cin, cout, err))
    assert err <= TOL, err


@pytest.mark.parametrize("k,s,p,C,shape", [([2, 2, 2], [2, 2, 2], [0, 0, 0], 32, (8, 12, 6)),
                                           ([1, 3, 3], [1, 2, 2], [0, 1, 1], 64, (1, 17, 12))])
def test_pair_maxpool_exact(k, s, p, C, shape):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, C, *shape, generator=g)
    hi, lo = ops.bf16_terms(x, 2)
    want = F.max_pool3d(hi + lo, k, s, p)
    y = ops.maxpool(to_pair(x), C, k, s, p)
    got = from_pair(y, C).float()
    assert torch.equal(got, want)


@pytest.mark.parametrize("C,cube", [(15, (16, 12, 8)), (1, (20, 20, 8))])
def test_unproject_pair_output_equals_float32_form(C, cube):
    """K1 with out_dtype SP3D_BF16X2: the float32 form's value, split into its two bf16 terms (bit-exact)."""
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [72, 96], [18, 24]
    layer = project_layer.ProjectLayer(cfg)
    cams = synthetic.ring_cameras(5, seed=0)
    meta = synthetic.make_meta(cams, 2, (72, 96))
    rs = np.random.RandomState(C)
    hms = [torch.from_numpy(rs.rand(2, C, 24, 18).astype(np.float32)).to(DEV) for _ in range(5)]
    table = ops.pack_cameras(meta, cfg.NETWORK.IMAGE_SIZE).to(DEV)
    centers = torch.tensor([[0.0, -500.0, 800.0], [300.0, -200.0, 900.0], [-900.0, 100.0, 700.0]], device=DEV)
    sample = torch.tensor([0, 1, 1], dtype=torch.int32, device=DEV)
    f32, _ = layer.project_cl(hms, table, centers, False, [2000.0, 2000.0, 2000.0], cube, cube_sample=sample)
    pair, _ = layer.project_cl(hms, table, centers, False, [2000.0, 2000.0, 2000.0], cube, cube_sample=sample, dtype="split")
    assert isinstance(pair, ops.SplitAct) and pair.planes.shape[-1] == 16
    hi, lo = ops.bf16_terms(f32[..., :C].cpu(), 2)
    assert torch.equal(pair.planes[0, ..., :C].cpu().float(), hi)
    assert torch.equal(pair.planes[1, ..., :C].cpu().float(), lo)
    assert not pair.planes[..., C:].any()
    assert float(f32.max()) > 0.1


# ---- CTA pairs sharing the weight stream (clusters of two, multicast TMA): same bits as the single-CTA launches
# (k, cin, cout, N cubes, spatial shape): each case is large enough for the pair form (>= 4 items per SM); the 422-long
# one has an ODD item count, so one CTA of a pair owns an item less than its twin and walks that item's stages unread
PAIR_CTA_CASES = [(3, 64, 64, 8, (32, 32, 32), True), (3, 32, 32, 4, (32, 32, 64), True), (3, 16, 32, 4, (32, 32, 64), False),
                  (3, 128, 128, 24, (16, 16, 16), True), (7, 15, 16, 3, (32, 32, 64), False),
                  (3, 64, 64, 3, (422, 16, 8), True), (3, 32, 32, 3, (422, 16, 16), True)]


@pytest.mark.parametrize("k,cin,cout,n,shape,with_res", PAIR_CTA_CASES)
def test_cta_pair_weight_multicast_is_bit_exact(k, cin, cout, n, shape, with_res):
    torch.manual_seed(k + cin + cout + n)
    conv = nn.Conv3d(cin, cout, k, 1, k // 2).to(DEV)
    bn = rand_bn(nn.BatchNorm3d(cout), cin + cout).float().to(DEV)
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, k // 2, relu=1)
    xin = ops.split_act(torch.randn(n, *shape, ops.split_pitch(cin), device=DEV), cin)
    rin = ops.split_act(torch.randn(n, *shape, ops.split_pitch(cout), device=DEV), cout) if with_res else None
    outs = []
    try:
        for pair in (0, 1, 1):
            ops._lib.load().sp3d_debug_conv_pair(pair)
            outs.append(pc(xin, residual=rin).planes.clone())
            torch.cuda.synchronize()
    finally:
        ops._lib.load().sp3d_debug_conv_pair(0)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert outs[0].float().abs().max() > 0
