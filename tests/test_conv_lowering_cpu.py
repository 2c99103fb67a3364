"""CPU check of the host-side convolution lowering (``ops.PackedConv``): BatchNorm folding, weight
packing, tap offsets and the per-phase decomposition of transposed convolutions are validated by
executing the exact ``sp3d_conv_fwd`` argument semantics (include/sp3d.h) with a small torch
emulator and comparing with ``torch.nn.functional`` on the reference-shaped parameters."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from selfpose3d_b200 import ops


def emulate_conv_launch(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0,
                        tap_step, ostride, ooffset, relu, algo=0, cin_real=None):
    N, D, H, W, _ = x.shape
    dims = (D, H, W)
    acc = torch.zeros((N,) + tuple(out_grid) + (cout,), dtype=torch.float64)
    t = 0
    for td in range(ksize[0]):
        for th in range(ksize[1]):
            for tw in range(ksize[2]):
                tt = (td, th, tw)
                idx, ok = [], []
                for a in range(3):
                    i = torch.arange(out_grid[a]) * stride[a] + tap_off0[a] + tt[a] * tap_step[a]
                    ok.append((i >= 0) & (i < dims[a]))
                    idx.append(i.clamp(0, dims[a] - 1))
                xs = x[:, idx[0]][:, :, idx[1]][:, :, :, idx[2]][..., :cin].double()
                m = (ok[0][:, None, None] & ok[1][None, :, None] & ok[2][None, None, :]).double()
                acc += torch.einsum("ndhwc,co->ndhwo", xs * m[None, ..., None], weight[t, :cin, :cout].double())
                t += 1
    v = acc
    if scale is not None:
        v = v * scale.double()
    if shift is not None:
        v = v + shift.double()
    sl = tuple(slice(ooffset[a], ooffset[a] + ostride[a] * out_grid[a], ostride[a]) for a in range(3))
    if relu == 2:
        v = v.clamp_min(0)
    if residual is not None:
        v = v + residual[(slice(None),) + sl][..., :cout].double()
    if relu == 1:
        v = v.clamp_min(0)
    out[(slice(None),) + sl + (slice(0, cout),)] = v.float()
    out[(slice(None),) + sl + (slice(cout, None),)] = 0


@pytest.fixture(autouse=True)
def _patch(monkeypatch):
    monkeypatch.setattr(ops, "conv_launch", emulate_conv_launch)


def cl(x):   # [N,C,*sp] -> channel-last 5-D with pitch rounded to 4
    if x.dim() == 4:
        x = x.unsqueeze(2)
    N, C = x.shape[:2]
    out = torch.zeros((N,) + tuple(x.shape[2:]) + (ops.round_up(C, 4),))
    out[..., :C] = x.permute(0, 2, 3, 4, 1)
    return out


def cf(y, C, nd):
    y = y[..., :C].permute(0, 4, 1, 2, 3)
    return y[:, :, 0] if nd == 2 else y


def rand_bn(bn):
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
    return bn.eval()


@pytest.mark.parametrize("k,s,p,cin,cout", [(7, 1, 3, 3, 16), (3, 1, 1, 5, 6), (1, 1, 0, 6, 15), (3, 2, 1, 4, 8)])
def test_conv3d_lowering(k, s, p, cin, cout):
    torch.manual_seed(0)
    conv, bn = nn.Conv3d(cin, cout, k, s, p), rand_bn(nn.BatchNorm3d(cout))
    x = torch.randn(2, cin, 6, 5, 7)
    want = F.relu(bn(conv(x)))
    got = cf(ops.PackedConv(conv.weight, conv.bias, bn, s, p, relu=1)(cl(x)), cout, 3)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k,s,p,cin,cout,hw", [(7, 2, 3, 3, 8, (13, 10)), (3, 2, 1, 8, 8, (9, 12)), (1, 2, 0, 8, 12, (9, 7))])
def test_conv2d_lowering_with_residual(k, s, p, cin, cout, hw):
    torch.manual_seed(1)
    conv, bn = nn.Conv2d(cin, cout, k, s, p, bias=False), rand_bn(nn.BatchNorm2d(cout))
    x = torch.randn(2, cin, *hw)
    y0 = bn(conv(x))
    res = torch.randn_like(y0)
    want = F.relu(y0 + res)
    got = cf(ops.PackedConv(conv.weight, None, bn, s, p, relu=1)(cl(x), residual=cl(res)), cout, 2)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


def test_conv_transpose3d_k2s2_lowering_relu_before_skip():
    torch.manual_seed(2)
    ct, bn = nn.ConvTranspose3d(6, 4, 2, 2), rand_bn(nn.BatchNorm3d(4))
    x = torch.randn(2, 6, 3, 4, 2)
    skip = torch.randn(2, 4, 6, 8, 4)
    want = F.relu(bn(ct(x))) + skip
    got = cf(ops.PackedConv(ct.weight, ct.bias, bn, 2, 0, transposed=True, relu=2)(cl(x), residual=cl(skip)), 4, 3)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("hw", [(3, 4), (5, 2)])
def test_conv_transpose2d_k4s2p1_lowering(hw):
    torch.manual_seed(3)
    ct, bn = nn.ConvTranspose2d(8, 6, 4, 2, 1, bias=False), rand_bn(nn.BatchNorm2d(6))
    x = torch.randn(2, 8, *hw)
    want = F.relu(bn(ct(x)))
    got = cf(ops.PackedConv(ct.weight, None, bn, 2, 1, transposed=True, relu=1)(cl(x)), 6, 2)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


def test_output_pitch_one_and_padding_lanes_zero():
    torch.manual_seed(4)
    conv = nn.Conv3d(8, 1, 1)
    x = torch.randn(1, 8, 4, 4, 4)
    y = ops.PackedConv(conv.weight, conv.bias, None, 1, 0)(cl(x), out_pitch=1)
    torch.testing.assert_close(y[..., 0], conv(x)[:, 0], rtol=1e-5, atol=1e-6)
    conv = nn.Conv3d(8, 15, 1)
    y = ops.PackedConv(conv.weight, conv.bias, None, 1, 0)(cl(x))
    assert y.shape[-1] == 16 and not y[..., 15].any()


# ---------------------------------------------------------------------------------------------- tcgen05 packing
def emulate_any_launch(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0,
                       tap_step, ostride, ooffset, relu, algo=0, cin_real=None, cout_pitch_w=None, fused_phases=False,
                       zfold=0, split_terms=0, pair_out=False):
    """Dispatch on the packing: the tensor-core launch carries [n_tiles, n_chunks, taps, N, chunk] bf16 weights;
    `fused_phases` and `zfold` follow the sp3d_conv_args field descriptions in include/sp3d.h."""
    res = None if residual is None else residual.float()
    if algo == 1:
        nt, nc, taps, n, chunk = weight.shape
        weight = weight.float().permute(2, 1, 4, 0, 3).reshape(taps, nc * chunk, nt * n)
        assert cout_pitch_w == n and cin == nc * chunk
    if fused_phases:
        # ksize (1,1,1), ostride (2,2,2): GEMM rows (px, py, pz, co); one 1x1 launch per phase is the same thing
        assert list(ksize) == [1, 1, 1] and list(ostride) == [2, 2, 2] and weight.shape[2] == 8 * cout
        for ph in range(8):
            off = [(ph >> 2) & 1, (ph >> 1) & 1, ph & 1]
            emulate_conv_launch(x.float(), weight[:, :, ph * cout:(ph + 1) * cout], scale, shift, res, out, cin, cout,
                                out_grid, ksize, stride, tap_off0, tap_step, ostride, off, relu)
        return
    if zfold and zfold > 1:
        # per (kd, kh): k + F - 1 windows e of rows (ro, co) holding tap kw = e - ro, zero rows elsewhere
        F_, k = zfold, ksize[2]
        pitch = out.shape[-1]
        assert cout_pitch_w == F_ * pitch and weight.shape[2] == F_ * pitch
        w = weight.reshape(ksize[0], ksize[1], k + F_ - 1, cin, F_, pitch)
        std = torch.zeros(ksize[0], ksize[1], k, cin, pitch)
        for e in range(k + F_ - 1):
            for ro in range(F_):
                kw = e - ro
                if 0 <= kw < k:
                    if ro == 0 or e - ro + 0 == kw:
                        if ro > 0:
                            assert torch.equal(std[:, :, kw], w[:, :, e, :, ro])    # every output position sees the same tap
                        std[:, :, kw] = w[:, :, e, :, ro]
                else:
                    assert not w[:, :, e, :, ro].any()
        weight = std.reshape(-1, cin, pitch)
    emulate_conv_launch(x.float(), weight, scale, shift, res, out, cin, cout, out_grid, ksize, stride, tap_off0,
                        tap_step, ostride, ooffset, relu)


def emulate_space_to_depth(x, channels, strides, n, h, w, dst_pitch):
    """include/sp3d.h, sp3d_s2d_args: dst[n, y', x', (py*2+px)*C + c] = src[n, c, 2y'+py, 2x'+px]."""
    src = torch.as_strided(x, (n, channels, h, w), strides).float()
    out = torch.zeros(n, 1, h // 2, w // 2, dst_pitch)
    for py in range(2):
        for px in range(2):
            q = py * 2 + px
            out[:, 0, :, :, q * channels:(q + 1) * channels] = src[:, :, py::2, px::2].permute(0, 2, 3, 1)
    return out.to(torch.bfloat16)


def cl16(x):   # bf16 channel-last, pitch rounded to 16
    if x.dim() == 4:
        x = x.unsqueeze(2)
    N, C = x.shape[:2]
    out = torch.zeros((N,) + tuple(x.shape[2:]) + (ops.round_up(C, 16),))
    out[..., :C] = x.permute(0, 2, 3, 4, 1)
    return out.to(torch.bfloat16)


def bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("k,p,cin,cout,hw", [(3, 1, 64, 64, (8, 6)), (7, 3, 3, 64, (12, 10))])
def test_space_to_depth_stride2_lowering(monkeypatch, k, p, cin, cout, hw):
    """ops.S2DConv: a stride-2 convolution as a stride-1 one over the 2x2 space-to-depth tensor (weight re-indexing,
    asymmetric tap origin, channel order)."""
    monkeypatch.setattr(ops, "conv_launch", emulate_any_launch)
    monkeypatch.setattr(ops, "space_to_depth", emulate_space_to_depth)
    torch.manual_seed(11)
    conv, bn = nn.Conv2d(cin, cout, k, 2, p, bias=False), rand_bn(nn.BatchNorm2d(cout))
    conv.weight.data = bf(conv.weight.data)
    x = bf(torch.randn(2, cin, *hw))
    want = F.relu(bn(conv(x)))
    sc = ops.S2DConv(conv.weight, bn, p, relu=1)
    y = sc(x, x.stride(), 2, hw[0], hw[1])
    got = y[:, 0, :, :, :cout].permute(0, 3, 1, 2).float()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 1e-2 * float(want.abs().max())    # bf16 output


def emulate_stack_x_shifts(x, taps, pad):
    """include/sp3d.h, sp3d_stack_args: dst[n, x, y, z, j] = src[n, x + j - pad, y, z, 0], zero outside."""
    N, X, Y, Z, _ = x.shape
    out = torch.zeros(N, X, Y, Z, 16)
    v = x[..., 0].float()
    for j in range(taps):
        lo, hi = max(0, pad - j), min(X, X + pad - j)
        out[:, lo:hi, :, :, j] = v[:, lo + j - pad:hi + j - pad]
    return out.to(torch.bfloat16)


def test_one_channel_stem_tap_stacking(monkeypatch):
    """Root-net stem (1 -> 16 channels, 7^3): x taps stacked into channels + 1 x 7 x 7 z-folded packing."""
    monkeypatch.setattr(ops, "conv_launch", emulate_any_launch)
    monkeypatch.setattr(ops, "stack_x_shifts", emulate_stack_x_shifts)
    torch.manual_seed(3)
    conv, bn = nn.Conv3d(1, 16, 7, 1, 3), rand_bn(nn.BatchNorm3d(16))
    conv.weight.data = bf(conv.weight.data)
    x = bf(torch.randn(2, 1, 5, 4, 6))
    want = F.relu(bn(conv(x)))
    pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, 3, relu=1)
    assert pc._tc_stack_ok(6, 16)
    got = cf(pc(cl16(x), out_pitch=16, out_dtype=torch.float32), 16, 3)
    assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())


@pytest.mark.parametrize("case", ["1x1", "1x1s2", "3x3", "deconv", "3d_k3", "3d_convT", "3d_stem_zfold"])
def test_tensorcore_lowering(monkeypatch, case):
    monkeypatch.setattr(ops, "conv_launch", emulate_any_launch)
    torch.manual_seed(7)
    if case in ("1x1", "1x1s2", "3x3"):
        k, s, p = {"1x1": (1, 1, 0), "1x1s2": (1, 2, 0), "3x3": (3, 1, 1)}[case]
        conv, bn = nn.Conv2d(64, 256, k, s, p, bias=False), rand_bn(nn.BatchNorm2d(256))
        conv.weight.data = bf(conv.weight.data)
        x = bf(torch.randn(3, 64, 9, 7))
        want = F.relu(bn(conv(x)))
        pc = ops.PackedConv(conv.weight, None, bn, s, p, relu=1)
        assert pc.tc_supported()
        got = cf(pc(cl16(x), out_dtype=torch.float32), 256, 2)
    elif case == "deconv":
        ct, bn = nn.ConvTranspose2d(64, 256, 4, 2, 1, bias=False), rand_bn(nn.BatchNorm2d(256))
        ct.weight.data = bf(ct.weight.data)
        x = bf(torch.randn(2, 64, 5, 3))
        want = F.relu(bn(ct(x)))
        pc = ops.PackedConv(ct.weight, None, bn, 2, 1, transposed=True, relu=1)
        assert pc.tc_supported()
        got = cf(pc(cl16(x), out_dtype=torch.float32), 256, 2)
    elif case == "3d_k3":
        conv, bn = nn.Conv3d(128, 128, 3, 1, 1), rand_bn(nn.BatchNorm3d(128))
        conv.weight.data = bf(conv.weight.data)
        x = bf(torch.randn(1, 128, 4, 5, 3))
        want = F.relu(bn(conv(x)))
        pc = ops.PackedConv(conv.weight, conv.bias, bn, 1, 1, relu=1)
        got = cf(pc(cl16(x), out_dtype=torch.float32), 128, 3)
    else:
        ct, bn = nn.ConvTranspose3d(128, 64, 2, 2), rand_bn(nn.BatchNorm3d(64))
        ct.weight.data = bf(ct.weight.data)
        x = bf(torch.randn(1, 128, 3, 2, 4))
        want = F.relu(bn(ct(x)))
        pc = ops.PackedConv(ct.weight, ct.bias, bn, 2, 0, transposed=True, relu=2)
        got = cf(pc(cl16(x), out_dtype=torch.float32), 64, 3)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------- split-operand mode
def emulate_split_bf16(x, channels, c_block, blocks):
    """include/sp3d.h, sp3d_split_args: plane s = term s of the bf16 expansion, channels >= C zero."""
    out = torch.zeros((blocks,) + tuple(x.shape[:-1]) + (c_block,))
    for s, t in enumerate(ops.bf16_terms(x[..., :channels], blocks)):
        out[s, ..., :channels] = t
    return out.to(torch.bfloat16)


def emulate_split_launch(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0,
                         tap_step, ostride, ooffset, relu, algo=0, cin_real=None, cout_pitch_w=None,
                         fused_phases=False, zfold=0, split_terms=0, pair_out=False):
    """SP3D_CONV_TC_BF16X3 (include/sp3d.h, split_terms): `x` is plane 0 of the term planes [S][N,D,H,W,cin]; K block
    b multiplies activation plane SPLIT_PAIRS[b][0] with weight block b, all blocks accumulate into one GEMM -- i.e.
    the plain tensor-core launch on the K-concatenated operands.  `pair_out` (out_dtype SP3D_BF16X2): `out` and
    `residual` are plane 0 of two-plane bf16 tensors holding float32 values as term pairs."""
    assert algo == 2 and split_terms in (2, 3, 6)
    if split_terms == 2:
        # the 3 pairs in two K blocks: per tile [chunk][tap][w0 rows | w1 rows] then [chunk][tap][w0 rows]
        nt, rows, chunk = weight.shape
        n, nc = cout_pitch_w, cin // chunk
        taps = rows // (3 * n * nc)
        assert rows == 3 * n * nc * taps and cin == nc * chunk
        wide = weight[:, :2 * n * nc * taps].reshape(nt, nc, taps, 2 * n, chunk)
        narrow = weight[:, 2 * n * nc * taps:].reshape(nt, nc, taps, n, chunk)
        assert torch.equal(wide[:, :, :, :n], narrow)
        weight = torch.stack([wide[:, :, :, :n], wide[:, :, :, n:], narrow], 1)      # K blocks x0 w0, x0 w1, x1 w0
        pairs = ((0, 0), (0, 1), (1, 0))
    else:
        pairs = ops.SPLIT_PAIRS[split_terms]
    nt, kb, nc, taps, n, chunk = weight.shape
    assert kb == len(pairs) and cin == nc * chunk and x.shape[-1] == cin

    def planes_of(t, count):
        return torch.as_strided(t, (count,) + tuple(t.shape), (t.numel(),) + tuple(t.stride()), t.storage_offset())

    planes = planes_of(x, max(a for a, _ in pairs) + 1)
    xcat = torch.cat([planes[a] for a, _ in pairs], -1)
    target = out
    if pair_out:
        assert split_terms in (2, 3) and out.dtype == torch.bfloat16 and out.is_contiguous()
        out_planes = planes_of(out, 2)
        target = out_planes[0].float() + out_planes[1].float()      # other phases' results stay what they are
        if residual is not None:
            rp = planes_of(residual, 2)
            residual = rp[0].float() + rp[1].float()
    emulate_any_launch(xcat, weight.reshape(nt, kb * nc, taps, n, chunk), scale, shift, residual, target, kb * cin, cout,
                       out_grid, ksize, stride, tap_off0, tap_step, ostride, ooffset, relu, algo=1,
                       cout_pitch_w=cout_pitch_w, fused_phases=fused_phases, zfold=zfold)
    if pair_out:
        hi, lo = ops.bf16_terms(target, 2)
        out_planes[0].copy_(hi)
        out_planes[1].copy_(lo)


def emulate_stack_planes(x, taps, pad):
    return emulate_stack_x_shifts(x, taps, pad)


@pytest.mark.parametrize("terms,tol", [(3, 4e-5), (6, 2e-6)])
@pytest.mark.parametrize("case", ["3d_k3_res", "3d_k3_zfold", "3d_k7_zfold", "3d_k7_zfold4", "3d_k7_stack", "3d_convT_fused",
                                  "2d_1x1s2", "2d_deconv"])
def test_split_operand_lowering(monkeypatch, case, terms, tol):
    """float32 activations / weights through the tensor-core packings (plain, z-folded, tap-stacked, fused transposed,
    per-phase transposed) as sums of bf16 terms: the result must be float32-faithful (3 term pairs: ~2^-16; 6:
    float32 rounding level) for un-rounded operands."""
    monkeypatch.setattr(ops, "conv_launch", emulate_split_launch)
    monkeypatch.setattr(ops, "split_bf16", emulate_split_bf16)
    monkeypatch.setattr(ops, "stack_x_shifts", emulate_stack_planes)
    monkeypatch.setattr(ops, "_F32_CONV", "bf16x3" if terms == 3 else "bf16x6")
    torch.manual_seed(5)
    res = None
    if case == "3d_k3_res":
        conv, bn = nn.Conv3d(32, 64, 3, 1, 1), rand_bn(nn.BatchNorm3d(64))
        x, res = torch.randn(1, 32, 4, 5, 6), torch.randn(1, 64, 4, 5, 6)
        want = F.relu(bn.double()(conv.double()(x.double())) + res.double())
        pc = ops.PackedConv(conv.float().weight, conv.bias, bn.float(), 1, 1, relu=1)
        nd, cout = 3, 64
    elif case == "3d_k3_zfold":
        conv, bn = nn.Conv3d(16, 32, 3, 1, 1), rand_bn(nn.BatchNorm3d(32))
        x, res = torch.randn(1, 16, 3, 4, 16), torch.randn(1, 32, 3, 4, 16)
        want = F.relu(bn.double()(conv.double()(x.double())) + res.double())
        pc = ops.PackedConv(conv.float().weight, conv.bias, bn.float(), 1, 1, relu=1)
        assert pc._tc_zfold_ok(16, 32)
        nd, cout = 3, 32
    elif case in ("3d_k7_zfold", "3d_k7_zfold4", "3d_k7_stack"):
        cin = 1 if case == "3d_k7_stack" else 15
        conv, bn = nn.Conv3d(cin, 16, 7, 1, 3), rand_bn(nn.BatchNorm3d(16))
        # (a z extent that is a multiple of 32 takes the 4-fold z-fold in the 3-pair mode: rows of 4 positions)
        x = torch.rand(1, cin, 4, 8, 32 if case == "3d_k7_zfold4" else 6)
        want = F.relu(bn.double()(conv.double()(x.double())))
        pc = ops.PackedConv(conv.float().weight, conv.bias, bn.float(), 1, 3, relu=1)
        assert pc._tc_zfold_ok(int(x.shape[-1]), 16) if cin == 15 else pc._tc_stack_ok(6, 16)
        nd, cout = 3, 16
    elif case == "3d_convT_fused":
        ct, bn = nn.ConvTranspose3d(128, 64, 2, 2), rand_bn(nn.BatchNorm3d(64))
        x, res = torch.randn(1, 128, 3, 2, 4), torch.randn(1, 64, 6, 4, 8)
        want = F.relu(bn.double()(ct.double()(x.double()))) + res.double()
        pc = ops.PackedConv(ct.float().weight, ct.bias, bn.float(), 2, 0, transposed=True, relu=2)
        assert pc._tc_fused_ok(64, torch.float32)
        nd, cout = 3, 64
    elif case == "2d_1x1s2":
        conv, bn = nn.Conv2d(64, 256, 1, 2, 0, bias=False), rand_bn(nn.BatchNorm2d(256))
        x = torch.randn(3, 64, 9, 7)
        want = bn.double()(conv.double()(x.double()))
        pc = ops.PackedConv(conv.float().weight, None, bn.float(), 2, 0, relu=0)
        nd, cout = 2, 256
    else:
        ct, bn = nn.ConvTranspose2d(64, 256, 4, 2, 1, bias=False), rand_bn(nn.BatchNorm2d(256))
        x = torch.randn(2, 64, 5, 3)
        want = F.relu(bn.double()(ct.double()(x.double())))
        pc = ops.PackedConv(ct.float().weight, None, bn.float(), 2, 1, transposed=True, relu=1)
        nd, cout = 2, 256
    assert pc.tc_supported()
    seen, folds = [], []
    monkeypatch.setattr(ops, "conv_launch", lambda *a, **k: (seen.append(k.get("split_terms")), folds.append(k.get("zfold", 0)),
                                                             emulate_split_launch(*a, **k))[2])
    got = cf(pc(cl(x), residual=None if res is None else cl(res)), cout, nd).double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    assert err <= tol, err
    # the stems and the z-folded 3^3 layers take the 2-K-block form of the 3-pair mode (split_terms 2: ops.TC_WIDE_CASES)
    if terms == 3 and case in ("3d_k7_zfold", "3d_k7_zfold4", "3d_k7_stack", "3d_k3_zfold"):
        assert seen == [2], seen
        if case.startswith("3d_k7_zfold"):
            assert folds == [4 if case == "3d_k7_zfold4" else 2], folds
    else:
        assert all(t == terms for t in seen), seen


# ---------------------------------------------------------------------------------------------- input gradients
@pytest.mark.parametrize("case", ["3x3s2_even", "3x3s2_odd", "1x1s2", "7x7s2", "3x3s1", "deconv4s2", "3d_k3", "3d_convT2"])
def test_conv_dgrad_lowering(monkeypatch, case):
    """grad_ops.conv_dgrad: the input gradient as the forward kernel on the adjoint operator (flipped kernel for
    stride 1, transposed convolution with an explicit output extent for strided convolutions -- incl. tap-less
    phases of a 1x1 / stride-2 kernel --, strided convolution for transposed ones), against torch autograd."""
    from selfpose3d_b200 import grad_ops
    monkeypatch.setattr(ops, "conv_launch", emulate_conv_launch)
    monkeypatch.setattr(grad_ops, "_f32", lambda *a: None)
    torch.manual_seed(21)
    nd, transposed = 2, False
    if case.startswith("3x3s2"):
        conv, x = nn.Conv2d(8, 12, 3, 2, 1), torch.randn(2, 8, *((10, 8) if case.endswith("even") else (9, 7)))
    elif case == "1x1s2":
        conv, x = nn.Conv2d(8, 12, 1, 2, 0), torch.randn(2, 8, 9, 8)
    elif case == "7x7s2":
        conv, x = nn.Conv2d(3, 8, 7, 2, 3), torch.randn(1, 3, 14, 12)
    elif case == "3x3s1":
        conv, x = nn.Conv2d(8, 8, 3, 1, 1), torch.randn(2, 8, 6, 5)
    elif case == "deconv4s2":
        conv, x, transposed = nn.ConvTranspose2d(8, 6, 4, 2, 1), torch.randn(2, 8, 5, 4), True
    elif case == "3d_k3":
        conv, x, nd = nn.Conv3d(4, 6, 3, 1, 1), torch.randn(1, 4, 4, 5, 3), 3
    else:
        conv, x, nd, transposed = nn.ConvTranspose3d(8, 4, 2, 2), torch.randn(1, 8, 3, 2, 4), 3, True
    x.requires_grad_(True)
    y = conv(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    pc = ops.PackedConv(conv.weight, conv.bias, None, conv.stride[0], conv.padding[0], transposed=transposed, relu=0)
    xcl = cl(x.detach())
    torch.testing.assert_close(cf(pc(xcl), y.shape[1], nd), y.detach(), rtol=1e-4, atol=1e-5)
    gx = grad_ops.conv_dgrad(pc, cl(gy), out_pitch=xcl.shape[-1], in_dims=xcl.shape[1:4])
    assert gx.shape == xcl.shape
    torch.testing.assert_close(cf(gx, x.shape[1], nd), x.grad, rtol=1e-4, atol=1e-5)


def test_tc_plan_predicts_the_launches(monkeypatch):
    """PackedConv.tc_plan (used to decide whether the split-operand tensor-core path has a compiled kernel for a
    layer) against the launches PackedConv._call_tc really makes, over every layer shape of V2VNet and PoseResNet."""
    seen = []

    def record(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0, tap_step, ostride,
               ooffset, relu, algo=0, cin_real=None, cout_pitch_w=None, fused_phases=False, zfold=0, split_terms=0,
               pair_out=False):
        seen.append(ops.tc_case(ksize, cin, cout_pitch_w, zfold))

    monkeypatch.setattr(ops, "conv_launch", record)
    monkeypatch.setattr(ops, "split_bf16", emulate_split_bf16)
    monkeypatch.setattr(ops, "stack_x_shifts", lambda x, taps, pad: torch.zeros(tuple(x.shape[:-1]) + (16,), dtype=torch.bfloat16))
    monkeypatch.setattr(ops, "_F32_CONV", "bf16x3")
    layers = [(nn.Conv3d(15, 16, 7, 1, 3), (4, 8, 6), False), (nn.Conv3d(15, 16, 7, 1, 3), (4, 8, 32), False),
              (nn.Conv3d(1, 16, 7, 1, 3), (4, 8, 6), False),
              (nn.Conv3d(1, 16, 7, 1, 3), (4, 8, 7), False), (nn.Conv3d(16, 32, 3, 1, 1), (4, 4, 16), True),
              (nn.Conv3d(32, 32, 3, 1, 1), (4, 4, 16), True), (nn.Conv3d(32, 32, 3, 1, 1), (4, 4, 20), False),
              (nn.Conv3d(32, 64, 3, 1, 1), (4, 4, 8), False), (nn.Conv3d(64, 64, 3, 1, 1), (4, 4, 8), True),
              (nn.Conv3d(64, 128, 3, 1, 1), (4, 4, 4), False), (nn.Conv3d(128, 128, 3, 1, 1), (4, 4, 4), True),
              (nn.Conv3d(16, 32, 1), (4, 4, 8), False), (nn.Conv3d(32, 15, 1), (4, 4, 8), False),
              (nn.Conv3d(32, 1, 1), (4, 4, 8), False), (nn.ConvTranspose3d(128, 64, 2, 2), (2, 2, 2), True),
              (nn.ConvTranspose3d(64, 32, 2, 2), (2, 2, 4), True), (nn.Conv2d(64, 256, 1), (6, 5), True),
              (nn.Conv2d(256, 64, 1), (6, 5), False), (nn.Conv2d(64, 64, 3, 1, 1), (6, 5), False),
              (nn.Conv2d(512, 512, 3, 1, 1), (6, 5), False), (nn.Conv2d(256, 512, 1, 2), (6, 4), False),
              (nn.ConvTranspose2d(2048, 256, 4, 2, 1, bias=False), (3, 2), False), (nn.Conv2d(256, 15, 1), (6, 5), False)]
    for mod, sp, with_res in layers:
        transposed = isinstance(mod, (nn.ConvTranspose3d, nn.ConvTranspose2d))
        pc = ops.PackedConv(mod.weight, mod.bias, None, mod.stride[0], mod.padding[0], transposed=transposed, relu=0)
        x = torch.randn(1, mod.in_channels, *sp)
        xcl = cl(x)
        o = pc.out_shape(tuple(xcl.shape[1:4]))
        pitch = ops.round_up(pc.cout, 4)
        res = torch.zeros(1, o[0], o[1], o[2], pitch) if with_res else None
        assert pc.tc_supported()
        seen.clear()
        plan = pc.tc_plan(int(xcl.shape[3]), pitch, torch.float32, with_res)
        pc._call_tc(xcl, res, None, None, terms=3)
        assert seen == plan, (mod, sp, seen, plan)
        assert pc.tc_available(int(xcl.shape[3]), pitch, torch.float32, with_res), (mod, plan)
    # adjoint shapes of the training path that have no compiled kernel must be reported as such (-> SIMT)
    adj = ops.PackedConv(nn.Conv3d(64, 16, 3, 1, 1).weight, None, None, 1, 1, relu=0)      # no such layer: 3^3 64 -> 16
    assert adj.tc_supported() and not adj.tc_available(8, 16, torch.float32, False)
    adj = ops.PackedConv(nn.Conv3d(32, 16, 3, 1, 1).weight, None, None, 1, 1, relu=0)      # dgrad of the 3^3 16 -> 32 layer
    assert adj.tc_available(8, 16, torch.float32, False)
