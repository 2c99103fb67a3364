"""CPU check of the float32-faithful tensor-core mode's host wiring (``ops.use_split()``): evaluation-mode V2VNet and
PoseResNet keep their activations as ``ops.SplitAct`` term pairs (``SP3D_BF16X2``) from layer to layer -- the
convolution epilogue, the max-pool, the space-to-depth pass and the un-projection write the pairs, no ``sp3d_split_bf16``
pass runs between layers.  Every kernel entry point is replaced by a torch emulation of its documented semantics
(include/sp3d.h); the nets are compared with the oracle's float64 restatement of the reference's layers."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nets
from selfpose3d_b200 import ops, synthetic
from selfpose3d_b200.config import default_config
from selfpose3d_b200.models import pose_resnet, v2v_net
from test_conv_lowering_cpu import (emulate_any_launch, emulate_conv_launch, emulate_split_bf16, emulate_split_launch,
                                    emulate_stack_x_shifts)

CALLS = {"split": 0, "merge": 0, "pair_out": 0, "f32_out": 0, "simt": 0}


def _launch(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0, tap_step, ostride,
            ooffset, relu, algo=0, **kw):
    args = (x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0, tap_step, ostride,
            ooffset, relu)
    if algo == 2:
        CALLS["pair_out" if kw.get("pair_out") else "f32_out"] += 1
        return emulate_split_launch(*args, algo=algo, **kw)
    if algo == 1:
        return emulate_any_launch(*args, algo=algo, **kw)
    CALLS["simt"] += 1
    return emulate_conv_launch(*args, algo=algo, cin_real=kw.get("cin_real"))


def _split(x, channels, c_block, blocks):
    CALLS["split"] += 1
    return emulate_split_bf16(x, channels, c_block, blocks)


def _merge(x, channels, pitch=None):
    CALLS["merge"] += 1
    pitch = ops.round_up(channels, 4) if pitch is None else pitch
    out = torch.zeros(tuple(x.planes.shape[1:-1]) + (pitch,))
    out[..., :channels] = (x.planes[0].float() + x.planes[1].float())[..., :channels]
    return out


def _maxpool(x, channels, k, s, p):
    """sp3d_maxpool_fwd; SP3D_BF16X2: the maximum of plane 0 + plane 1, re-split."""
    pair = isinstance(x, ops.SplitAct)
    v = (x.planes[0].float() + x.planes[1].float()) if pair else x.float()
    y = F.max_pool3d(v.permute(0, 4, 1, 2, 3), k, s, p).permute(0, 2, 3, 4, 1).contiguous()
    if not pair:
        return y.to(x.dtype)
    hi, lo = ops.bf16_terms(y, 2)
    return ops.SplitAct(torch.stack([hi, lo]).to(torch.bfloat16).contiguous())


def _s2d(x, channels, strides, n, h, w, dst_pitch, pair=False):
    """sp3d_space_to_depth (dst_dtype SP3D_BF16X2 with `pair`)."""
    src = torch.as_strided(x, (n, channels, h, w), strides).float()
    out = torch.zeros(n, 1, h // 2, w // 2, dst_pitch)
    for py in range(2):
        for px in range(2):
            q = py * 2 + px
            out[:, 0, :, :, q * channels:(q + 1) * channels] = src[:, :, py::2, px::2].permute(0, 2, 3, 1)
    if not pair:
        return out.to(torch.bfloat16)
    hi, lo = ops.bf16_terms(out, 2)
    return torch.stack([hi, lo]).to(torch.bfloat16).contiguous()


def _cl(x, c_pitch=None, dtype=None):
    x = x.contiguous()
    N, C = x.shape[:2]
    out = torch.zeros((N,) + tuple(x.shape[2:]) + (ops.round_up(C, 4) if c_pitch is None else c_pitch,))
    out[..., :C] = x.movedim(1, -1)
    return out.to(dtype or x.dtype)


def _cf(x, channels, dtype=None):
    return x[..., :channels].movedim(-1, 1).contiguous().to(dtype or x.dtype)


@pytest.fixture
def split_mode(monkeypatch):
    monkeypatch.setattr(ops, "conv_launch", _launch)
    monkeypatch.setattr(ops, "split_bf16", _split)
    monkeypatch.setattr(ops, "merge_act", _merge)
    monkeypatch.setattr(ops, "maxpool", _maxpool)
    monkeypatch.setattr(ops, "space_to_depth", _s2d)
    monkeypatch.setattr(ops, "stack_x_shifts", emulate_stack_x_shifts)
    monkeypatch.setattr(ops, "to_channel_last", _cl)
    monkeypatch.setattr(ops, "to_channel_first", _cf)
    monkeypatch.setattr(ops, "_require_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "_VOLUME_DTYPE", torch.float32)
    monkeypatch.setattr(ops, "_F32_CONV", "bf16x3")
    for k in CALLS:
        CALLS[k] = 0
    assert ops.use_split()


@pytest.mark.parametrize("cin,cout,shape", [(15, 15, (8, 8, 16)), (1, 1, (8, 8, 12))])
def test_v2v_net_keeps_term_pairs_between_layers(split_mode, cin, cout, shape):
    net = v2v_net.V2VNet(cin, cout)
    sd = synthetic.trained_like_state_dict(net, seed=3)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = torch.rand(1, cin, *shape, generator=torch.Generator().manual_seed(1))
    want = nets.v2v_forward(x.double(), sd, dtype=torch.float64)
    with torch.no_grad():
        got = net(x)
    err = float((got.double() - want).abs().max() / want.abs().max())
    # one split in front of the net (reference-contract entry), none between the layers, float32 out of the head
    assert CALLS["split"] == 1 and CALLS["merge"] == 0 and CALLS["simt"] == 0, CALLS
    assert CALLS["pair_out"] >= 20 and CALLS["f32_out"] == 1, CALLS
    assert err < 1e-4, err


def test_pose_resnet_keeps_term_pairs_between_layers(split_mode):
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 15
    net = pose_resnet.get_pose_net(cfg, is_train=False)
    sd = synthetic.trained_like_state_dict(net, seed=4)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = torch.randn(2, 3, 64, 32, generator=torch.Generator().manual_seed(2))
    want = nets.pose_resnet_forward(x.double(), sd, dtype=torch.float64)
    with torch.no_grad():
        got = net(x)
    err = float((got.double() - want).abs().max() / want.abs().max())
    # the image enters through the space-to-depth pair form; every layer incl. the stride-2 ones on the tensor-core path
    assert CALLS["split"] == 0 and CALLS["merge"] == 0 and CALLS["simt"] == 0, CALLS
    assert CALLS["f32_out"] == 1, CALLS
    assert err < 1e-4, err
