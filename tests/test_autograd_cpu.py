"""CPU check of the training-path WIRING (``selfpose3d_b200/autograd.py`` + the ``.train()`` branches of
``models/v2v_net.py``): every kernel entry point is replaced by a small torch emulation of its documented semantics
(include/sp3d.h), and a whole V2VNet training step (forward, running statistics, gradients of the input and of all
156 parameters) is compared with autograd through the oracle's functional restatement of the reference net.  The
kernels themselves are checked on the GPU (tests/test_gpu_backward.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nets
from selfpose3d_b200 import autograd as ag
from selfpose3d_b200 import grad_ops, ops, synthetic
from selfpose3d_b200.models import v2v_net
from test_conv_lowering_cpu import emulate_conv_launch


def _cl(x, c_pitch=None, dtype=None):
    x = x.float()
    N, C = x.shape[:2]
    pitch = ops.round_up(C, 4) if c_pitch is None else c_pitch
    out = torch.zeros((N,) + tuple(x.shape[2:]) + (pitch,))
    out[..., :C] = x.permute(0, *range(2, x.dim()), 1)
    return out


def _cf(x, channels, dtype=None):
    return x[..., :channels].permute(0, x.dim() - 1, *range(1, x.dim() - 1)).contiguous()


def _maxpool(x, channels, k, s, p):
    y = F.max_pool3d(x.permute(0, 4, 1, 2, 3), k, s, p)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def _maxpool_bwd(x, channels, k, s, p, gy):
    with torch.enable_grad():      # (called from inside Function.backward, where grad mode is off)
        xc = x.permute(0, 4, 1, 2, 3).detach().clone().requires_grad_(True)
        F.max_pool3d(xc, k, s, p).backward(gy.permute(0, 4, 1, 2, 3))
    return xc.grad.permute(0, 2, 3, 4, 1).contiguous()


def _group_slices(groups):
    out, o = [], 0
    for c in groups.counts:
        out.append(slice(o, o + c))
        o += c
    return out


def _bn_stats(x, channels, groups=None, running=None):
    if groups is not None:      # grouped statistics: every group of items on its own (running statistics: in group order)
        st = [_bn_stats(x[sl], channels, running=running) for sl in _group_slices(groups)]
        return torch.stack([m for m, _ in st]), torch.stack([v for _, v in st])
    v = x.reshape(-1, x.shape[-1])[:, :channels].double()
    mean, var = v.mean(0).float(), v.var(0, unbiased=False).float()
    if running is not None:     # (running_mean, running_var, momentum): updated in place, the variance unbiased
        rm, rv, m = running
        n = v.shape[0]
        rm.mul_(1.0 - m).add_(mean, alpha=m)
        rv.mul_(1.0 - m).add_(var * (n / max(n - 1, 1)), alpha=m)
    return mean, var


def _bn_apply(x, channels, scale, shift, relu=0, residual=None, groups=None):
    if groups is not None:
        return torch.cat([_bn_apply(x[sl], channels, scale[g], shift[g], relu, None if residual is None else residual[sl])
                          for g, sl in enumerate(_group_slices(groups))])
    y = torch.zeros_like(x)
    r = x[..., :channels] * scale + shift
    if relu == 2:
        r = r.clamp_min(0)
    if residual is not None:
        r = r + residual[..., :channels]
    if relu == 1:
        r = r.clamp_min(0)
    y[..., :channels] = r
    return y


def _bn_bwd(x, channels, grad_y, mean, var, gamma, eps, y=None, groups=None):
    if groups is not None:
        parts = [_bn_bwd(x[sl], channels, grad_y[sl], mean[g], var[g], gamma, eps, None if y is None else y[sl])
                 for g, sl in enumerate(_group_slices(groups))]
        return (torch.cat([p[0] for p in parts]), sum(p[1] for p in parts), sum(p[2] for p in parts))
    P = x.numel() // x.shape[-1]
    xs, dz = x[..., :channels].double(), grad_y[..., :channels].double()
    if y is not None:
        dz = dz * (y[..., :channels] > 0)
    istd = 1.0 / torch.sqrt(var.double() + eps)
    xh = (xs - mean.double()) * istd
    red = tuple(range(x.dim() - 1))
    db, dg = dz.sum(red), (dz * xh).sum(red)
    gx = torch.zeros_like(x)
    gx[..., :channels] = (gamma.double() * istd * (dz - db / P - xh * dg / P)).float()
    return gx, dg.float(), db.float()


def _relu_bwd(grad_y, y):
    return grad_y * (y > 0)


def _conv_wgrad(pc, x, grad_out, with_bias=True):
    # independent route: autograd of torch's own convolution on the reference-shaped parameter
    w5 = pc._subs[0] if not pc.transposed else None
    xc = x[..., :pc.cin].permute(0, 4, 1, 2, 3)
    go = grad_out[..., :pc.cout].permute(0, 4, 1, 2, 3)
    with torch.enable_grad():
        if not pc.transposed:
            w = w5.detach().clone().requires_grad_(True)
            y = F.conv3d(xc, w, None, stride=pc.stride, padding=pc.padding)
        else:
            full = torch.zeros(pc.cout, pc.cin, *pc.k)
            for sub, (phase, _, _) in zip(pc._subs, pc.phases):
                full[:, :, phase[0]::pc.stride[0], phase[1]::pc.stride[1], phase[2]::pc.stride[2]] = sub
            w = full.permute(1, 0, 2, 3, 4).detach().clone().requires_grad_(True)
            y = F.conv_transpose3d(xc, w, None, stride=pc.stride, padding=pc.padding)
        y.backward(go)
    return w.grad, (go.sum((0, 2, 3, 4)) if with_bias else None)


@pytest.fixture
def emulated(monkeypatch):
    monkeypatch.setattr(ops, "conv_launch", emulate_conv_launch)
    monkeypatch.setattr(ops, "to_channel_last", _cl)
    monkeypatch.setattr(ops, "to_channel_first", _cf)
    monkeypatch.setattr(ops, "maxpool", _maxpool)
    monkeypatch.setattr(grad_ops, "_f32", lambda *a: None)
    for name, fn in (("maxpool_bwd", _maxpool_bwd), ("bn_stats", _bn_stats), ("bn_apply", _bn_apply), ("bn_bwd", _bn_bwd),
                     ("relu_bwd", _relu_bwd), ("conv_wgrad", _conv_wgrad)):
        monkeypatch.setattr(grad_ops, name, fn)


@pytest.mark.parametrize("cin,shape", [(3, (2, 3, 8, 8, 4)), (1, (1, 1, 4, 8, 8))])
def test_v2v_net_training_step_wiring(emulated, cin, shape):
    torch.manual_seed(cin)
    net = v2v_net.V2VNet(cin, cin)
    sd0 = synthetic.trained_like_state_dict(net, seed=50 + cin)
    net.load_state_dict(sd0, strict=True)
    net.train()
    x = torch.rand(*shape)
    gy = torch.randn(shape[0], cin, *shape[2:])

    # oracle: autograd through the functional restatement of the reference net in training mode (float64)
    sd = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd0.items()}
    xo = x.double().clone().requires_grad_(True)
    yo = nets.v2v_forward(xo, sd, dtype=torch.float64, training=True)
    (yo * gy.double()).sum().backward()

    xin = x.clone().requires_grad_(True)
    y = net(xin)
    assert y.shape == yo.shape
    (y * gy).sum().backward()

    def rel(a, b):
        return float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), 1e-30)

    assert rel(y, yo) < 1e-4, rel(y, yo)
    assert rel(xin.grad, xo.grad) < 1e-3, rel(xin.grad, xo.grad)
    n_checked = 0
    for name, p in net.named_parameters():
        ref = sd[name].grad
        assert p.grad is not None and p.grad.shape == ref.shape, name
        if float(ref.abs().max()) < 1e-9:          # conv biases in front of a batch normalisation: exactly cancelled
            assert float(p.grad.abs().max()) < 1e-4 * max(1.0, float(gy.abs().max())), name
            continue
        assert rel(p.grad, ref) < 2e-3, (name, rel(p.grad, ref))
        n_checked += 1
    assert n_checked >= 60
    # running statistics: momentum update with the unbiased batch variance, as nn.BatchNorm3d in .train()
    bn = net.front_layers[0].block[1]
    z = F.conv3d(x.double(), sd0["front_layers.0.block.0.weight"].double(), sd0["front_layers.0.block.0.bias"].double(),
                 padding=3)
    m = bn.momentum
    want_mean = (1 - m) * sd0["front_layers.0.block.1.running_mean"].double() + m * z.mean((0, 2, 3, 4))
    want_var = (1 - m) * sd0["front_layers.0.block.1.running_var"].double() + m * z.var((0, 2, 3, 4), unbiased=True)
    assert rel(bn.running_mean, want_mean) < 1e-5 and rel(bn.running_var, want_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(sd0["front_layers.0.block.1.num_batches_tracked"]) + 1

    # eval mode afterwards: the fused inference path, no graph
    net.eval()
    with torch.no_grad():
        assert net(x).shape == y.shape


def _maxpool_any(x, channels, k, s, p):
    y = F.max_pool3d(x.permute(0, 4, 1, 2, 3), tuple(k), tuple(s), tuple(p))
    return y.permute(0, 2, 3, 4, 1).contiguous()


def _maxpool_bwd_any(x, channels, k, s, p, gy):
    with torch.enable_grad():
        xc = x.permute(0, 4, 1, 2, 3).detach().clone().requires_grad_(True)
        F.max_pool3d(xc, tuple(k), tuple(s), tuple(p)).backward(gy.permute(0, 4, 1, 2, 3))
    return xc.grad.permute(0, 2, 3, 4, 1).contiguous()


def _conv_wgrad_any(pc, x, grad_out, with_bias=True):
    """Independent route for 2-D / strided / transposed layers: autograd of torch's convolution on the reference-shaped
    parameter (3-D functional calls with a unit leading extent)."""
    xc = x[..., :pc.cin].permute(0, 4, 1, 2, 3)
    go = grad_out[..., :pc.cout].permute(0, 4, 1, 2, 3)
    with torch.enable_grad():
        if not pc.transposed:
            w = pc._subs[0].detach().clone().requires_grad_(True)
            y = F.conv3d(xc, w, None, stride=pc.stride, padding=pc.padding)
        else:
            full = torch.zeros(pc.cout, pc.cin, *pc.k)
            for sub, (phase, _, _) in zip(pc._subs, pc.phases):
                t0 = [(phase[i] + pc.padding[i]) % pc.stride[i] for i in range(3)]
                full[:, :, t0[0]::pc.stride[0], t0[1]::pc.stride[1], t0[2]::pc.stride[2]] = sub
            w = full.permute(1, 0, 2, 3, 4).detach().clone().requires_grad_(True)
            y = F.conv_transpose3d(xc, w, None, stride=pc.stride, padding=pc.padding)
        y.backward(go)
    g = w.grad
    g = g.reshape(list(g.shape[:2]) + list(g.shape[2 + (3 - pc.nd):]))
    return g, (go.sum((0, 2, 3, 4)) if with_bias else None)


def test_pose_resnet_training_step_wiring(emulated, monkeypatch):
    """PoseResNet-50 in .train() on a 64 x 96 image batch: output, running statistics and the gradients of all 161
    parameters (stride-2 3x3 / 1x1 convolutions and the k4/s2 transposed convolutions included) against autograd
    through the oracle's functional restatement."""
    from selfpose3d_b200.config import default_config
    from selfpose3d_b200.models import pose_resnet
    monkeypatch.setattr(ops, "maxpool", _maxpool_any)
    monkeypatch.setattr(grad_ops, "maxpool_bwd", _maxpool_bwd_any)
    monkeypatch.setattr(grad_ops, "conv_wgrad", _conv_wgrad_any)
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 5
    net = pose_resnet.get_pose_net(cfg, is_train=False)
    sd0 = synthetic.trained_like_state_dict(net, seed=90)
    net.load_state_dict(sd0, strict=True)
    net.train()
    torch.manual_seed(2)
    x = torch.randn(2, 3, 64, 96)
    gy = torch.randn(2, 5, 16, 24)
    def oracle(dtype):
        sd_ = {k: (v.to(dtype).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
               for k, v in sd0.items()}
        yo_ = nets.pose_resnet_forward(x.to(dtype), sd_, dtype=dtype, training=True)
        (yo_ * gy.to(dtype)).sum().backward()
        return yo_.detach(), sd_

    yo, sd = oracle(torch.float64)
    _, sd32 = oracle(torch.float32)     # yardstick: ReLU gates decided in the last bits differ between float32 / float64
    y = net(x)
    assert y.shape == yo.shape
    (y * gy).sum().backward()

    def rel(a, b):
        return float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), 1e-30)

    assert rel(y, yo) < 1e-3, rel(y, yo)
    top = max(float(v.grad.abs().max()) for v in sd.values() if v.is_floating_point() and v.grad is not None)
    n_checked = 0
    for name, p in net.named_parameters():
        ref = sd[name].grad
        assert p.grad is not None and p.grad.shape == ref.shape, name
        if float(ref.abs().max()) < 1e-7 * top:
            continue
        # tiny feature maps (2 x 3) and a batch of 2 make single ReLU gates flip between float32 and float64 on either
        # side (the float32 oracle is itself up to 0.14 of the range away from the float64 one on single entries), so
        # the gradients are compared in the L2 norm, which a wiring error would move by O(1) and a flipped gate by ~1e-2
        l2 = float((p.grad.double() - ref).norm() / ref.norm())
        assert l2 < max(3e-2, 3 * float((sd32[name].grad.double() - ref).norm() / ref.norm())), (name, l2)
        n_checked += 1
    assert n_checked >= 100, n_checked
    bn = net.layer2[0].downsample[1]                   # BatchNorm behind the 1x1 / stride-2 convolution
    assert int(bn.num_batches_tracked) == int(sd0["layer2.0.downsample.1.num_batches_tracked"]) + 1
    assert not torch.equal(bn.running_mean, sd0["layer2.0.downsample.1.running_mean"])


def test_grouped_batches_context_and_group_table():
    """``autograd.grouped_batches`` / ``grad_ops.BnGroups``: the item -> group table, nesting, and the refusal of a batch
    that does not match the groups (a silent fall-back to whole-batch statistics would change the training result)."""
    from selfpose3d_b200 import _lib
    g = grad_ops.BnGroups([2, 1, 3], "cpu")
    assert g.n_groups == 3 and g.n_items == 6
    assert g.item_group.tolist() == [0, 0, 1, 2, 2, 2] and g.group_items.tolist() == [2, 1, 3]
    with pytest.raises(ValueError):
        grad_ops.BnGroups([2, 0], "cpu")
    assert ag._active_groups(torch.zeros(4, 1)) is None
    with ag.grouped_batches([2, 2], "cpu") as outer:
        assert ag._active_groups(torch.zeros(4, 1)) is outer
        with ag.grouped_batches([3], "cpu") as inner:          # a single group = plain batch statistics
            assert inner is None and ag._active_groups(torch.zeros(3, 1)) is None
        assert ag._active_groups(torch.zeros(4, 1)) is outer
        with pytest.raises(_lib.Sp3dError):
            ag._active_groups(torch.zeros(5, 1))
    assert ag._active_groups(torch.zeros(5, 1)) is None
