"""``torch.autograd.Function`` wrappers over the forward / backward operators of the C ABI: the training path
(batch-statistics BatchNorm, gradients) of the voxel networks, on channel-last float32 activations.

``torch.autograd`` only records the graph; every forward and backward step below is one of the hand-written kernels
(``ops`` / ``grad_ops``).  The reference trains by running autograd through its ATen / cuDNN modules
(``lib/models/multi_person_posenet_ssv.py:222-501``); these functions are what replaces that for
``lib/models/v2v_net.py`` (all blocks), ``ProjectLayer`` and ``SoftArgmaxLayer``.
"""
from __future__ import annotations

import threading

import torch
from torch.autograd import Function

from . import _lib, grad_ops, ops

_TLS = threading.local()      # per thread: nn.DataParallel replicas run their forwards on threads


class grouped_batches:
    """``with grouped_batches(counts, device):`` -- every training-mode BatchNorm inside normalises the leading
    (cube) dimension in GROUPS of ``counts[g]`` consecutive items, each with its own batch statistics, and updates the
    running statistics as ``len(counts)`` successive calls would.  This is how ONE batched pose-net pass reproduces
    the reference's one-call-per-proposal-slot training forward (``lib/models/multi_person_posenet.py:88-99``)."""

    def __init__(self, counts, device):
        self.groups = grad_ops.BnGroups(counts, device) if len(counts) > 1 else None

    def __enter__(self):
        self.prev = getattr(_TLS, "groups", None)
        _TLS.groups = self.groups
        return self.groups

    def __exit__(self, *exc):
        _TLS.groups = self.prev
        return False


def _active_groups(x):
    g = getattr(_TLS, "groups", None)
    if g is not None and int(x.shape[0]) != g.n_items:
        raise _lib.Sp3dError("grouped BatchNorm: the batch has %d items, the groups cover %d" % (int(x.shape[0]), g.n_items))
    return g


class ToChannelLast(Function):
    """``[N, C, *spatial]`` -> channel-last ``[N, *spatial, pitch]`` (a permutation: the backward is its inverse)."""

    @staticmethod
    def forward(ctx, x):
        ctx.channels = int(x.shape[1])
        return ops.to_channel_last(x.float())

    @staticmethod
    def backward(ctx, g):
        return ops.to_channel_first(g.contiguous(), ctx.channels)


class ToChannelFirst(Function):
    @staticmethod
    def forward(ctx, x, channels):
        ctx.pitch = int(x.shape[-1])
        return ops.to_channel_first(x, channels)

    @staticmethod
    def backward(ctx, g):
        return ops.to_channel_last(g.contiguous(), c_pitch=ctx.pitch), None


class Conv(Function):
    """Raw convolution / transposed convolution (+ bias) of an ``nn.Conv*`` / ``nn.ConvTranspose*`` parameter pair."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, transposed, pc=None):
        if pc is None:
            pc = ops.PackedConv(weight, bias, None, stride, padding, transposed=transposed, relu=0)
        ctx.pc, ctx.has_bias = pc, bias is not None
        ctx.save_for_backward(x)
        # float32 FMA kernel, or -- ops.set_float32_conv("bf16x3" / "bf16x6") -- the tcgen05 kernel on split operands
        # where an instantiation exists for the shape (PackedConv.tc_available)
        return pc(x)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = gy.contiguous()
        gx = (grad_ops.conv_dgrad(ctx.pc, gy, out_pitch=int(x.shape[-1]), in_dims=tuple(x.shape[1:4]))
              if ctx.needs_input_grad[0] else None)
        gw = gb = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):     # (frozen layers: dgrad only)
            gw, gb = grad_ops.conv_wgrad(ctx.pc, x, gy, with_bias=ctx.has_bias)
        return gx, gw, gb, None, None, None, None


class BatchNormAct(Function):
    """Training-mode BatchNorm (batch statistics) with an optional ReLU right after it.  Returns ``(y, mean, var)``;
    the statistics are for the caller's running-average update and carry no gradient."""

    @staticmethod
    def forward(ctx, x, gamma, beta, channels, eps, relu, groups=None, running=None):
        mean, var = grad_ops.bn_stats(x, channels, groups=groups, running=running)     # [C], or [n_groups, C]
        scale = gamma * torch.rsqrt(var + eps)                 # [C] vectors: a handful of scalars, not a hot path
        y = grad_ops.bn_apply(x, channels, scale.contiguous(), (beta - mean * scale).contiguous(), relu=1 if relu else 0,
                              groups=groups)
        ctx.channels, ctx.eps, ctx.relu, ctx.groups = channels, eps, bool(relu), groups
        ctx.save_for_backward(x, mean, var, gamma, y if relu else None)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, _gm, _gv):
        x, mean, var, gamma, y = ctx.saved_tensors
        gx, gg, gb = grad_ops.bn_bwd(x, ctx.channels, gy.contiguous(), mean, var, gamma, ctx.eps, y=y, groups=ctx.groups)
        return gx, gg, gb, None, None, None, None, None


class AddAct(Function):
    """``a + b`` or ``relu(a + b)`` on channel-last tensors (the residual joins of the V2V blocks)."""

    @staticmethod
    def forward(ctx, a, b, channels, relu):
        one = torch.ones(channels, device=a.device, dtype=torch.float32)
        y = grad_ops.bn_apply(a, channels, one, torch.zeros_like(one), relu=1 if relu else 0, residual=b)
        ctx.relu = bool(relu)
        ctx.save_for_backward(y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        g = grad_ops.relu_bwd(gy, y) if ctx.relu else gy
        return g, g, None, None


class MaxPool(Function):
    @staticmethod
    def forward(ctx, x, channels, k, s, p):
        ctx.args = (channels, list(k), list(s), list(p))
        ctx.save_for_backward(x)
        return ops.maxpool(x, channels, k, s, p)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        channels, k, s, p = ctx.args
        return grad_ops.maxpool_bwd(x, channels, k, s, p, gy.contiguous()), None, None, None, None


def conv(x, module, transposed=False):
    """``module``: ``nn.Conv{2,3}d`` / ``nn.ConvTranspose{2,3}d`` -> raw convolution of channel-last ``x``.
    The packed weight (and its adjoint for the input gradient) is kept on the module until the parameter changes
    (optimizer step, ``load_state_dict``, ``.to()``)."""
    w, b = module.weight, module.bias
    key = (w.data_ptr(), w._version, None if b is None else (b.data_ptr(), b._version), transposed)
    hit = module.__dict__.get("_sp3d_train_pack")
    if hit is None or hit[0] != key:
        pc = ops.PackedConv(w, b, None, int(module.stride[0]), int(module.padding[0]), transposed=transposed, relu=0)
        hit = module.__dict__["_sp3d_train_pack"] = (key, pc)
    return Conv.apply(x, w, b, int(module.stride[0]), int(module.padding[0]), transposed, hit[1])


def batch_norm(x, bn, relu=False):
    """``nn.BatchNorm{2,3}d`` in training mode on channel-last ``x`` (+ ReLU), with the module's running-statistics
    update (momentum, unbiased running variance, ``num_batches_tracked``) done as ``F.batch_norm`` does it."""
    channels = int(bn.num_features)
    groups = _active_groups(x)
    track = bn.track_running_stats and bn.running_mean is not None
    # momentum updates of the float32 running statistics ride on the statistics kernel (one launch instead of ~8 tiny
    # tensor ops per layer); the cumulative-average form (momentum None) keeps the tensor expression
    fused = (track and bn.momentum is not None and bn.running_mean.dtype == torch.float32
             and bn.running_var.dtype == torch.float32 and bn.running_mean.is_contiguous() and bn.running_var.is_contiguous())
    running = (bn.running_mean, bn.running_var, float(bn.momentum)) if fused else None
    y, mean, var = BatchNormAct.apply(x, bn.weight, bn.bias, channels, float(bn.eps), relu, groups, running)
    if track:
        with torch.no_grad():
            bn.num_batches_tracked += groups.n_groups if groups is not None else 1
            if not fused:
                n = x.numel() // int(x.shape[-1])
                counts = groups.counts if groups is not None else [1]
                per_item = n // sum(counts)
                nbt = int(bn.num_batches_tracked) - len(counts)
                for g, c in enumerate(counts):
                    nbt += 1
                    m = 1.0 / float(nbt) if bn.momentum is None else float(bn.momentum)
                    ng = c * per_item
                    mg, vg = (mean[g], var[g]) if groups is not None else (mean, var)
                    bn.running_mean.mul_(1.0 - m).add_(mg.to(bn.running_mean.dtype), alpha=m)
                    bn.running_var.mul_(1.0 - m).add_((vg * (ng / max(ng - 1, 1))).to(bn.running_var.dtype), alpha=m)
    return y


def add(a, b, channels, relu=False):
    return AddAct.apply(a, b, channels, relu)


def max_pool(x, channels, k, s, p):
    return MaxPool.apply(x, channels, k, s, p)


class Unproject(Function):
    """Per-cube un-projection of heat-maps into channel-last cubes, differentiable with respect to the heat-maps
    (``sp3d_unproject_fwd`` / ``sp3d_unproject_bwd``).  ``hms``: one contiguous float32 ``[B,C,h,w]`` per view."""

    @staticmethod
    def forward(ctx, cams, centers, cube_sample, spec, *hms):
        grid_size, cube_size, img_size, hm_cfg_wh, channels, pitch = spec
        if any(h.dtype != torch.float32 or not h.is_contiguous() for h in hms):
            raise _lib.Sp3dError("autograd.Unproject expects contiguous float32 heat-maps")
        X, Y, Z = [int(s) for s in cube_size]
        n = int(centers.shape[0])
        cubes = torch.empty(n, X, Y, Z, pitch, device=hms[0].device, dtype=torch.float32)
        ops.unproject(list(hms), hms[0].stride(), cams, centers, grid_size, (X, Y, Z), img_size, tuple(hms[0].shape[2:]),
                      channels, cubes, (X * Y * Z * pitch, 1, pitch), out_c_pad=pitch, check_flag=False,
                      cube_sample=cube_sample, heatmap_cfg_wh=hm_cfg_wh)
        ctx.spec = spec
        ctx.save_for_backward(cams, centers, cube_sample, *hms)
        return cubes

    @staticmethod
    def backward(ctx, g):
        cams, centers, cube_sample, *hms = ctx.saved_tensors
        grid_size, cube_size, img_size, hm_cfg_wh, channels, pitch = ctx.spec
        X, Y, Z = [int(s) for s in cube_size]
        grads = [torch.zeros_like(h) for h in hms]
        grad_ops.unproject_bwd(hms, hms[0].stride(), cams, centers, grid_size, (X, Y, Z), img_size,
                               tuple(hms[0].shape[2:]), channels, g.contiguous(), (X * Y * Z * pitch, 1, pitch), grads,
                               check_flag=False, cube_sample=cube_sample, heatmap_cfg_wh=hm_cfg_wh)
        return (None, None, None, None) + tuple(grads)


class SoftArgmax(Function):
    """Soft-argmax of channel-last float32 volumes ``[n,X,Y,Z,pitch]`` -> ``[n,C,3]`` (``sp3d_softargmax3d_fwd`` /
    ``sp3d_softargmax3d_bwd``)."""

    @staticmethod
    def forward(ctx, y, centers, spec):
        channels, cube_size, grid_size, beta = spec
        n, pitch = int(y.shape[0]), int(y.shape[-1])
        X, Y, Z = [int(s) for s in cube_size]
        y = y.contiguous()
        out = ops.softargmax(y, (X * Y * Z * pitch, 1, pitch), n, channels, (X, Y, Z), centers, grid_size, beta)
        ctx.spec = spec
        ctx.save_for_backward(y, centers, out)
        return out

    @staticmethod
    def backward(ctx, g):
        y, centers, out = ctx.saved_tensors
        channels, cube_size, grid_size, beta = ctx.spec
        n, pitch = int(y.shape[0]), int(y.shape[-1])
        X, Y, Z = [int(s) for s in cube_size]
        gx = grad_ops.softargmax_bwd(y, (X * Y * Z * pitch, 1, pitch), n, channels, (X, Y, Z), centers, grid_size, beta,
                                     out, g.contiguous())
        return gx, None, None


class RenderGaussians(Function):
    """Joint pixels ``[V,B,P,J,2]`` (+ people per sample) -> clipped Gaussian heat-maps ``[V,B,J,h,w]``
    (``sp3d_gauss_render_fwd`` / ``sp3d_gauss_render_bwd``)."""

    @staticmethod
    def forward(ctx, kps, n_people, hw, inv_scale, sigma):
        ctx.args = (tuple(hw), float(inv_scale), float(sigma))
        ctx.save_for_backward(kps, n_people)
        return grad_ops.gauss_render(kps, n_people, hw, inv_scale, sigma)

    @staticmethod
    def backward(ctx, g):
        kps, n_people = ctx.saved_tensors
        hw, inv_scale, sigma = ctx.args
        return grad_ops.gauss_render_bwd(kps, n_people, hw, g, inv_scale, sigma), None, None, None, None
