"""Torch-tensor front end of the backward operators of the C ABI (``include/sp3d.h``, "Backward operators").

Same rules as ``ops.py``: CUDA tensors in, POD argument structs, launches on torch's current stream, no CPU path.
All tensors are float32; activations are channel-last ``[N, D, H, W, pitch]``.

What is here is the operator level of the training path: gradients of the un-projection,
the soft-argmax, max pooling, the convolution family (weight gradient kernel; the input gradient is the forward
kernel on the adjoint weight) and training-mode BatchNorm.  The module-level ``autograd`` wiring of whole nets is the
next step and is not part of this file.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .ops import _require_cuda, _set3, _stream


def _f32(*tensors):
    for t in tensors:
        if t is not None and (t.dtype != torch.float32 or not t.is_cuda):
            raise _lib.Sp3dError("backward operators take float32 CUDA tensors")


# --------------------------------------------------------------------------------------------- un-projection
def unproject_bwd(heatmaps, hm_strides, cams, centers, grid_size, cube_size, image_size, heatmap_hw, channels,
                  grad_cubes, grad_strides, grad_heatmaps, check_flag=False, cubes_per_sample=1, cube_sample=None,
                  heatmap_cfg_wh=None):
    """Accumulate ``dL/dheatmaps`` into ``grad_heatmaps`` (list[V], addressed like ``heatmaps`` through
    ``hm_strides``; zero them first) given ``grad_cubes`` addressed through ``grad_strides = (cube, channel, voxel)``.
    The remaining arguments are those of the forward ``ops.unproject`` call."""
    _f32(grad_cubes, *heatmaps, *grad_heatmaps)
    b = _lib.UnprojectBwdArgs()
    ops.unproject_args(heatmaps, hm_strides, cams, centers, grid_size, cube_size, image_size, heatmap_hw, channels,
                       None, grad_strides, 0, check_flag, cubes_per_sample, cube_sample, None, None, False,
                       heatmap_cfg_wh, False, a=b.fwd)
    b.grad_cubes = grad_cubes.data_ptr()
    for v, g in enumerate(grad_heatmaps):
        b.grad_heatmaps[v] = g.data_ptr()
    n_vox = b.fwd.X * b.fwd.Y * b.fwd.Z
    _lib.call("sp3d_unproject_bwd", b, _stream(), kind="unproject_bwd",
              work=b.fwd.n_cubes * b.fwd.C * n_vox * 4 + 2 * b.fwd.V * b.fwd.B * b.fwd.C * b.fwd.h * b.fwd.w * 4)


# --------------------------------------------------------------------------------------------- soft-argmax
def softargmax_bwd(x, strides, n_cubes, channels, cube_size, centers, grid_size, beta, out, grad_out, check_flag=False,
                   lin=None):
    """``dL/dx`` of ``ops.softargmax`` (same arguments; ``out`` = its result, ``grad_out [n_cubes, C, 3]``).
    Returns a tensor shaped and strided like ``x`` (padding channels zero)."""
    _f32(x, out, grad_out)
    _require_cuda(centers)
    gx = torch.zeros_like(x)
    b = _lib.SoftargmaxBwdArgs()
    a = b.fwd
    a.x, a.x_dtype = x.data_ptr(), _lib.F32
    a.stride_cube, a.stride_c, a.stride_vox = [int(s) for s in strides]
    a.n_cubes, a.C = int(n_cubes), int(channels)
    a.X, a.Y, a.Z = [int(s) for s in cube_size]
    a.centers, a.center_stride = centers.data_ptr(), int(centers.stride(0))
    a.check_flag = int(bool(check_flag))
    if lin is None:
        lin = ops.linspace_axes(grid_size, cube_size, x.device)
    a.lin_x, a.lin_y, a.lin_z = lin[0].data_ptr(), lin[1].data_ptr(), lin[2].data_ptr()
    a.beta = float(beta)
    out, grad_out = out.contiguous(), grad_out.contiguous()
    a.out = out.data_ptr()
    b.grad_out = grad_out.data_ptr()
    b.grad_x = gx.data_ptr()
    ws = torch.empty(max(a.n_cubes * 64 * a.C * 2, 1), device=x.device, dtype=torch.float32)   # streaming form's partials
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    _lib.call("sp3d_softargmax3d_bwd", b, _stream(), launches=2, kind="softargmax_bwd",
              work=4 * a.n_cubes * a.C * a.X * a.Y * a.Z * 4)
    return gx


# --------------------------------------------------------------------------------------------- max pool
def maxpool_bwd(x, channels, k, s, p, grad_out):
    """``dL/dx`` of ``ops.maxpool(x, channels, k, s, p)`` for float32 channel-last ``x``."""
    _f32(x, grad_out)
    N, D, H, W, pitch = [int(v) for v in x.shape]
    dims = (D, H, W)
    o = [(dims[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3)]
    if tuple(grad_out.shape) != (N, o[0], o[1], o[2], pitch) or not grad_out.is_contiguous() or not x.is_contiguous():
        raise _lib.Sp3dError("maxpool_bwd: grad_out must be the contiguous pooled shape")
    gi = torch.empty_like(x)
    b = _lib.MaxpoolBwdArgs()
    a = b.fwd
    a.in_ = x.data_ptr()
    a.N, a.D, a.H, a.W, a.C, a.c_pitch = N, D, H, W, int(channels), pitch
    a.OD, a.OH, a.OW = o
    _set3(a.k, k)
    _set3(a.s, s)
    _set3(a.p, p)
    a.dtype = _lib.F32
    b.grad_out, b.grad_in = grad_out.data_ptr(), gi.data_ptr()
    _lib.call("sp3d_maxpool_bwd", b, _stream(), launches=2, kind="maxpool_bwd", work=3 * x.numel() * 4)
    return gi


# --------------------------------------------------------------------------------------------- convolutions
def _conv_geometry(a, x, out_shape, cin, cout, cout_pitch_w, out_grid, ksize, stride, tap_off0, tap_step, ostride, ooffset):
    a.N, a.D, a.H, a.W = [int(s) for s in x.shape[:4]]
    a.cin, a.cin_pitch = int(cin), int(x.shape[4])
    a.OD, a.OH, a.OW = [int(s) for s in out_grid]
    a.TD, a.TH, a.TW = [int(s) for s in out_shape[1:4]]
    a.cout, a.cout_pitch, a.cout_pitch_w = int(cout), int(out_shape[4]), int(cout_pitch_w)
    _set3(a.ksize, ksize)
    _set3(a.stride, stride)
    _set3(a.tap_off0, tap_off0)
    _set3(a.tap_step, tap_step)
    _set3(a.ostride, ostride)
    _set3(a.ooffset, ooffset)
    a.in_dtype = a.out_dtype = _lib.F32


def conv_wgrad(pc, x, grad_out, with_bias=True):
    """Weight / bias gradient of the convolution ``pc`` (an ``ops.PackedConv``; its folded scale / shift / ReLU are
    NOT part of this: ``grad_out`` is the gradient of the raw convolution result) for the forward input ``x``.
    Returns ``(grad_weight, grad_bias)`` in the layout of the reference parameter (``nn.Conv*``: ``[Cout,Cin,k..]``,
    ``nn.ConvTranspose*``: ``[Cin,Cout,k..]``)."""
    _f32(x, grad_out)
    if not x.is_contiguous() or not grad_out.is_contiguous():
        raise _lib.Sp3dError("conv_wgrad expects contiguous channel-last tensors")
    N, D, H, W, pitch = [int(v) for v in x.shape]
    if pitch < pc.cin_p:
        raise _lib.Sp3dError("activation pitch %d smaller than packed cin %d" % (pitch, pc.cin_p))
    o = pc.out_shape((D, H, W))
    if tuple(grad_out.shape[:4]) != (N, o[0], o[1], o[2]) or int(grad_out.shape[4]) < pc.cout:
        raise _lib.Sp3dError("conv_wgrad: grad_out does not match the convolution's output shape")
    gb = torch.zeros(pc.cout, device=x.device, dtype=torch.float32) if with_bias else None
    use_tc = _wgrad_tc_ok(pc, x, grad_out)
    subs = []
    launches = [(None, [-p for p in pc.padding], pc.k, o, [1, 1, 1], [1, 1, 1], [0, 0, 0])] if not pc.transposed else [
        (i, off0, ks, [(o[d] - phase[d] + pc.stride[d] - 1) // pc.stride[d] for d in range(3)], [-1, -1, -1], pc.stride, phase)
        for i, (phase, off0, ks) in enumerate(pc.phases)]
    for (_, off0, ks, grid, step, ostride, ooff) in launches:
        taps = int(ks[0] * ks[1] * ks[2])
        gw = torch.zeros(taps, pc.cin_p, pc.cout_pw, device=x.device, dtype=torch.float32)
        if taps == 0 or min(grid) <= 0:        # a phase without taps / without output positions
            subs.append(gw[:, :pc.cin, :pc.cout].reshape(int(ks[0]), int(ks[1]), int(ks[2]), pc.cin, pc.cout)
                        .permute(4, 3, 0, 1, 2))
            continue
        if use_tc:
            # transposed phases walk their taps backwards (tap_step -1 from off0): the same window read forwards starts
            # at off0 - (k - 1), with the tap order reversed in the result
            fwd_off = [int(off0[d]) if step[d] == 1 else int(off0[d]) - (int(ks[d]) - 1) for d in range(3)]
            _conv_wgrad_tc(pc, x, grad_out, gb, gw, ks, fwd_off, ostride, ooff)
            if step[0] == -1:
                gw = gw.reshape(int(ks[0]), int(ks[1]), int(ks[2]), pc.cin_p, pc.cout_pw).flip(0, 1, 2).reshape(
                    taps, pc.cin_p, pc.cout_pw)
        else:
            b = _lib.ConvWgradArgs()
            _conv_geometry(b.fwd, x, grad_out.shape, pc.cin_p, pc.cout, pc.cout_pw, grid, ks,
                           pc.stride if not pc.transposed else [1, 1, 1], off0, step, ostride, ooff)
            b.fwd.in_ = x.data_ptr()
            b.grad_out, b.grad_weight = grad_out.data_ptr(), gw.data_ptr()
            b.grad_bias = gb.data_ptr() if gb is not None else None
            flops = 2.0 * N * grid[0] * grid[1] * grid[2] * pc.cout * pc.cin * taps
            _lib.call("sp3d_conv_wgrad", b, _stream(), kind="conv_wgrad", work=flops)
        subs.append(gw[:, :pc.cin, :pc.cout].reshape(int(ks[0]), int(ks[1]), int(ks[2]), pc.cin, pc.cout)
                    .permute(4, 3, 0, 1, 2))                       # [Cout, Cin, a, b, c]
    if not pc.transposed:
        w5 = subs[0]
    else:
        w5 = torch.zeros(pc.cout, pc.cin, *pc.k, device=x.device, dtype=torch.float32)
        s_, p_ = pc.stride, pc.padding
        for sub, (phase, _, _) in zip(subs, pc.phases):
            t0 = [(phase[i] + p_[i]) % s_[i] for i in range(3)]
            w5[:, :, t0[0]::s_[0], t0[1]::s_[1], t0[2]::s_[2]] = sub
        w5 = w5.permute(1, 0, 2, 3, 4)                             # nn.ConvTranspose layout [Cin, Cout, k..]
    shape = list(w5.shape[:2]) + list(w5.shape[2 + (3 - pc.nd):])
    return w5.reshape(shape).contiguous(), gb


_WGRAD_TC = __import__("os").environ.get("SP3D_WGRAD_TC", "1") != "0"


def _wgrad_tc_ok(pc, x, grad_out):
    """The tcgen05 weight gradient (``sp3d_conv_wgrad_tc``) takes the convolutions whose taps step by one input position
    per output position: stride-1 convolutions (3-D and 2-D) and stride-s transposed convolutions (phase by phase), with a
    z extent (row length) of at most 128, ``round_up(cin, 16)`` in {16, 32, 64} or more than 64 input channels, at most 128
    output channels or a multiple of 128.  Not in the float32 FMA mode."""
    if not _WGRAD_TC or ops.float32_conv() == "simt":
        return False
    if not pc.transposed and pc.stride != [1, 1, 1]:
        return False
    if max(pc.k) > 7 or int(x.shape[3]) > 128:
        return False
    cp, co16 = ops.round_up(pc.cin, 16), ops.round_up(pc.cout, 16)
    if cp not in (16, 32, 64) and cp < 65:
        return False
    if co16 > 128 and co16 % 128:
        return False
    if pc.transposed:
        # every phase must cover the whole input grid (output extent = stride * input extent)
        D, H, W = [int(v) for v in x.shape[1:4]]
        o = pc.out_shape((D, H, W))
        for phase, _, ks in pc.phases:
            grid = [(o[d] - phase[d] + pc.stride[d] - 1) // pc.stride[d] for d in range(3)]
            if int(ks[0] * ks[1] * ks[2]) and grid != [D, H, W]:
                return False
    return True


def _conv_wgrad_tc(pc, x, grad_out, gb, gw, ks, tap_off, g_stride, g_off):
    """One ``sp3d_conv_wgrad_tc`` launch: ``gw [taps, cin_p, cout_pw] +=`` the weight gradient of the tap window ``ks``
    starting at input offset ``tap_off``, with the gradient read at ``p * g_stride + g_off`` (bias gradient added into
    ``gb``)."""
    N, X, Y, Z, pitch = [int(v) for v in x.shape]
    a = _lib.ConvWgradTcArgs()
    a.x, a.grad_out = x.data_ptr(), grad_out.data_ptr()
    a.N, a.X, a.Y, a.Z = N, X, Y, Z
    a.cin, a.x_pitch, a.cout, a.g_pitch = pc.cin, pitch, pc.cout, int(grad_out.shape[4])
    _set3(a.ksize, ks)
    _set3(a.tap_off, tap_off)
    a.GX, a.GY, a.GZ = [int(v) for v in grad_out.shape[1:4]]
    _set3(a.g_stride, g_stride)
    _set3(a.g_off, g_off)
    a.grad_weight, a.gw_cin, a.gw_pitch = gw.data_ptr(), pc.cin_p, pc.cout_pw
    a.grad_bias = gb.data_ptr() if gb is not None else None
    nbytes = int(_lib.load().sp3d_conv_wgrad_tc_workspace(a))
    if nbytes < 0:
        raise _lib.Sp3dError("sp3d_conv_wgrad_tc: shape not supported")
    ws = torch.empty(nbytes + 16, device=x.device, dtype=torch.uint8)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    taps = int(ks[0] * ks[1] * ks[2])
    flops = 2.0 * N * X * Y * Z * pc.cout * pc.cin * taps
    _lib.call("sp3d_conv_wgrad_tc", a, _stream(), launches=3, kind="conv_wgrad_tc", work=flops,
              detail="wgrad tc k%dx%dx%d %d->%d @%dx%dx%dx%d" % (int(ks[0]), int(ks[1]), int(ks[2]), pc.cin, pc.cout, N, X, Y, Z))


def conv_dgrad(pc, grad_out, out_pitch=None, in_dims=None):
    """Input gradient of the convolution ``pc`` (raw convolution, no scale / shift / ReLU) -- the forward kernel
    on the adjoint operator:
      * stride-1 convolution           -> convolution with the flipped, channel-transposed kernel, padding k-1-p;
      * strided convolution            -> transposed convolution with the same kernel (``in_dims``: the forward input's
                                          spatial extent ``(D, H, W)``, which the output extent alone does not determine);
      * transposed convolution         -> strided convolution with the same kernel."""
    _f32(grad_out)
    adj = pc.__dict__.get("_adjoint")
    if adj is None:
        def as_param(w5):        # [A, B, kd, kh, kw] -> drop the leading unit dims of a 2-D kernel
            return w5.reshape(list(w5.shape[:2]) + list(w5.shape[2 + (3 - pc.nd):])).contiguous()
        if not pc.transposed and pc.stride == [1, 1, 1]:
            wa = pc._subs[0].permute(1, 0, 2, 3, 4).flip(2, 3, 4)      # an nn.Conv weight with Cout' = Cin
            pads = [pc.k[i] - 1 - pc.padding[i] for i in range(3)][3 - pc.nd:]
            if len(set(pads)) != 1:
                raise _lib.Sp3dError("conv_dgrad: anisotropic padding is not covered")
            adj = ops.PackedConv(as_param(wa), None, None, 1, pads[0], relu=0)
        elif not pc.transposed:
            # adjoint of a strided convolution: nn.ConvTranspose weight layout [in = Cout, out = Cin, k..] = the kernel as is
            adj = ops.PackedConv(as_param(pc._subs[0]), None, None, pc.stride[-1], pc.padding[-1], transposed=True, relu=0)
        else:
            full = torch.zeros(pc.cout, pc.cin, *pc.k, device=grad_out.device, dtype=torch.float32)
            s_, p_ = pc.stride, pc.padding
            for sub, (phase, _, _) in zip(pc._subs, pc.phases):
                if sub.numel():
                    t0 = [(phase[i] + p_[i]) % s_[i] for i in range(3)]
                    full[:, :, t0[0]::s_[0], t0[1]::s_[1], t0[2]::s_[2]] = sub
            # [Cin_t, Cout_t, k..]: an nn.Conv weight (Cout' = Cin_t) applied with the transposed convolution's stride / padding
            adj = ops.PackedConv(as_param(full.permute(1, 0, 2, 3, 4)), None, None, pc.stride[-1], pc.padding[-1], relu=0)
        pc.__dict__["_adjoint"] = adj
    if adj.transposed:
        if in_dims is None:
            raise _lib.Sp3dError("conv_dgrad of a strided convolution needs in_dims (the forward input's extent)")
        return adj(grad_out, algo=_lib.CONV_SIMT_F32, out_pitch=out_pitch, out_dims=[int(v) for v in in_dims])
    return adj(grad_out, out_pitch=out_pitch)      # float32 FMA kernel, or split operands on tcgen05 where compiled


# --------------------------------------------------------------------------------------------- BatchNorm (training)
def _flat(x):
    if not x.is_contiguous():
        raise _lib.Sp3dError("BatchNorm operators expect contiguous channel-last tensors")
    pitch = int(x.shape[-1])
    return x.numel() // pitch, pitch


class BnGroups:
    """Statistic groups of a batched launch: item ``i`` (a cube of the leading dimension) belongs to group
    ``item_group[i]``; every group is normalised with its own batch statistics (``include/sp3d.h``, grouped
    BatchNorm).  ``counts``: items per group (host list, in group order)."""

    def __init__(self, counts, device):
        self.counts = [int(c) for c in counts]
        if not self.counts or min(self.counts) < 1:
            raise ValueError("every BatchNorm group needs at least one item")
        self.n_groups, self.n_items = len(self.counts), sum(self.counts)
        ids = [g for g, c in enumerate(self.counts) for _ in range(c)]
        self.item_group = torch.tensor(ids, dtype=torch.int32).to(device)
        self.group_items = torch.tensor(self.counts, dtype=torch.int32).to(device)

    def fill(self, a, x, with_counts=True):
        if int(x.shape[0]) != self.n_items:
            raise ValueError("grouped BatchNorm: %d items expected, tensor has %d" % (self.n_items, int(x.shape[0])))
        a.n_items, a.n_groups, a.item_group = self.n_items, self.n_groups, self.item_group.data_ptr()
        if with_counts:
            a.group_items = self.group_items.data_ptr()


def bn_stats(x, channels, groups=None, running=None):
    """Per-channel batch mean and biased variance of channel-last ``x`` -> ``(mean, var)``, each ``[C]`` or, with
    ``groups`` (``BnGroups``), ``[n_groups, C]``.  ``running = (running_mean, running_var, momentum)``: the module's
    float32 running statistics are updated in place by the same launch (once per group, in group order)."""
    _f32(x)
    P, pitch = _flat(x)
    G = groups.n_groups if groups is not None else 1
    mean = torch.empty((G, channels) if groups is not None else (channels,), device=x.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    ws = torch.empty(G * 2 * channels, device=x.device, dtype=torch.float64)
    a = _lib.BnStatsArgs()
    a.x, a.P, a.C, a.pitch = x.data_ptr(), P, int(channels), pitch
    a.mean, a.var = mean.data_ptr(), var.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 8
    if groups is not None:
        groups.fill(a, x)
    if running is not None:
        _f32(running[0], running[1])
        a.running_mean, a.running_var, a.momentum = running[0].data_ptr(), running[1].data_ptr(), float(running[2])
    _lib.call("sp3d_bn_stats", a, _stream(), launches=2, kind="bn", work=x.numel() * 4)
    if running is not None:       # the kernel wrote them in place: caches keyed on the buffers' versions must notice
        torch.autograd.graph.increment_version(running[0])
        torch.autograd.graph.increment_version(running[1])
    return mean, var


def bn_apply(x, channels, scale, shift, relu=0, residual=None, groups=None):
    """``act(x * scale[c] + shift[c] (+ residual))`` on channel-last ``x`` (``relu`` as in ``sp3d_conv_args``);
    ``scale`` / ``shift`` are ``[n_groups, C]`` with ``groups``."""
    _f32(x, scale, shift, residual)
    P, pitch = _flat(x)
    y = torch.empty_like(x)
    a = _lib.BnApplyArgs()
    a.x, a.y = x.data_ptr(), y.data_ptr()
    a.residual = residual.data_ptr() if residual is not None else None
    a.P, a.C, a.pitch = P, int(channels), pitch
    a.scale, a.shift, a.relu = scale.data_ptr(), shift.data_ptr(), int(relu)
    if groups is not None:
        groups.fill(a, x, with_counts=False)
    _lib.call("sp3d_bn_apply", a, _stream(), kind="bn", work=2 * x.numel() * 4)
    return y


def bn_bwd(x, channels, grad_y, mean, var, gamma, eps, y=None, groups=None):
    """Backward of training-mode BatchNorm (+ the ReLU right after it when its output ``y`` is given) ->
    ``(grad_x, grad_gamma, grad_beta)``; with ``groups`` the statistics are ``[n_groups, C]`` and the parameter
    gradients are summed over the groups."""
    _f32(x, grad_y, mean, var, gamma, y)
    P, pitch = _flat(x)
    G = groups.n_groups if groups is not None else 1
    gx = torch.empty_like(x)
    gg = torch.empty(channels, device=x.device, dtype=torch.float32)
    gb = torch.empty_like(gg)
    ws = torch.empty(G * 2 * channels, device=x.device, dtype=torch.float64)
    a = _lib.BnBwdArgs()
    grad_y = grad_y.contiguous()
    a.x, a.grad_y = x.data_ptr(), grad_y.data_ptr()
    a.y = y.data_ptr() if y is not None else None
    a.P, a.C, a.pitch = P, int(channels), pitch
    a.mean, a.var = mean.data_ptr(), var.data_ptr()
    a.gamma = gamma.data_ptr() if gamma is not None else None
    a.eps = float(eps)
    a.grad_x, a.grad_gamma, a.grad_beta = gx.data_ptr(), gg.data_ptr(), gb.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 8
    if groups is not None:
        groups.fill(a, x)
    _lib.call("sp3d_bn_bwd", a, _stream(), launches=3, kind="bn", work=4 * x.numel() * 4)
    return gx, gg, gb


def relu_bwd(grad_y, y):
    """``grad_y`` where ``y > 0`` else 0 (backward of the ReLU that produced ``y``)."""
    _f32(grad_y, y)
    grad_y, y = grad_y.contiguous(), y.contiguous()
    gx = torch.empty_like(y)
    a = _lib.ReluBwdArgs()
    a.grad_y, a.y, a.grad_x, a.n = grad_y.data_ptr(), y.data_ptr(), gx.data_ptr(), y.numel()
    _lib.call("sp3d_relu_bwd", a, _stream(), kind="elementwise", work=3 * y.numel() * 4)
    return gx


# --------------------------------------------------------------------------------------------- Gaussian joint rendering
def _gauss_args(a, kps, n_people, hw, inv_scale, sigma):
    V, B, P, J, _ = [int(v) for v in kps.shape]
    a.kps, a.n_people = kps.data_ptr(), n_people.data_ptr()
    a.V, a.B, a.P, a.J, a.h, a.w = V, B, P, J, int(hw[0]), int(hw[1])
    a.inv_scale, a.sigma = float(inv_scale), float(sigma)


def gauss_render(kps, n_people, hw, inv_scale=0.25, sigma=3.0):
    """``kps [V,B,P,J,2]`` joint pixels (network input), ``n_people [B]`` int32 -> heat-maps ``[V,B,J,h,w]``:
    clipped sum of the people's Gaussians (``sp3d_gauss_render_fwd``)."""
    _f32(kps)
    _require_cuda(n_people)
    kps = kps.contiguous()
    V, B, P, J, _ = [int(v) for v in kps.shape]
    out = torch.empty(V, B, J, int(hw[0]), int(hw[1]), device=kps.device, dtype=torch.float32)
    a = _lib.GaussRenderArgs()
    _gauss_args(a, kps, n_people, hw, inv_scale, sigma)
    a.heatmaps = out.data_ptr()
    _lib.call("sp3d_gauss_render_fwd", a, _stream(), kind="render", work=out.numel() * 4)
    return out


def gauss_render_bwd(kps, n_people, hw, grad_heatmaps, inv_scale=0.25, sigma=3.0):
    """Gradient of ``gauss_render`` with respect to ``kps``."""
    _f32(kps, grad_heatmaps)
    kps, grad_heatmaps = kps.contiguous(), grad_heatmaps.contiguous()
    gk = torch.zeros_like(kps)
    b = _lib.GaussRenderBwdArgs()
    _gauss_args(b.fwd, kps, n_people, hw, inv_scale, sigma)
    b.grad_heatmaps, b.grad_kps = grad_heatmaps.data_ptr(), gk.data_ptr()
    _lib.call("sp3d_gauss_render_bwd", b, _stream(), kind="render", work=grad_heatmaps.numel() * 4)
    return gk
