"""V2VNet -- 3-D encoder/decoder over voxel cubes, on the sm_100a convolution kernels.

Module/attribute names, constructor signatures, parameter shapes and state-dict keys are
those of the reference's ``lib/models/v2v_net.py``, so checkpoints load with
``strict=True``.  The arithmetic does not go through ATen/cuDNN: every block lowers to
``sp3d_conv_fwd`` launches on channel-last activations with evaluation-mode BatchNorm, bias,
ReLU and the residual add fused into the convolution epilogue, and ``sp3d_maxpool_fwd``.

``forward(x)`` keeps the reference contract (``[N,C,X,Y,Z]`` in and out); ``forward_cl`` is the
layout-native entry used by the proposal / regression nets (channel-last ``[N,X,Y,Z,pitch]``).

Modules in ``.train()`` mode take the training path: float32 activations, batch-statistics BatchNorm with the
running-average update, and gradients through ``selfpose3d_b200.autograd`` (raw convolution -> ``sp3d_bn_stats`` /
``sp3d_bn_apply`` -> joins; backward through ``sp3d_bn_bwd``, the forward kernel on the adjoint weight,
``sp3d_conv_wgrad``, ``sp3d_maxpool_bwd``).  Modules in ``.eval()`` mode take the fused inference path, which
records no gradient.
"""
from __future__ import annotations

import threading

import torch
import torch.nn as nn

from .. import autograd as ag
from .. import ops


def _to_volume_cl(x):
    """Channel-first float tensor -> channel-last activations in the configured volume dtype."""
    if ops.volume_dtype() == torch.bfloat16:
        return ops.to_channel_last(x.float(), c_pitch=ops.round_up(x.shape[1], 16), dtype=torch.bfloat16)
    if ops.use_split():      # float32-faithful tensor-core mode: two bf16 term planes
        return ops.split_act(ops.to_channel_last(x.float()), int(x.shape[1]))
    return ops.to_channel_last(x.float())


def _to_channel_first(y, channels):
    """Channel-last activations (float32 / bf16 tensor or ``SplitAct``) -> float32-or-bf16 ``[N,C,*spatial]``."""
    if isinstance(y, ops.SplitAct):
        y = ops.merge_act(y, channels)
    return ops.to_channel_first(y, channels)


def _forward_cf(module, x, out_channels):
    """The reference contract ``[N,C,X,Y,Z] -> [N,C',X,Y,Z]`` around a module's channel-last ``forward_cl``."""
    if module.training:
        return ag.ToChannelFirst.apply(module.forward_cl(ag.ToChannelLast.apply(x)), out_channels)
    return _to_channel_first(module.forward_cl(_to_volume_cl(x)), out_channels)


def _no_train(module):
    if module.training:
        raise NotImplementedError(
            "selfpose3d_b200: the training-mode forward (batch-statistics BatchNorm + backward kernels) of %s is not "
            "implemented yet (V2VNet and its blocks are); call .eval() for the inference path" % type(module).__name__)


class _PackedCache:
    """Re-packs a module's parameters for the kernels when they change (load_state_dict, .to()).

    One slot per device: ``nn.DataParallel`` replicas are shallow copies that share this object while their parameters
    live on different GPUs and their forwards run in concurrent threads (``tools/train_3d.py:139-140`` wraps the model
    that way), so the packed weights are keyed by the device of the parameters and guarded by a lock."""

    def __init__(self):
        self.slots = {}
        self.lock = threading.Lock()

    # copy.deepcopy(model) / pickling: a copy starts with an empty cache (the packed tensors are derived data keyed by
    # the ORIGINAL parameters' storage; the lock is not copyable)
    def __deepcopy__(self, memo):
        return _PackedCache()

    def __reduce__(self):
        return (_PackedCache, ())

    def get(self, module, build):
        tensors = list(module.parameters()) + list(module.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        dev = tensors[0].device if tensors else None
        with self.lock:
            hit = self.slots.get(dev)
            if hit is not None and hit[0] == key:
                return hit[1]
        value = build()
        with self.lock:
            self.slots[dev] = (key, value)
        return value


class Basic3DBlock(nn.Module):
    """conv k (pad (k-1)//2) + BN + ReLU -- reference v2v_net.py:10-20."""

    def __init__(self, in_planes, out_planes, kernel_size):
        super().__init__()
        self.block = nn.Sequential(
            nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, stride=1, padding=(kernel_size - 1) // 2),
            nn.BatchNorm3d(out_planes),
            nn.ReLU(True))
        self._cache = _PackedCache()

    def _packed(self):
        c, b = self.block[0], self.block[1]
        return self._cache.get(self, lambda: ops.PackedConv(c.weight, c.bias, b, 1, c.padding[0], relu=1))

    def forward_cl(self, x):
        if self.training:
            return ag.batch_norm(ag.conv(x, self.block[0]), self.block[1], relu=True)
        return self._packed()(x)

    def forward(self, x):
        return _forward_cf(self, x, self.block[0].out_channels)


class Res3DBlock(nn.Module):
    """relu(BN(conv3(relu(BN(conv3 x)))) + skip(x)) -- reference v2v_net.py:23-45."""

    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.res_branch = nn.Sequential(
            nn.Conv3d(in_planes, out_planes, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm3d(out_planes),
            nn.ReLU(True),
            nn.Conv3d(out_planes, out_planes, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm3d(out_planes))
        if in_planes == out_planes:
            self.skip_con = nn.Sequential()
        else:
            self.skip_con = nn.Sequential(
                nn.Conv3d(in_planes, out_planes, kernel_size=1, stride=1, padding=0),
                nn.BatchNorm3d(out_planes))
        self.out_planes = out_planes
        self._cache = _PackedCache()

    def _packed(self):
        def build():
            rb = self.res_branch
            a = ops.PackedConv(rb[0].weight, rb[0].bias, rb[1], 1, 1, relu=1)
            b = ops.PackedConv(rb[3].weight, rb[3].bias, rb[4], 1, 1, relu=1)  # ReLU after the residual add
            s = None
            if len(self.skip_con) > 0:
                s = ops.PackedConv(self.skip_con[0].weight, self.skip_con[0].bias, self.skip_con[1], 1, 0, relu=0)
            return a, b, s
        return self._cache.get(self, build)

    def forward_cl(self, x):
        if self.training:
            rb = self.res_branch
            r = ag.batch_norm(ag.conv(x, rb[0]), rb[1], relu=True)
            r = ag.batch_norm(ag.conv(r, rb[3]), rb[4])
            skip = x if len(self.skip_con) == 0 else ag.batch_norm(ag.conv(x, self.skip_con[0]), self.skip_con[1])
            return ag.add(r, skip, self.out_planes, relu=True)
        a, b, s = self._packed()
        skip = x if s is None else s(x)
        return b(a(x), residual=skip)

    def forward(self, x):
        return _forward_cf(self, x, self.out_planes)


class Pool3DBlock(nn.Module):
    """max-pool k = s = pool_size -- reference v2v_net.py:48-54."""

    def __init__(self, pool_size):
        super().__init__()
        self.pool_size = pool_size

    def forward_cl(self, x, channels):
        k = [self.pool_size] * 3
        if self.training:
            return ag.max_pool(x, channels, k, k, [0, 0, 0])
        return ops.maxpool(x, channels, k, k, [0, 0, 0])

    def forward(self, x):
        c = int(x.shape[1])
        if self.training:
            return ag.ToChannelFirst.apply(self.forward_cl(ag.ToChannelLast.apply(x), c), c)
        return _to_channel_first(self.forward_cl(_to_volume_cl(x), c), c)


class Upsample3DBlock(nn.Module):
    """ConvTranspose3d k2 s2 + BN + ReLU -- reference v2v_net.py:57-69."""

    def __init__(self, in_planes, out_planes, kernel_size, stride):
        super().__init__()
        assert kernel_size == 2
        assert stride == 2
        self.block = nn.Sequential(
            nn.ConvTranspose3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=0,
                               output_padding=0),
            nn.BatchNorm3d(out_planes),
            nn.ReLU(True))
        self.out_planes = out_planes
        self._cache = _PackedCache()

    def _packed(self):
        c, b = self.block[0], self.block[1]
        # relu=2: ReLU before the (optional) skip tensor is added, as `up(x) + skip` in EncoderDecorder
        return self._cache.get(self, lambda: ops.PackedConv(c.weight, c.bias, b, 2, 0, transposed=True, relu=2))

    def forward_cl(self, x, skip=None):
        if self.training:
            y = ag.batch_norm(ag.conv(x, self.block[0], transposed=True), self.block[1], relu=True)
            return y if skip is None else ag.add(y, skip, self.out_planes)
        return self._packed()(x, residual=skip)

    def forward(self, x):
        return _forward_cf(self, x, self.out_planes)


class EncoderDecorder(nn.Module):
    """Two-level encoder/decoder with skip branches -- reference v2v_net.py:72-110."""

    def __init__(self):
        super().__init__()
        self.encoder_pool1 = Pool3DBlock(2)
        self.encoder_res1 = Res3DBlock(32, 64)
        self.encoder_pool2 = Pool3DBlock(2)
        self.encoder_res2 = Res3DBlock(64, 128)
        self.mid_res = Res3DBlock(128, 128)
        self.decoder_res2 = Res3DBlock(128, 128)
        self.decoder_upsample2 = Upsample3DBlock(128, 64, 2, 2)
        self.decoder_res1 = Res3DBlock(64, 64)
        self.decoder_upsample1 = Upsample3DBlock(64, 32, 2, 2)
        self.skip_res1 = Res3DBlock(32, 32)
        self.skip_res2 = Res3DBlock(64, 64)

    def forward_cl(self, x):
        skip_x1 = self.skip_res1.forward_cl(x)
        x = self.encoder_res1.forward_cl(self.encoder_pool1.forward_cl(x, 32))
        skip_x2 = self.skip_res2.forward_cl(x)
        x = self.encoder_res2.forward_cl(self.encoder_pool2.forward_cl(x, 64))
        x = self.decoder_res2.forward_cl(self.mid_res.forward_cl(x))
        x = self.decoder_upsample2.forward_cl(x, skip=skip_x2)
        x = self.decoder_res1.forward_cl(x)
        return self.decoder_upsample1.forward_cl(x, skip=skip_x1)

    def forward(self, x):
        return _forward_cf(self, x, 32)


class V2VNet(nn.Module):
    """reference v2v_net.py:113-144."""

    def __init__(self, input_channels, output_channels):
        super().__init__()
        self.front_layers = nn.Sequential(
            Basic3DBlock(input_channels, 16, 7),
            Res3DBlock(16, 32))
        self.encoder_decoder = EncoderDecorder()
        self.output_layer = nn.Conv3d(32, output_channels, kernel_size=1, stride=1, padding=0)
        self.input_channels = input_channels
        self.output_channels = output_channels
        self._cache = _PackedCache()
        self._initialize_weights()

    def _out_packed(self):
        o = self.output_layer
        return self._cache.get(o, lambda: ops.PackedConv(o.weight, o.bias, None, 1, 0, relu=0))

    def forward_cl(self, x, out_pitch=None):
        """``x``: channel-last ``[N,X,Y,Z,pitch]`` float32, spatial extents divisible by 4.
        Returns channel-last ``[N,X,Y,Z,out_pitch]`` (default pitch: channels rounded up to 4)."""
        if any(int(s) % 4 for s in x.shape[1:4]):
            raise ValueError("V2VNet needs spatial extents divisible by 4, got %s" % (tuple(x.shape[1:4]),))
        if self.training:
            if x.dtype != torch.float32:
                raise ValueError("the training path runs on float32 activations (ops.set_volume_dtype(torch.float32))")
            if out_pitch not in (None, ops.round_up(self.output_channels, 4)):
                raise ValueError("the training path writes the default channel pitch")
        x = self.front_layers[0].forward_cl(x)
        x = self.front_layers[1].forward_cl(x)
        x = self.encoder_decoder.forward_cl(x)
        if self.training:
            return ag.conv(x, self.output_layer)
        # the score / heat-map volume leaves the net in float32 in either mode (NMS and soft-argmax read it)
        return self._out_packed()(x, out_pitch=out_pitch, out_dtype=torch.float32)

    def forward_softargmax_cl(self, x, head):
        """bf16 volume mode: like ``forward_cl`` but the 1x1x1 output layer feeds ``head`` (``ops.SoftargmaxHead``)
        on chip -- the heat-map volume is never written; returns ``[N, output_channels, 3]``."""
        _no_train(self)
        if any(int(s) % 4 for s in x.shape[1:4]):
            raise ValueError("V2VNet needs spatial extents divisible by 4, got %s" % (tuple(x.shape[1:4]),))
        x = self.front_layers[0].forward_cl(x)
        x = self.front_layers[1].forward_cl(x)
        x = self.encoder_decoder.forward_cl(x)
        return self._out_packed()(x, head=head)

    def forward(self, x):
        return _forward_cf(self, x, self.output_channels)

    def _initialize_weights(self):
        # reference v2v_net.py:135-144: N(0, 0.001) weights, zero bias for conv and transposed conv
        for m in self.modules():
            if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d)):
                nn.init.normal_(m.weight, 0, 0.001)
                nn.init.constant_(m.bias, 0)
