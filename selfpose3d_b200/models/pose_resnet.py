"""PoseResNet -- 2-D heat-map backbone (ResNet trunk + 3 transposed convs + 1x1 head).

Interface mirror of the reference's ``lib/models/pose_resnet.py``: ``BasicBlock`` (:25-55),
``Bottleneck`` (:58-93), ``PoseResNet`` (:96-262), ``resnet_spec`` (:265-271), ``get_pose_net``
(:274-284), ``PoseResAttnNet`` (:287-299), ``get_pose_attn_net`` (:323-333); identical module
tree and state-dict keys (338 tensors for ResNet-50).  The forward pass lowers to
``sp3d_conv_fwd`` / ``sp3d_maxpool_fwd`` launches on channel-last activations with
evaluation-mode BatchNorm, ReLU and the residual add fused into the convolution epilogue;
a transposed 4x4 stride-2 convolution is four stride-1 2x2 sub-convolutions, one per output phase.

In ``.train()`` mode the net takes the training path of ``selfpose3d_b200.autograd`` (float32, raw convolution ->
batch-statistics BatchNorm with the running-average update -> ReLU / residual joins; backward through
``sp3d_conv_wgrad``, the forward kernel on the adjoint operator, ``sp3d_bn_bwd``, ``sp3d_maxpool_bwd``).
"""
from __future__ import annotations

import logging
import os

import torch
import torch.nn as nn

from .. import autograd as ag
from .. import ops
from .v2v_net import _PackedCache

BN_MOMENTUM = 0.1
logger = logging.getLogger(__name__)


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


def _pc(conv, bn, relu):
    return ops.PackedConv(conv.weight, conv.bias, bn, conv.stride[0], conv.padding[0], relu=relu)


def _s2d(conv, bn, relu):
    """Space-to-depth form of a stride-2 3x3 / 7x7 convolution for the tcgen05 path (None if not applicable)."""
    if conv.stride[0] != 2 or conv.bias is not None or (conv.kernel_size[0], conv.padding[0]) not in ((3, 1), (7, 3)):
        return None
    if not (conv.in_channels * 4 < 64 or conv.in_channels % 16 == 0) or conv.out_channels % 64:
        return None
    return ops.S2DConv(conv.weight, bn, conv.padding[0], relu)


def _strided_conv_cl(pc, s2d, x):
    """``x``: channel-last ``[N,1,H,W,C]``.  bf16 activations with even extents take the space-to-depth tensor-core
    form of a stride-2 convolution; everything else the generic kernel."""
    N, _, H, W, C = [int(v) for v in x.shape]
    if s2d is not None and H % 2 == 0 and W % 2 == 0 and C == s2d.cin:
        if isinstance(x, ops.SplitAct):      # float32-faithful tensor-core mode
            return s2d.call_split(x)
        if x.dtype == torch.bfloat16:
            return s2d(x, (H * W * C, 1, W * C, C), N, H, W)
    return pc(x)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.downsample = downsample
        self.stride = stride
        self._cache = _PackedCache()

    def _packed(self):
        def build():
            d = _pc(self.downsample[0], self.downsample[1], 0) if self.downsample is not None else None
            return _pc(self.conv1, self.bn1, 1), _pc(self.conv2, self.bn2, 1), d
        return self._cache.get(self, build)

    def forward_cl(self, x):
        if self.training:
            r = ag.batch_norm(ag.conv(x, self.conv1), self.bn1, relu=True)
            r = ag.batch_norm(ag.conv(r, self.conv2), self.bn2)
            skip = x if self.downsample is None else ag.batch_norm(ag.conv(x, self.downsample[0]), self.downsample[1])
            return ag.add(r, skip, self.conv2.out_channels, relu=True)
        a, b, d = self._packed()
        return b(a(x), residual=x if d is None else d(x))


# Training forwards over several views run as ONE pass with one BatchNorm statistic group per view
# (PoseResNet.forward_views); SP3D_SLOT_BATCH=0 keeps one pass per view (A/B checks).
_VIEW_BATCH = os.environ.get("SP3D_SLOT_BATCH", "1") != "0"


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self._cache = _PackedCache()

    def _packed(self):
        def build():
            d = _pc(self.downsample[0], self.downsample[1], 0) if self.downsample is not None else None
            return (_pc(self.conv1, self.bn1, 1), _pc(self.conv2, self.bn2, 1), _pc(self.conv3, self.bn3, 1), d,
                    _s2d(self.conv2, self.bn2, 1))
        return self._cache.get(self, build)

    def forward_cl(self, x):
        if self.training:
            r = ag.batch_norm(ag.conv(x, self.conv1), self.bn1, relu=True)
            r = ag.batch_norm(ag.conv(r, self.conv2), self.bn2, relu=True)
            r = ag.batch_norm(ag.conv(r, self.conv3), self.bn3)
            skip = x if self.downsample is None else ag.batch_norm(ag.conv(x, self.downsample[0]), self.downsample[1])
            return ag.add(r, skip, self.conv3.out_channels, relu=True)
        a, b, c, d, b2 = self._packed()
        return c(_strided_conv_cl(b, b2, a(x)), residual=x if d is None else d(x))


class PoseResNet(nn.Module):
    def __init__(self, block, layers, cfg, **kwargs):
        self.inplanes = 64
        self.deconv_with_bias = cfg.POSE_RESNET.DECONV_WITH_BIAS
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.deconv_layers = self._make_deconv_layer(
            cfg.POSE_RESNET.NUM_DECONV_LAYERS, cfg.POSE_RESNET.NUM_DECONV_FILTERS, cfg.POSE_RESNET.NUM_DECONV_KERNELS)
        self.final_layer = nn.Conv2d(
            in_channels=cfg.POSE_RESNET.NUM_DECONV_FILTERS[-1], out_channels=cfg.NETWORK.NUM_JOINTS,
            kernel_size=cfg.POSE_RESNET.FINAL_CONV_KERNEL, stride=1,
            padding=1 if cfg.POSE_RESNET.FINAL_CONV_KERNEL == 3 else 0)
        self.num_joints = cfg.NETWORK.NUM_JOINTS
        self._cache = _PackedCache()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion, momentum=BN_MOMENTUM))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    @staticmethod
    def _get_deconv_cfg(deconv_kernel, index):
        return {4: (4, 1, 0), 3: (3, 1, 1), 2: (2, 0, 0)}[deconv_kernel]

    def _make_deconv_layer(self, num_layers, num_filters, num_kernels):
        assert num_layers == len(num_filters) == len(num_kernels)
        layers = []
        for i in range(num_layers):
            kernel, padding, output_padding = self._get_deconv_cfg(num_kernels[i], i)
            layers += [
                nn.ConvTranspose2d(self.inplanes, num_filters[i], kernel_size=kernel, stride=2, padding=padding,
                                   output_padding=output_padding, bias=self.deconv_with_bias),
                nn.BatchNorm2d(num_filters[i], momentum=BN_MOMENTUM),
                nn.ReLU(inplace=True)]
            self.inplanes = num_filters[i]
        return nn.Sequential(*layers)

    def _packed(self):
        def build():
            stem = _pc(self.conv1, self.bn1, 1)
            deconvs = []
            for i in range(0, len(self.deconv_layers), 3):
                ct, bn = self.deconv_layers[i], self.deconv_layers[i + 1]
                if ct.kernel_size[0] % ct.stride[0] != 0 or ct.output_padding[0] != 0:
                    raise NotImplementedError("transposed conv with kernel %d stride %d output_padding %d"
                                              % (ct.kernel_size[0], ct.stride[0], ct.output_padding[0]))
                deconvs.append(ops.PackedConv(ct.weight, ct.bias, bn, ct.stride[0], ct.padding[0], transposed=True,
                                              relu=1))
            head = _pc(self.final_layer, None, 0)
            return stem, deconvs, head, _s2d(self.conv1, self.bn1, 1)
        mods = nn.ModuleList([self.conv1, self.bn1, self.deconv_layers, self.final_layer])
        return self._cache.get(mods, build)

    def forward_cl(self, x, out_pitch=None, image=None):
        """``x``: channel-last ``[N,1,H,W,4]`` float32 image batch, or ``image``: the ``[N,3,H,W]`` float32 tensor
        itself -> ``(heat-maps [N,1,h,w,pitch], features)``."""
        if self.training:
            return self._forward_train_cl(x, out_pitch, image)
        stem, deconvs, head, stem_s2d = self._packed()
        # bf16 mode: the 7x7 stem reads the float32 image and writes bf16; from there on activations are bf16
        # (tcgen05 convolutions where the shape is covered) and the heat-maps leave the net in float32
        bf16 = ops.volume_dtype() == torch.bfloat16
        split = ops.use_split()   # float32 values as bf16 term pairs from layer to layer (3-pair tensor-core mode)
        if image is not None and (bf16 or split) and stem_s2d is not None and image.shape[2] % 2 == 0 and image.shape[3] % 2 == 0:
            # straight from the NCHW float32 image: 2x2 space-to-depth (bf16 / term pairs) + 4x4 tensor-core convolution
            n, _, h, w = [int(v) for v in image.shape]
            x = stem_s2d.call_split(image, image.stride(), n, h, w) if split else stem_s2d(image, image.stride(), n, h, w)
        else:
            if x is None:
                x = ops.to_channel_last(image.unsqueeze(2), c_pitch=4)
            x = stem(x, out_dtype=torch.bfloat16 if bf16 else None)
            if split:
                x = ops.split_act(x, 64)
        x = ops.maxpool(x, 64, [1, 3, 3], [1, 2, 2], [0, 1, 1])
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk.forward_cl(x)
        for d in deconvs:
            x = d(x)
        return head(x, out_pitch=out_pitch, out_dtype=torch.float32), x

    def _forward_train_cl(self, x, out_pitch, image):
        """Training path (float32): every convolution raw, BatchNorm on batch statistics, gradients recorded."""
        if out_pitch not in (None, ops.round_up(self.num_joints, 4)):
            raise ValueError("the training path writes the default channel pitch")
        if x is None:
            x = ag.ToChannelLast.apply(image.unsqueeze(2))
        x = ag.batch_norm(ag.conv(x, self.conv1), self.bn1, relu=True)
        x = ag.max_pool(x, 64, [1, 3, 3], [1, 2, 2], [0, 1, 1])
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk.forward_cl(x)
        for i in range(0, len(self.deconv_layers), 3):
            ct, bn = self.deconv_layers[i], self.deconv_layers[i + 1]
            if ct.output_padding[0] != 0:
                raise NotImplementedError("transposed convolution with output_padding in the training path")
            x = ag.batch_norm(ag.conv(x, ct, transposed=True), bn, relu=True)
        return ag.conv(x, self.final_layer), x

    def forward(self, x, attn=False):
        """``[N,3,H,W]`` -> ``[N,J,H/4,W/4]`` (reference :191-207).  The result is a zero-copy
        channel-last view (``stride(1) == 1``), which the un-projection kernel reads directly."""
        y, feat = self.forward_cl(None, out_pitch=ops.round_up(self.num_joints, 4), image=x.float().contiguous())
        out = y[:, 0].permute(0, 3, 1, 2)[:, :self.num_joints]
        if attn:
            if isinstance(feat, ops.SplitAct):
                feat = ops.merge_act(feat, int(feat.shape[-1]))
            return out, feat[:, 0].permute(0, 3, 1, 2)[:, :feat.shape[-1]]
        return out

    def forward_views(self, views):
        """``[backbone(view) for view in views]`` -- the reference's one call per view
        (``lib/models/multi_person_posenet_ssv.py:227-277``, ``multi_person_posenet.py:38-41``) -- in ONE pass in
        training mode: the views are concatenated and every view is a BatchNorm statistic group of its own
        (``autograd.grouped_batches``), so batch statistics, gradients and the running-average updates are those of the
        per-view calls, at a fifth of the launches."""
        views = list(views)
        if not self.training or len(views) < 2 or not _VIEW_BATCH:
            return [self(view) for view in views]
        counts = [int(v.shape[0]) for v in views]
        x = torch.cat([v.float() for v in views], dim=0)
        with ag.grouped_batches(counts, x.device):
            y = self(x)
        return list(torch.split(y, counts, dim=0))

    def init_weights(self, pretrained="", mapping=None):
        """Reference :209-262: load an ImageNet/COCO checkpoint if the file exists (remapping the
        final layer to the Panoptic joint order), else N(0, 0.001) weights and unit BatchNorm."""
        path = os.path.join(os.path.dirname(__file__), "../..", pretrained)
        if pretrained and os.path.isfile(path):
            state = torch.load(path)
            own = self.state_dict()
            for k in list(state.keys()):
                if "final_layer" in k:
                    if state[k].shape[0] != own[k].shape[0]:
                        state[k] = state[k][mapping]
                    else:
                        state[k] = torch.zeros_like(own[k])
            missing, unexpected = self.load_state_dict(state, strict=False)
            logger.info("=> loaded %s (missing %s, unexpected %s)", path, missing, unexpected)
            if missing and unexpected:
                self._init_head()
            return
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, std=0.001)
                if isinstance(m, nn.ConvTranspose2d) and self.deconv_with_bias:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _init_head(self):
        for m in self.deconv_layers.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.normal_(m.weight, std=0.001)
                if self.deconv_with_bias:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.final_layer.weight, std=0.001)
        nn.init.constant_(self.final_layer.bias, 0)


resnet_spec = {
    18: (BasicBlock, [2, 2, 2, 2]),
    34: (BasicBlock, [3, 4, 6, 3]),
    50: (Bottleneck, [3, 4, 6, 3]),
    101: (Bottleneck, [3, 4, 23, 3]),
    152: (Bottleneck, [3, 8, 36, 3]),
}


def get_pose_net(cfg, is_train, **kwargs):
    block_class, layers = resnet_spec[cfg.POSE_RESNET.NUM_LAYERS]
    model = PoseResNet(block_class, layers, cfg, **kwargs)
    if is_train:
        model.init_weights(cfg.NETWORK.PRETRAINED, mapping=cfg.COCO_TO_PANOPTIC_MAPPING)
    return model


class PoseResAttnNet(nn.Module):
    """Attention branch of the SSL model: a PoseResNet followed by a sigmoid (reference :287-299).
    Only used by the training losses and ``visualize_attn``; the sigmoid is a torch elementwise op."""

    def __init__(self, block, layers, cfg, **kwargs):
        super().__init__()
        self.backbone = PoseResNet(block, layers, cfg, **kwargs)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        return self.sigmoid(self.backbone(x))

    def forward_views(self, views):
        return [self.sigmoid(y) for y in self.backbone.forward_views(views)]

    def init_weights(self, pretrained="", mapping=None):
        self.backbone.init_weights(pretrained, mapping)


def get_pose_attn_net(cfg, is_train, **kwargs):
    block_class, layers = resnet_spec[cfg.ATTN_NUM_LAYERS]
    model = PoseResAttnNet(block_class, layers, cfg, **kwargs)
    if is_train:
        model.init_weights(cfg.NETWORK.PRETRAINED, mapping=cfg.COCO_TO_PANOPTIC_MAPPING)
    return model
