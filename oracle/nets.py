"""Oracle: V2VNet and PoseResNet forward passes from a state dict (torch CPU functional).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

The reference's dense arithmetic is PyTorch's ``conv3d`` / ``conv_transpose3d`` /
``batch_norm`` / ``max_pool`` (``lib/models/v2v_net.py:10-144``,
``lib/models/pose_resnet.py:58-207``); this file restates the layer wiring as
plain functions over a ``{key: tensor}`` state dict (keys as in SURVEY.md
App. B / the reference ``state_dict()``), evaluation mode (running statistics).
``dtype=torch.float64`` gives the up-cast truth for tolerance budgeting.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


_BATCH_STATS = False   # True inside v2v_forward(training=True): nn.BatchNorm3d in .train() mode (batch statistics)


def _bn(x, sd, pfx, eps=1e-5):
    if _BATCH_STATS:
        return F.batch_norm(x, None, None, sd[pfx + ".weight"], sd[pfx + ".bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[pfx + ".running_mean"], sd[pfx + ".running_var"],
                        sd[pfx + ".weight"], sd[pfx + ".bias"], False, 0.0, eps)


def _cast(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# ----------------------------------------------------------------------------- V2V
def basic3d(x, sd, pfx):
    """Basic3DBlock (v2v_net.py:10-20): conv k (pad (k-1)//2) + BN + ReLU."""
    w = sd[pfx + ".block.0.weight"]
    x = F.conv3d(x, w, sd[pfx + ".block.0.bias"], padding=(w.shape[2] - 1) // 2)
    return F.relu(_bn(x, sd, pfx + ".block.1"))


def res3d(x, sd, pfx):
    """Res3DBlock (v2v_net.py:23-45)."""
    r = F.conv3d(x, sd[pfx + ".res_branch.0.weight"], sd[pfx + ".res_branch.0.bias"], padding=1)
    r = F.relu(_bn(r, sd, pfx + ".res_branch.1"))
    r = F.conv3d(r, sd[pfx + ".res_branch.3.weight"], sd[pfx + ".res_branch.3.bias"], padding=1)
    r = _bn(r, sd, pfx + ".res_branch.4")
    if pfx + ".skip_con.0.weight" in sd:
        s = F.conv3d(x, sd[pfx + ".skip_con.0.weight"], sd[pfx + ".skip_con.0.bias"])
        s = _bn(s, sd, pfx + ".skip_con.1")
    else:
        s = x
    return F.relu(r + s)


def upsample3d(x, sd, pfx):
    """Upsample3DBlock (v2v_net.py:57-69): ConvTranspose3d k2 s2 + BN + ReLU."""
    x = F.conv_transpose3d(x, sd[pfx + ".block.0.weight"], sd[pfx + ".block.0.bias"], stride=2)
    return F.relu(_bn(x, sd, pfx + ".block.1"))


def v2v_forward(x, sd, pfx="", dtype=torch.float32, training=False):
    """V2VNet.forward (v2v_net.py:126-131) with EncoderDecorder.forward (:91-110).  ``training``: the module in
    ``.train()`` mode (batch-statistics BatchNorm); differentiable with respect to ``x`` and the entries of ``sd``."""
    global _BATCH_STATS
    _BATCH_STATS = bool(training)
    try:
        return _v2v_forward(x, sd, pfx, dtype)
    finally:
        _BATCH_STATS = False


def _v2v_forward(x, sd, pfx, dtype):
    sd = _cast(sd, dtype)
    x = x.to(dtype)
    x = basic3d(x, sd, pfx + "front_layers.0")
    x = res3d(x, sd, pfx + "front_layers.1")
    e = pfx + "encoder_decoder."
    skip1 = res3d(x, sd, e + "skip_res1")
    x = F.max_pool3d(x, 2, 2)
    x = res3d(x, sd, e + "encoder_res1")
    skip2 = res3d(x, sd, e + "skip_res2")
    x = F.max_pool3d(x, 2, 2)
    x = res3d(x, sd, e + "encoder_res2")
    x = res3d(x, sd, e + "mid_res")
    x = res3d(x, sd, e + "decoder_res2")
    x = upsample3d(x, sd, e + "decoder_upsample2") + skip2
    x = res3d(x, sd, e + "decoder_res1")
    x = upsample3d(x, sd, e + "decoder_upsample1") + skip1
    return F.conv3d(x, sd[pfx + "output_layer.weight"], sd[pfx + "output_layer.bias"])


# ----------------------------------------------------------------------------- PoseResNet
def _bottleneck(x, sd, pfx, stride):
    """Bottleneck (pose_resnet.py:58-93); stride sits on the 3x3 conv (:65)."""
    out = F.relu(_bn(F.conv2d(x, sd[pfx + ".conv1.weight"]), sd, pfx + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[pfx + ".conv2.weight"], stride=stride, padding=1), sd, pfx + ".bn2"))
    out = _bn(F.conv2d(out, sd[pfx + ".conv3.weight"]), sd, pfx + ".bn3")
    if pfx + ".downsample.0.weight" in sd:
        x = _bn(F.conv2d(x, sd[pfx + ".downsample.0.weight"], stride=stride), sd, pfx + ".downsample.1")
    return F.relu(out + x)


def pose_resnet_forward(x, sd, pfx="", layers=(3, 4, 6, 3), dtype=torch.float32, training=False):
    """PoseResNet.forward (pose_resnet.py:191-207) for Bottleneck nets, deconv k4 s2 p1 without bias.  ``training``:
    the module in ``.train()`` mode (batch-statistics BatchNorm); differentiable."""
    global _BATCH_STATS
    _BATCH_STATS = bool(training)
    try:
        return _pose_resnet_forward(x, sd, pfx, layers, dtype)
    finally:
        _BATCH_STATS = False


def _pose_resnet_forward(x, sd, pfx, layers, dtype):
    sd = _cast(sd, dtype)
    x = x.to(dtype)
    x = F.relu(_bn(F.conv2d(x, sd[pfx + "conv1.weight"], stride=2, padding=3), sd, pfx + "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, n in enumerate(layers):
        for b in range(n):
            stride = 2 if (b == 0 and li > 0) else 1
            x = _bottleneck(x, sd, "%slayer%d.%d" % (pfx, li + 1, b), stride)
    i = 0
    while pfx + "deconv_layers.%d.weight" % i in sd:
        x = F.conv_transpose2d(x, sd[pfx + "deconv_layers.%d.weight" % i],
                               sd.get(pfx + "deconv_layers.%d.bias" % i), stride=2, padding=1)
        x = F.relu(_bn(x, sd, pfx + "deconv_layers.%d" % (i + 1)))
        i += 3
    fw = sd[pfx + "final_layer.weight"]
    return F.conv2d(x, fw, sd[pfx + "final_layer.bias"], padding=1 if fw.shape[2] == 3 else 0)
