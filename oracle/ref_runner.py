"""Runs the UNMODIFIED reference (staged by ``oracle/make_ref.sh`` into ``oracle/_ref/``) on the host CPU.

TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY -- see ``oracle/__init__.py``.  ``bench.py --impl reference`` and the
``cpu_baseline`` leg time ``MultiPersonPoseNetSSV.forward(..., inference=True)``
(``lib/models/multi_person_posenet_ssv.py:105-153,197-220``) through the reference's own module API (``kind:
"reference"``); the GPU tests use it as a second checker beside the oracle port.  Two import-time-only dependencies
that the image lacks are shimmed (``easydict`` for ``lib/core/config.py:15``, ``vedo`` for
``lib/models/cuboid_proposal_net_soft.py:14``); no reference file is modified.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


class _EasyDict(dict):
    """Attribute-style dict with the subset of ``easydict.EasyDict`` behaviour ``lib/core/config.py`` relies on."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) else x for x in v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setitem__ = __setattr__


def install():
    """Put the staged reference ``lib/`` first on ``sys.path`` (with the two import shims)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `sh oracle/make_ref.sh` where /root/reference exists")
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    if "vedo" not in sys.modules:
        m = types.ModuleType("vedo")
        m.Volume = m.show = None
        sys.modules["vedo"] = m
    lib = os.path.join(REF_ROOT, "lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)


def build_model(image_size, heatmap_size, space_size, space_center, initial_cube_size, grid_size, cube_size,
                max_people, threshold, beta=100.0, root_idx=2, num_joints=15, state_dict=None):
    """The reference's ``multi_person_posenet_ssv`` model (ResNet-50 backbone, root-heat-map RootNet, PoseNet) with the
    given geometry on the CPU, in ``eval()`` mode; ``state_dict`` is loaded with ``strict=True``."""
    install()
    import torch
    from core.config import config as cfg
    import models  # noqa: F401  (the reference's package: lib/models/__init__.py)
    from models import multi_person_posenet_ssv
    cfg.NETWORK.IMAGE_SIZE = np.array(image_size)
    cfg.NETWORK.HEATMAP_SIZE = np.array(heatmap_size)
    cfg.NETWORK.NUM_JOINTS = int(num_joints)
    cfg.NETWORK.ROOTNET_ROOTHM = True
    cfg.NETWORK.ROOTNET_TRAIN_SYNTH = False
    cfg.NETWORK.USE_GT = False
    cfg.NETWORK.BETA = float(beta)
    cfg.DATASET.ROOTIDX = int(root_idx)
    cfg.DATASET.ROOTIDX_PSEUDO = int(root_idx)
    cfg.WITH_ATTN = False
    cfg.POSE_RESNET.NUM_LAYERS = 50
    cfg.MULTI_PERSON.SPACE_SIZE = [float(v) for v in space_size]
    cfg.MULTI_PERSON.SPACE_CENTER = [float(v) for v in space_center]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [int(v) for v in initial_cube_size]
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = int(max_people)
    cfg.MULTI_PERSON.THRESHOLD = float(threshold)
    cfg.PICT_STRUCT.GRID_SIZE = [float(v) for v in grid_size]
    cfg.PICT_STRUCT.CUBE_SIZE = [int(v) for v in cube_size]
    cfg.BACKBONE_MODEL = "pose_resnet"
    cfg.MODEL = "multi_person_posenet_ssv"
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    for p in model.parameters():
        assert p.device.type == "cpu"
    return model, torch


def inference(model, images, meta):
    """``model(views1=images, meta1=meta, inference=True)`` under ``no_grad`` -> ``(pred, heatmaps, grid_centers)``."""
    import torch
    with torch.no_grad():
        return model(views1=images, meta1=meta, inference=True)
