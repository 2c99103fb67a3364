"""MultiPersonPoseNetSSV -- top-level SSL model (backbone -> root net -> pose net).

Interface mirror of the reference's ``lib/models/multi_person_posenet_ssv.py``: constructor
``(backbone, cfg, attn=None)`` (:28-101), ``do_inference`` (:105-153), the 18-argument
``forward`` (:197-220) and ``get_multi_person_pose_net`` (:504-514); attributes ``backbone``,
``root_net``, ``pose_net``, ``attn`` and every state-dict key as in the reference.
``forward(..., inference=True)`` runs on the sm_100a kernels.  The SSL training branch of
``forward`` (:226-501: three augmented view sets, re-projection, Gaussian rendering, attention
and Hungarian losses) is in ``_ssl_train.py``: the networks take their training paths through the
backward kernels, the loss assembly around them is plain tensor expressions on small tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import pose_resnet
from . import _inference
from . import _ssl_train
from .cuboid_proposal_net_soft import CuboidProposalNetSoft
from .pose_regression_net import PoseRegressionNet


class MultiPersonPoseNetSSV(nn.Module):
    def __init__(self, backbone, cfg, attn=None):
        super().__init__()
        self.num_cand = cfg.MULTI_PERSON.MAX_PEOPLE_NUM
        self.num_joints = cfg.NETWORK.NUM_JOINTS
        self.backbone = backbone
        self.WITH_ATTN = cfg.WITH_ATTN
        if self.WITH_ATTN:
            self.attn = attn
            self.attn_weight = cfg.ATTN_WEIGHT
        self.USE_L1, self.L1_WEIGHT, self.L1_ATTN = cfg.USE_L1, cfg.L1_WEIGHT, cfg.L1_ATTN
        self.L1_EPOCH = cfg.TRAIN.L1_EPOCH
        self.width, self.height = list(cfg.NETWORK.IMAGE_SIZE)[0], list(cfg.NETWORK.IMAGE_SIZE)[1]
        self.use_root_gt = cfg.NETWORK.USE_GT
        self.train_only_2d = cfg.NETWORK.TRAIN_ONLY_2D
        self.root_id = cfg.DATASET.ROOTIDX
        self.dataset_name = cfg.DATASET.TEST_DATASET
        self.train_only_rootnet = cfg.NETWORK.TRAIN_ONLY_ROOTNET
        self.rootnet_train_synth = cfg.NETWORK.ROOTNET_TRAIN_SYNTH
        self.freeze_rootnet = cfg.NETWORK.FREEZE_ROOTNET
        self.single_aug_training_posenet = cfg.NETWORK.SINGLE_AUG_TRAINING_POSENET
        self.init_train_epochs_rootnet = cfg.NETWORK.INIT_TRAIN_EPOCHS_ROOTNET
        self.root_reg_loss = cfg.NETWORK.ROOT_CONSISTENCY_LOSS
        self.weight_root_syn, self.weight_root_reg = cfg.NETWORK.WEIGHT_ROOT_SYN, cfg.NETWORK.WEIGHT_ROOT_REG
        self.eval_rootnet_only = cfg.EVAL_ROOTNET_ONLY
        # heat-map pixel-index grids of the Gaussian rendering (reference :88-101; not part of the state dict)
        hw, hh = int(cfg.NETWORK.HEATMAP_SIZE[0]), int(cfg.NETWORK.HEATMAP_SIZE[1])
        yy, xx = torch.meshgrid(torch.arange(hh, dtype=torch.float32), torch.arange(hw, dtype=torch.float32), indexing="ij")
        self.register_buffer("hm_xx", xx.view(1, 1, hh, hw), persistent=False)
        self.register_buffer("hm_yy", yy.view(1, 1, hh, hw), persistent=False)
        if self.train_only_2d:
            self.use_root_gt = True
        elif not self.train_only_rootnet:
            self.pose_net = PoseRegressionNet(cfg)
        if not self.use_root_gt:
            self.root_net = CuboidProposalNetSoft(cfg)

    def do_inference(self, views=None, meta=None, input_heatmaps=None, visualize_attn=False):
        skip_pose = self.train_only_rootnet or self.train_only_2d
        pred, all_heatmaps, grid_centers, _ = _inference.infer(
            self, views, meta, input_heatmaps, self.use_root_gt, self.eval_rootnet_only, skip_pose)
        if visualize_attn and views is not None:
            attns = torch.stack([self.attn(view) for view in views], 0)
            return pred, all_heatmaps, grid_centers, attns
        return pred, all_heatmaps, grid_centers

    def forward(self, views1=None, meta1=None, targets_2d1=None, weights_2d1=None, targets_3d1=None,
                input_heatmaps1=None, views2=None, meta2=None, targets_2d2=None, weights_2d2=None,
                targets_3d2=None, input_heatmaps2=None, views3=None, meta3=None, targets_2d3=None,
                weights_2d3=None, targets_3d3=None, input_heatmaps3=None, inference=False,
                visualize_attn=False, epoch=0):
        if inference:
            with torch.no_grad():
                return self.do_inference(views1, meta1, input_heatmaps1, visualize_attn)
        return _ssl_train.forward_train(
            self, views1, meta1, targets_2d1, weights_2d1, targets_3d1, input_heatmaps1,
            views2, meta2, targets_2d2, weights_2d2, targets_3d2, input_heatmaps2,
            views3, meta3, targets_2d3, weights_2d3, targets_3d3, input_heatmaps3, epoch)


def get_multi_person_pose_net(cfg, is_train=True):
    backbone = None
    if cfg.BACKBONE_MODEL:
        backbone = getattr(_backbones()[cfg.BACKBONE_MODEL], "get_pose_net")(cfg, is_train=is_train)
    if cfg.WITH_ATTN:
        attn = getattr(_backbones()[cfg.BACKBONE_MODEL], "get_pose_attn_net")(cfg, is_train=is_train)
        return MultiPersonPoseNetSSV(backbone, cfg, attn)
    return MultiPersonPoseNetSSV(backbone, cfg)


def _backbones():
    return {"pose_resnet": pose_resnet}
