"""MultiPersonPoseNet -- supervised (VoxelPose) top-level model.

Interface mirror of the reference's ``lib/models/multi_person_posenet.py:20-111``: same
constructor, ``forward(views, meta, targets_2d, weights_2d, targets_3d, input_heatmaps)`` and
6-tuple return ``(pred, all_heatmaps, grid_centers, loss_2d, loss_3d, loss_cord)`` (or
``(loss_2d, all_heatmaps)`` with ``TRAIN_ONLY_2D``).  Evaluation mode runs on the sm_100a kernels;
the loss values are the reference's trivial MSE reductions evaluated with torch under ``no_grad``
(SURVEY.md section 2.1 row 11).  Training mode (autograd) raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import pose_resnet
from . import _inference
from .cuboid_proposal_net import CuboidProposalNet
from .pose_regression_net import PoseRegressionNet


def _per_joint_mse(output, target, weight=None):
    """PerJointMSELoss (reference lib/core/loss.py:40-58)."""
    if weight is not None:
        B, J = output.shape[:2]
        return torch.mean((output.reshape(B, J, -1) * weight - target.reshape(B, J, -1) * weight) ** 2)
    return torch.mean((output - target) ** 2)


class MultiPersonPoseNet(nn.Module):
    def __init__(self, backbone, cfg):
        super().__init__()
        self.num_cand = cfg.MULTI_PERSON.MAX_PEOPLE_NUM
        self.num_joints = cfg.NETWORK.NUM_JOINTS
        self.train_only_2d = cfg.NETWORK.TRAIN_ONLY_2D
        self.backbone = backbone
        if not self.train_only_2d:
            self.root_net = CuboidProposalNet(cfg)
            self.pose_net = PoseRegressionNet(cfg)
        self.USE_GT = cfg.NETWORK.USE_GT
        self.root_id = cfg.DATASET.ROOTIDX
        self.dataset_name = cfg.DATASET.TEST_DATASET

    def forward(self, views=None, meta=None, targets_2d=None, weights_2d=None, targets_3d=None,
                input_heatmaps=None):
        if self.training:
            raise NotImplementedError(
                "selfpose3d_b200: supervised training forward/backward is not implemented in this backend yet; "
                "call .eval()")
        with torch.no_grad():
            if self.train_only_2d:
                all_heatmaps = (_inference.backbone_heatmaps(self.backbone, views) if views is not None
                                else list(input_heatmaps))
                pred = grid_centers = root_cubes = None
            else:
                pred, all_heatmaps, grid_centers, root_cubes = _inference.infer(
                    self, views, meta, input_heatmaps, self.USE_GT)
            device = all_heatmaps[0].device
            loss_2d = torch.zeros((), device=device)
            if targets_2d is not None:
                for t, w, o in zip(targets_2d, weights_2d, all_heatmaps):
                    loss_2d = loss_2d + _per_joint_mse(o, t.to(device), w.to(device))
                loss_2d = loss_2d / len(all_heatmaps)
            if self.train_only_2d:
                return loss_2d, all_heatmaps
            loss_3d = torch.zeros((), device=device)
            if targets_3d is not None and root_cubes is not None:
                loss_3d = _per_joint_mse(root_cubes, targets_3d.to(device))
            loss_cord = torch.zeros((), device=device)   # only accumulated when training (reference :90-99)
            return pred, all_heatmaps, grid_centers, loss_2d, loss_3d, loss_cord


def get_multi_person_pose_net(cfg, is_train=True):
    backbone = None
    if cfg.BACKBONE_MODEL:
        backbone = {"pose_resnet": pose_resnet}[cfg.BACKBONE_MODEL].get_pose_net(cfg, is_train=is_train)
    return MultiPersonPoseNet(backbone, cfg)
