"""MultiPersonPoseNet -- supervised (VoxelPose) top-level model.

Interface mirror of the reference's ``lib/models/multi_person_posenet.py:20-111``: same
constructor, ``forward(views, meta, targets_2d, weights_2d, targets_3d, input_heatmaps)`` and
6-tuple return ``(pred, all_heatmaps, grid_centers, loss_2d, loss_3d, loss_cord)`` (or
``(loss_2d, all_heatmaps)`` with ``TRAIN_ONLY_2D``).  Evaluation mode runs on the sm_100a kernels;
the loss values are the reference's trivial MSE reductions evaluated with torch under ``no_grad``
(SURVEY.md section 2.1 row 11).  In ``.train()`` mode ``forward`` is the reference's supervised training step
(``_forward_train`` below): the nets run their training paths (batch-statistics BatchNorm, gradients through the
backward kernels).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import pose_resnet
from . import _inference, pose_regression_net
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
            return self._forward_train(views, meta, targets_2d, weights_2d, targets_3d, input_heatmaps)
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


def _per_joint_l1(output, target, weight):
    """PerJointL1Loss with target weights (reference lib/core/loss.py:61-78)."""
    B, J = output.shape[:2]
    return torch.mean(torch.abs(output.reshape(B, J, -1) * weight - target.reshape(B, J, -1) * weight))


def _forward_train(self, views, meta, targets_2d, weights_2d, targets_3d, input_heatmaps):
    """The supervised training forward (reference multi_person_posenet.py:36-102 with ``self.training``): root net and
    pose net run their training paths (batch-statistics BatchNorm, gradients through the backward kernels), the pose
    net once per proposal slot on the rows matched to a ground-truth person -- as the reference does, so that the
    batch statistics see the same batches.  The three losses are plain reductions over small tensors.  A backbone
    in ``.train()`` mode runs once per view (the reference's batches, :38-41) on its own training path; a frozen one
    (``.eval()``) runs the fused inference kernels on all views at once and its heat-maps enter as constants."""
    if views is not None:
        if self.backbone.training:
            all_heatmaps = self.backbone.forward_views(views)      # one pass, one BatchNorm statistic group per view
        else:
            with torch.no_grad():
                all_heatmaps = _inference.backbone_heatmaps(self.backbone, views)
    else:
        all_heatmaps = [h if h.is_cuda else h.cuda() for h in input_heatmaps]
    device = all_heatmaps[0].device
    B = int(all_heatmaps[0].shape[0])
    loss_2d = torch.zeros((), device=device)
    if targets_2d is not None:
        for t, w, o in zip(targets_2d, weights_2d, all_heatmaps):
            loss_2d = loss_2d + _per_joint_mse(o, t.to(device), w.to(device))
        loss_2d = loss_2d / len(all_heatmaps)
    if self.train_only_2d:
        return loss_2d, all_heatmaps
    loss_3d = torch.zeros((), device=device)
    if self.USE_GT:
        grid_centers = _inference.gt_grid_centers(meta, B, self.num_cand, device)
    else:
        root_cubes, grid_centers = self.root_net(all_heatmaps, meta)
        if targets_3d is not None:
            loss_3d = _per_joint_mse(root_cubes, targets_3d.to(device))
    pred = torch.zeros(B, self.num_cand, self.num_joints, 5, device=device)
    pred[:, :, :, 3:] = grid_centers[:, :, 3:].reshape(B, -1, 1, 2)
    loss_cord = torch.zeros((), device=device)
    count = 0
    has_gt = "joints_3d" in meta[0] and "joints_3d_vis" in meta[0]
    flags = grid_centers[:, :, 3].detach().cpu()                       # one host sync (the reference: one per slot)
    batched = None
    if pose_regression_net.SLOT_BATCH:
        # all slots in one pass, every slot's rows one BatchNorm statistic group (= the reference's per-slot calls)
        got = self.pose_net.regress_slots(all_heatmaps, meta, grid_centers, flags)
        if got is not None:
            joints, slots, samples = got
            idx = (torch.tensor(samples, device=device), torch.tensor(slots, device=device))
            batched = torch.zeros(B, self.num_cand, self.num_joints, 3, device=device).index_put(idx, joints)
    if has_gt:
        gt_3d = meta[0]["joints_3d"].float().to(device)
        vis = meta[0]["joints_3d_vis"].float().to(device)
    for n in range(self.num_cand):
        if not bool((flags[:, n] >= 0).any()):
            continue
        single_pose = batched[:, n] if batched is not None else self.pose_net(all_heatmaps, meta, grid_centers[:, n])
        pred[:, n, :, 0:3] = single_pose.detach()
        if has_gt:
            for i in range(B):
                if flags[i, n] >= 0:
                    g = int(flags[i, n])
                    count += 1
                    term = _per_joint_l1(single_pose[i:i + 1], gt_3d[i:i + 1, g], vis[i:i + 1, g, :, 0:1])
                    loss_cord = (loss_cord * (count - 1) + term) / count
    return pred, all_heatmaps, grid_centers, loss_2d, loss_3d, loss_cord


MultiPersonPoseNet._forward_train = _forward_train


def get_multi_person_pose_net(cfg, is_train=True):
    backbone = None
    if cfg.BACKBONE_MODEL:
        backbone = {"pose_resnet": pose_resnet}[cfg.BACKBONE_MODEL].get_pose_net(cfg, is_train=is_train)
    return MultiPersonPoseNet(backbone, cfg)
