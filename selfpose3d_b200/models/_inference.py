"""Batched inference shared by the two top-level models.

Same result as the reference's ``do_inference`` (``lib/models/multi_person_posenet_ssv.py:105-153``)
and the evaluation branch of ``MultiPersonPoseNet.forward`` (``multi_person_posenet.py:36-102``),
organised for the GPU instead of as Python loops: the V views go through the backbone as one
``V*B`` batch, the cameras are packed once, and every valid (sample, proposal) cube is un-projected
and regressed in one batched call -- one host synchronisation per forward (to read which proposal
slots are valid) instead of one per proposal slot.
"""
from __future__ import annotations

import torch

from .. import ops


def backbone_heatmaps(backbone, views):
    """list[V] of ``[B,3,H,W]`` -> list[V] of ``[B,J,h,w]`` channel-last views of one buffer."""
    V, B = len(views), int(views[0].shape[0])
    hm = backbone(torch.cat(list(views), dim=0))
    return [hm[v * B:(v + 1) * B] for v in range(V)]


def gt_grid_centers(meta, batch_size, num_cand, device):
    """Proposals from ground-truth roots (reference multi_person_posenet_ssv.py:123-131)."""
    num_person = meta[0]["num_person"]
    gc = torch.zeros(batch_size, num_cand, 5, device=device)
    gc[:, :, 0:3] = meta[0]["roots_3d"].float().to(device)
    gc[:, :, 3] = -1.0
    slot = torch.arange(num_cand, device=device)[None]
    mask = slot < num_person.to(device)[:, None]
    gc[:, :, 3] = torch.where(mask, slot.float().expand(batch_size, -1), gc[:, :, 3])
    gc[:, :, 4] = mask.float()
    return gc


def regress_valid(pose_net, all_heatmaps, cams, grid_centers, pred):
    """Fill ``pred[b, n, :, 0:3]`` for every proposal with ``flag >= 0`` (one sync to list them)."""
    valid = torch.nonzero(grid_centers[:, :, 3] >= 0)          # [n_valid, 2] (sample, slot)
    if valid.shape[0] == 0:
        return pred
    centers = grid_centers[valid[:, 0], valid[:, 1]].contiguous()
    joints = pose_net.regress(all_heatmaps, cams, centers, valid[:, 0].to(torch.int32).contiguous())
    pred[valid[:, 0], valid[:, 1], :, 0:3] = joints
    return pred


def infer(model, views, meta, input_heatmaps, use_root_gt, eval_rootnet_only=False, skip_pose=False):
    if views is not None:
        all_heatmaps = backbone_heatmaps(model.backbone, views)
    else:
        all_heatmaps = [h if h.is_cuda else h.cuda() for h in input_heatmaps]
    device = all_heatmaps[0].device
    B = int(all_heatmaps[0].shape[0])
    K, J = model.num_cand, model.num_joints

    # (TRAIN_ONLY_2D models have neither a root net nor a pose net: the heat-maps are the whole result)
    sub = getattr(model, "pose_net", None) or getattr(model, "root_net", None)
    if sub is None:
        return torch.zeros(B, K, J, 5, device=device), all_heatmaps, torch.zeros(B, K, 5, device=device), None
    img_size = sub.project_layer.img_size
    cams = ops.pack_cameras(meta, img_size).to(device, non_blocking=True)
    root_cubes = None
    if use_root_gt:
        grid_centers = gt_grid_centers(meta, B, K, device)
    else:
        root_cubes = model.root_net.root_volume(all_heatmaps, meta, cams=cams)
        grid_centers = model.root_net.proposal_layer(root_cubes, meta)

    pred = torch.zeros(B, K, J, 5, device=device)
    pred[:, :, :, 3:] = grid_centers[:, :, 3:].reshape(B, -1, 1, 2)
    if not (eval_rootnet_only or skip_pose):
        regress_valid(model.pose_net, all_heatmaps, cams, grid_centers, pred)
    return pred, all_heatmaps, grid_centers, root_cubes
