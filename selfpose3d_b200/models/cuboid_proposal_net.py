"""CuboidProposalNet -- root localiser: un-project -> V2VNet -> 3-D NMS / top-K proposals.

Interface mirror of the reference's ``lib/models/cuboid_proposal_net.py:15-122``
(``ProposalLayer``, ``CuboidProposalNet``); state-dict keys ``v2v_net.*`` identical.  The whole
forward is three kernel families with no intermediate layout change: ``sp3d_unproject_fwd``
writes the channel-last cube the convolutions read, the 1-channel V2V output *is* the
``[B,X,Y,Z]`` score volume, and ``sp3d_nms_topk3d`` turns it into ``grid_centers`` in one launch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .project_layer import ProjectLayer
from .v2v_net import V2VNet


def _is_f64_config(*values):
    """The reference builds ``torch.tensor(cfg...)`` (cuboid_proposal_net.py:18-20): float64 numpy
    config defaults make ``get_real_loc`` run in float64, YAML lists in float32."""
    return any(isinstance(v, np.ndarray) and v.dtype == np.float64 for v in values)


class ProposalLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.loc_f64 = _is_f64_config(cfg.MULTI_PERSON.SPACE_SIZE, cfg.MULTI_PERSON.SPACE_CENTER)
        self.grid_size = torch.tensor(cfg.MULTI_PERSON.SPACE_SIZE)
        self.cube_size = torch.tensor(cfg.MULTI_PERSON.INITIAL_CUBE_SIZE)
        self.grid_center = torch.tensor(cfg.MULTI_PERSON.SPACE_CENTER)
        self.num_cand = cfg.MULTI_PERSON.MAX_PEOPLE_NUM
        self.root_id = cfg.DATASET.ROOTIDX
        self.num_joints = cfg.NETWORK.NUM_JOINTS
        self.threshold = cfg.MULTI_PERSON.THRESHOLD

    def get_real_loc(self, index):
        """Voxel index -> world mm, ``idx / (n-1) * size + centre - size/2`` (reference :42-52)."""
        device = index.device
        cube_size = self.cube_size.to(device=device, dtype=torch.float)
        grid_size = self.grid_size.to(device=device)
        grid_center = self.grid_center.to(device=device)
        return index.float() / (cube_size - 1) * grid_size + grid_center - grid_size / 2.0

    def filter_proposal(self, topk_index, gt_3d, num_person):
        """Nearest ground-truth root within 500 mm per candidate, else -1 (reference :25-40)."""
        dist = torch.cdist(topk_index.float(), gt_3d.float().to(topk_index.device))          # [B,K,G]
        valid = torch.arange(gt_3d.shape[1], device=dist.device)[None, None, :] < num_person.to(dist.device)[:, None, None]
        dist = torch.where(valid, dist, torch.full_like(dist, float("inf")))
        min_dist, min_gt = dist.min(dim=-1)
        cand2gt = min_gt.float()
        cand2gt[min_dist > 500.0] = -1.0
        return cand2gt

    def forward(self, root_cubes, meta):
        root_cubes = root_cubes.detach().float().contiguous()
        gc = ops.nms_topk(root_cubes, self.num_cand, self.threshold, self.grid_size.tolist(),
                          self.grid_center.tolist(), loc_f64=self.loc_f64)
        if self.training and meta is not None and ("roots_3d" in meta[0] and "num_person" in meta[0]):
            # supervised training/validation of the pose net: match proposals to ground truth (reference :69-77)
            gc[:, :, 3] = self.filter_proposal(gc[:, :, 0:3], meta[0]["roots_3d"].float(), meta[0]["num_person"])
        return gc


class CuboidProposalNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.grid_size = cfg.MULTI_PERSON.SPACE_SIZE
        self.cube_size = cfg.MULTI_PERSON.INITIAL_CUBE_SIZE
        self.grid_center = cfg.MULTI_PERSON.SPACE_CENTER
        self.rootnet_roothm = cfg.NETWORK.ROOTNET_ROOTHM
        self.root_id = cfg.DATASET.ROOTIDX_PSEUDO

        self.project_layer = ProjectLayer(cfg)
        self.v2v_net = V2VNet(1 if self.rootnet_roothm else cfg.NETWORK.NUM_JOINTS, 1)
        self.proposal_layer = ProposalLayer(cfg)

    def root_volume(self, all_heatmaps, meta, flip_xcoords=None, cams=None, select_root=True):
        """heat-maps -> ``root_cubes [B,X,Y,Z]`` (un-projection + V2VNet), layout-native.  ``select_root=False``: the
        maps already hold exactly the channels the root net reads (synthetic root heat-maps)."""
        if self.rootnet_roothm and select_root:   # only the root joint's heat-map feeds the root net (reference :103-108)
            hms = [a[:, self.root_id:self.root_id + 1] for a in all_heatmaps]
        else:
            hms = all_heatmaps
        device = hms[0].device
        B = int(hms[0].shape[0])
        if cams is None:
            cams = ops.pack_cameras(meta, self.project_layer.img_size, flip_xcoords).to(device, non_blocking=True)
        centers, _ = self.project_layer.centers_tensor([list(self.grid_center)], B, device)
        if self.v2v_net.training:
            # training path (float32): differentiable un-projection -> V2VNet with batch statistics; gradients reach the
            # heat-maps and the net's parameters
            from .. import autograd as ag
            if ops.volume_dtype() != torch.float32:
                raise ValueError("the training path runs on float32 volumes (ops.set_volume_dtype(torch.float32))")
            C = int(hms[0].shape[1])
            spec = ([float(v) for v in self.grid_size], [int(v) for v in self.cube_size], self.project_layer.img_size,
                    self.project_layer.heatmap_size, C, ops.round_up(C, 4))
            cubes = ag.Unproject.apply(cams, centers, None, spec, *[h.float().contiguous() for h in hms])
            return self.v2v_net.forward_cl(cubes)[..., 0].contiguous()
        bf16 = ops.volume_dtype() == torch.bfloat16
        cubes, _ = self.project_layer.project_cl(hms, cams, centers, False, self.grid_size, self.cube_size,
                                                 dtype="split" if ops.use_split() else ops.volume_dtype(),
                                                 c_pitch=ops.round_up(hms[0].shape[1], 16) if bf16 else None)
        root = self.v2v_net.forward_cl(cubes, out_pitch=1)
        return root.view(root.shape[0], root.shape[1], root.shape[2], root.shape[3])

    def forward(self, all_heatmaps, meta, flip_xcoords=None):
        root_cubes = self.root_volume(all_heatmaps, meta, flip_xcoords)
        grid_centers = self.proposal_layer(root_cubes, meta)
        return root_cubes, grid_centers
