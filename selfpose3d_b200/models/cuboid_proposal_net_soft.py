"""CuboidProposalNetSoft -- the SSL variant of the root localiser.

Interface mirror of the reference's ``lib/models/cuboid_proposal_net_soft.py``
(``ProposalLayerSoft`` :18-68, ``CuboidProposalNetSoft`` :71-276): same constructor, the
4-tuple return ``(root_cubes, root_cubes_syn, target_cubes, grid_centers)`` and the
``get_grid_centres`` helper.  The synthetic-root training branch ``train_rootnet`` (:151-241) draws random roots,
builds their 3-D Gaussian target volume and their noisy 2-D heat-maps in every view, and trains the V2VNet on that
pair; its random draws are made in the reference's order, so a seeded CPU run reproduces the reference's.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .cuboid_proposal_net import CuboidProposalNet, ProposalLayer


class ProposalLayerSoft(ProposalLayer):
    """No ground-truth matching: ``flag = (score > threshold) - 1`` always (reference :54-68)."""

    def forward(self, root_cubes, meta, grids=None):
        training, self.training = self.training, False
        try:
            return super().forward(root_cubes, None)
        finally:
            self.training = training


class CuboidProposalNetSoft(CuboidProposalNet):
    def __init__(self, cfg):
        nn.Module.__init__(self)
        from .project_layer import ProjectLayer
        from .v2v_net import V2VNet
        self.grid_size = cfg.MULTI_PERSON.SPACE_SIZE
        self.cube_size = cfg.MULTI_PERSON.INITIAL_CUBE_SIZE
        self.grid_center = cfg.MULTI_PERSON.SPACE_CENTER
        self.root_id = cfg.DATASET.ROOTIDX          # the Soft variant reads ROOTIDX, not ROOTIDX_PSEUDO (:77)
        self.rootnet_roothm = cfg.NETWORK.ROOTNET_ROOTHM
        self.rootnet_train_synth = cfg.NETWORK.ROOTNET_TRAIN_SYNTH
        self.max_num_people = cfg.MULTI_PERSON.MAX_PEOPLE_NUM
        self.rootnet_syn_range = cfg.NETWORK.ROOTNET_SYN_RANGE

        self.project_layer = ProjectLayer(cfg)
        self.v2v_net = V2VNet(1 if self.rootnet_roothm else cfg.NETWORK.NUM_JOINTS, 1)
        self.proposal_layer = ProposalLayerSoft(cfg)
        if self.rootnet_train_synth:
            # voxel-centre axes of the root grid, the box the synthetic roots are drawn from, heat-map pixel grids
            # (reference :90-127; non-persistent buffers: not part of the state dict)
            self.cur_sigma = 200.0
            axes = [np.linspace(-self.grid_size[a] / 2, self.grid_size[a] / 2, self.cube_size[a]) + self.grid_center[a]
                    for a in range(3)]
            rng = cfg.NETWORK.ROOTNET_SYN_RANGE
            self.syn_box = [(float(axes[a].min() + rng[a][0]), float(axes[a].max() + rng[a][1])) for a in range(3)]
            for name, ax in zip(("grid1Dx", "grid1Dy", "grid1Dz"), axes):
                self.register_buffer(name, torch.from_numpy(ax).to(torch.float32), persistent=False)
            hw, hh = int(cfg.NETWORK.HEATMAP_SIZE[0]), int(cfg.NETWORK.HEATMAP_SIZE[1])
            yy, xx = torch.meshgrid(torch.arange(hh, dtype=torch.float32), torch.arange(hw, dtype=torch.float32), indexing="ij")
            self.register_buffer("hm_xx", xx.view(1, 1, hh, hw), persistent=False)
            self.register_buffer("hm_yy", yy.view(1, 1, hh, hw), persistent=False)

    def get_grid_centres(self, all_heatmaps, meta, flip_xcoords):
        root_cubes = self.root_volume(all_heatmaps, meta, flip_xcoords)
        return root_cubes, self.proposal_layer(root_cubes, meta, None)

    def synthetic_roots(self, batch_size):
        """Random root positions ``[B, R, 3]`` in the reference's draw order (:155-163): the number of roots, x, y, one
        z per sample shared by its roots, a 50 mm z jitter per root."""
        n = int(torch.randint(1, self.max_num_people, (1,)).item())
        (x0, x1), (y0, y1), (z0, z1) = self.syn_box
        x = (x1 - x0) * torch.rand(batch_size, n, 1) + x0
        y = (y1 - y0) * torch.rand(batch_size, n, 1) + y0
        z = ((z1 - z0) * torch.rand(batch_size, 1, 1) + z0).expand(-1, n, 1)
        z = z + torch.randn_like(z) * 50
        return torch.cat((x, y, z), -1).to(device=self.grid1Dx.device, dtype=torch.float32)

    def target_volume(self, roots):
        """``[B, R, 3]`` -> ``[B, X, Y, Z]``: the maximum over the roots of a Gaussian (sigma ``cur_sigma`` mm) cut off at
        3 sigma along every axis (reference :167-203)."""
        s = self.cur_sigma
        out = []
        for pts in roots:
            vol = torch.zeros(len(self.grid1Dx), len(self.grid1Dy), len(self.grid1Dz), device=roots.device)
            for mu in pts:
                d2, inside = 0.0, True
                for a, ax in enumerate((self.grid1Dx, self.grid1Dy, self.grid1Dz)):
                    shape = [1, 1, 1]
                    shape[a] = -1
                    win = ((ax >= mu[a] - 3 * s) & (ax <= mu[a] + 3 * s)).view(shape)
                    inside = inside & win
                    d2 = d2 + ((ax - mu[a]) ** 2).view(shape)
                vol = torch.maximum(vol, torch.exp(-d2 / (2 * s ** 2)) * inside)
            out.append(torch.clip(vol, 0, 1))
        return torch.stack(out, 0)

    def train_rootnet(self, batch_size, meta, pred_hms, flip_xcoords=False):
        """Synthetic-root step (reference :151-241): random roots -> target volume and noisy per-view root heat-maps
        (both without gradient) -> un-projection + V2VNet.  Returns ``(root_cubes_syn, target_cubes)``."""
        from ._ssl_train import project_to_views, GAUSS_SIGMA, IMAGE_TO_HEATMAP
        with torch.no_grad():
            roots = self.synthetic_roots(batch_size)
            target = self.target_volume(roots)
            pts = [roots[b][None] for b in range(batch_size)]                    # one "person" of R joints per sample
            maps = []
            for m in meta:
                kps = project_to_views(pts, m["camera"], meta[0]["trans"])       # list[B] of [1, R, 2]
                per_sample = []
                for kp in kps:
                    x = (kp[0, :, 0] / IMAGE_TO_HEATMAP)[:, None, None, None]
                    y = (kp[0, :, 1] / IMAGE_TO_HEATMAP)[:, None, None, None]
                    g = torch.exp(-(((self.hm_xx - x) / GAUSS_SIGMA) ** 2) / 2 - (((self.hm_yy - y) / GAUSS_SIGMA) ** 2) / 2)
                    hm = torch.clip(g.sum(0), min=0.0, max=1.0)                  # [1, h, w]
                    per_sample.append(torch.clip(hm + 0.02 * torch.randn_like(hm), min=0.0, max=1.0))
                maps.append(torch.stack(per_sample, 0))                          # [B, 1, h, w]
        return self.root_volume(maps, meta, flip_xcoords, select_root=not self.rootnet_roothm), target

    def forward(self, all_heatmaps, meta, flip_xcoords=None):
        root_cubes, grid_centers = self.get_grid_centres(all_heatmaps, meta, flip_xcoords)
        if self.rootnet_train_synth and self.training:
            syn, target = self.train_rootnet(int(all_heatmaps[0].shape[0]), meta, all_heatmaps, flip_xcoords)
            return root_cubes, syn, target, grid_centers
        return root_cubes, None, None, grid_centers
