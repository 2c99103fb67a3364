"""CuboidProposalNetSoft -- the SSL variant of the root localiser.

Interface mirror of the reference's ``lib/models/cuboid_proposal_net_soft.py``
(``ProposalLayerSoft`` :18-68, ``CuboidProposalNetSoft`` :71-276): same constructor, the
4-tuple return ``(root_cubes, root_cubes_syn, target_cubes, grid_centers)`` and the
``get_grid_centres`` helper.  The inference branch runs on the sm_100a kernels; the
synthetic-root training branch ``train_rootnet`` (:151-241) is a "next" row of the scope table
 and raises here.
"""
from __future__ import annotations

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

    def get_grid_centres(self, all_heatmaps, meta, flip_xcoords):
        root_cubes = self.root_volume(all_heatmaps, meta, flip_xcoords)
        return root_cubes, self.proposal_layer(root_cubes, meta, None)

    def train_rootnet(self, *args, **kwargs):
        raise NotImplementedError(
            "selfpose3d_b200: the synthetic-root RootNet training branch (reference "
            "cuboid_proposal_net_soft.py:151-241) is not part of this backend yet")

    def forward(self, all_heatmaps, meta, flip_xcoords=None):
        root_cubes, grid_centers = self.get_grid_centres(all_heatmaps, meta, flip_xcoords)
        if self.rootnet_train_synth and self.training:
            self.train_rootnet()
        return root_cubes, None, None, grid_centers
