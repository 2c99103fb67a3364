"""The self-supervised training forward of ``MultiPersonPoseNetSSV`` (reference
``lib/models/multi_person_posenet_ssv.py:222-501``).

Three augmented view sets enter: sets 1 and 2 carry the affine / flip augmentation and train the pose net against
each other (the joints regressed from one set are re-projected into the other set's views, rendered as Gaussian
heat-maps and compared with that set's pseudo heat-maps, optionally weighted by the attention net), set 3 is
un-augmented and feeds the root net.  The networks run their training paths (``selfpose3d_b200.autograd``: every
forward / backward step is one of the kernels); what is restated here is the loss assembly around them -- small
tensors (a few people x 15 joints, ``V x B x J x h x w`` heat-maps): plain tensor expressions for now, the fused
re-projection + rendering kernel is the next row of SURVEY.md section 8(f).

The synthetic-root RootNet branch (``NETWORK.ROOTNET_TRAIN_SYNTH``) lives in ``cuboid_proposal_net_soft.py``.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import _inference, pose_regression_net

GAUSS_SIGMA = 3.0        # reference :419 -- rendering sigma in heat-map pixels
IMAGE_TO_HEATMAP = 4.0   # reference :416 -- network-input pixels per heat-map pixel (hard-coded there too)


def _heatmaps(backbone, views, given):
    """One backbone call per view, as the reference (:227-277): the BatchNorm batches are the per-view batches."""
    if views is None:
        return [h if h.is_cuda else h.cuda() for h in given]
    return backbone.forward_views(views)      # (one pass, one BatchNorm statistic group per view)


def project_to_views(poses, camera, trans):
    """World joints -> network-input pixels of one view (``cameras.project_pose_batch``,
    ``lib/utils/cameras.py:58-108,116-118``).  ``poses``: list over samples of ``[P_b, J, 3]``; ``camera``: the
    collated camera dict of the view (``R [B,3,3]``, ``T [B,3,1]``, ``f, c [B,2,1]``, ``k [B,3,1]``, ``p [B,2,1]``);
    ``trans [B,2,3]`` the augmentation affine.  Returns a list of ``[P_b, J, 2]``."""
    out = []
    for b, x in enumerate(poses):
        dev = x.device
        def cam(key, *shape):
            return camera[key][b].to(device=dev, dtype=x.dtype).reshape(*shape)
        R, T, f, c, k, p = cam("R", 3, 3), cam("T", 1, 1, 3), cam("f", 1, 1, 2), cam("c", 1, 1, 2), cam("k", 3), cam("p", 2)
        xc = torch.matmul(x - T, R.transpose(0, 1))                                                        # [P,J,3]
        y = xc[..., :2] / (xc[..., 2:3] + 1e-5)
        r2 = (y ** 2).sum(-1, keepdim=True)
        radial = 1 + k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3
        tan = p[0] * y[..., 1:2] + p[1] * y[..., 0:1]
        y = y * (radial + 2 * tan) + torch.stack([p[1], p[0]]).reshape(1, 1, 2) * r2
        pix = f * y + c                                                                                    # [P,J,2]
        A = trans[b].to(device=dev, dtype=x.dtype)
        out.append(torch.matmul(pix, A[:, :2].transpose(0, 1)) + A[:, 2])
    return out


def render_gaussians(kps_views, xx, yy):
    """Joint pixels -> heat-maps (reference :410-448): per (view, sample) a Gaussian of sigma 3 heat-map pixels around
    every person's joint, summed over the people and clipped to [0, 1].  ``kps_views``: list over views of lists over
    samples of ``[P_b, J, 2]``; ``xx, yy``: ``[1,1,h,w]`` pixel-index grids.  Returns ``[V, B, J, h, w]``."""
    # the fused rendering kernel (sp3d_gauss_render_fwd / _bwd) on the device; the tensor expression below is what the
    # CPU wiring tests (kernels emulated) evaluate, and SP3D_RENDER_KERNEL=0 keeps it for A/B checks
    if kps_views[0][0].is_cuda and os.environ.get("SP3D_RENDER_KERNEL", "1") != "0":
        return _render_gaussians_kernel(kps_views, xx)
    views = []
    for kps_samples in kps_views:
        maps = []
        for kp in kps_samples:
            x = (kp[..., 0] / IMAGE_TO_HEATMAP)[..., None, None]
            y = (kp[..., 1] / IMAGE_TO_HEATMAP)[..., None, None]
            g = torch.exp(-(((xx - x) / GAUSS_SIGMA) ** 2) / 2 - (((yy - y) / GAUSS_SIGMA) ** 2) / 2)          # [P,J,h,w]
            maps.append(torch.clip(g.sum(0), min=0.0, max=1.0))
        views.append(torch.stack(maps, 0))
    return torch.stack(views, 0)


def _render_gaussians_kernel(kps_views, xx):
    """The same through ``sp3d_gauss_render_fwd/bwd``: the ragged people lists are padded to ``[V,B,P,J,2]``."""
    from .. import autograd as ag
    counts = [int(kp.shape[0]) for kp in kps_views[0]]
    P = max(max(counts), 1)
    J = int(kps_views[0][0].shape[1])
    dev = kps_views[0][0].device
    padded = torch.stack([torch.stack([torch.cat([kp, kp.new_zeros(P - kp.shape[0], J, 2)], 0) for kp in samples], 0)
                          for samples in kps_views], 0)
    n_people = torch.tensor(counts, device=dev, dtype=torch.int32)
    return ag.RenderGaussians.apply(padded.float(), n_people, (int(xx.shape[-2]), int(xx.shape[-1])),
                                    1.0 / IMAGE_TO_HEATMAP, GAUSS_SIGMA)


def hungarian_l1(kps_views, meta, width, height, drop_worst):
    """2-D joint L1 loss under the optimal pseudo-label <-> prediction assignment (reference ``l1_matching_loss``
    :155-194): per (view, sample) the summed cost of the Hungarian matching between the re-projected people and the
    pseudo 2-D poses ``meta[v]['joints']`` (people whose joints are all zero are padding), coordinates normalised by
    the network-input size; mean over (view, sample), or -- ``drop_worst`` (``L1_ATTN``) -- the mean without the
    largest entry.

    The reference synchronises with the host once per (view, sample) (``d_matrix.cpu()`` at :182).  Here all ``V * B``
    cost matrices are built on the device in one padded ``[V*B, G, P]`` tensor, cross to the host in ONE copy, are
    assigned there (``scipy.optimize.linear_sum_assignment``, a few people per matrix), and the matched entries are
    gathered back with one index tensor."""
    from scipy.optimize import linear_sum_assignment
    V, B = len(meta), len(kps_views[0])
    dev = kps_views[0][0].device
    size = torch.tensor([float(width), float(height)], device=dev)
    joints = torch.stack([meta[v]["joints"].to(dev) for v in range(V)])                 # [V, B, G, J, 2]
    vis = torch.stack([meta[v]["joints_vis"].to(dev) for v in range(V)])                # [V, B, G, J, 2]
    G = int(joints.shape[2])
    n_pred = [int(kp.shape[0]) for kp in kps_views[0]]                                  # people per sample (same in every view)
    P = max(max(n_pred), 1)
    J = int(joints.shape[3])
    pred = torch.stack([torch.stack([torch.cat([kp, kp.new_zeros(P - kp.shape[0], J, 2)], 0) for kp in samples])
                        for samples in kps_views])                                      # [V, B, P, J, 2]
    target = joints / size.to(joints.dtype)
    cost = (((pred / size)[:, :, None] - target[:, :, :, None]) * vis[:, :, :, None]).abs().mean((-1, -2)).to(torch.float32)
    cost = cost.reshape(V * B, G, P)                                                    # [V*B, G, P]
    n_gt = (joints.sum(-1).sum(-1) != 0).sum(-1).reshape(V * B)                         # pseudo people per (view, sample)
    host_cost, host_gt = cost.detach().cpu().numpy(), n_gt.cpu().numpy()                # the one device -> host copy
    idx_m, idx_r, idx_c = [], [], []
    for m in range(V * B):
        g, p_ = int(host_gt[m]), n_pred[m % B]
        if g == 0 or p_ == 0:
            continue
        rows, cols = linear_sum_assignment(host_cost[m, :g, :p_])
        idx_m += [m] * len(rows)
        idx_r += rows.tolist()
        idx_c += cols.tolist()
    per = torch.zeros(V * B, device=dev)
    if idx_m:
        sel = torch.tensor([idx_m, idx_r, idx_c], device=dev)
        per = per.index_add(0, sel[0], cost[sel[0], sel[1], sel[2]])
    if drop_worst:
        keep = torch.ones_like(per)
        keep[torch.argmax(per)] = 0.0
        return (per * keep).sum() / (V * B - 1)
    return per.mean()


def forward_train(self, views1, meta1, targets_2d1, weights_2d1, targets_3d1, input_heatmaps1,
                  views2, meta2, targets_2d2, weights_2d2, targets_3d2, input_heatmaps2,
                  views3, meta3, targets_2d3, weights_2d3, targets_3d3, input_heatmaps3, epoch):
    heatmaps3 = _heatmaps(self.backbone, views3, input_heatmaps3)
    attn1 = attn2 = None
    if self.WITH_ATTN:
        if views1 is not None:
            attn1 = torch.stack(self.attn.forward_views(views1), 0)
        if views2 is not None:
            attn2 = torch.stack(self.attn.forward_views(views2), 0)
    heatmaps1 = _heatmaps(self.backbone, views1, input_heatmaps1)
    heatmaps2 = _heatmaps(self.backbone, views2, input_heatmaps2)
    device = heatmaps1[0].device
    B = int(heatmaps1[0].shape[0])
    K, J = self.num_cand, self.num_joints
    # a zero that stays attached to the graph.  The reference uses `pose_net.v2v_net(zero_tensor).mean() * 0`
    # (multi_person_posenet_ssv.py:91-97,342-349): it always reaches the pose net's parameters, so that every loss it
    # stands in for requires grad and those parameters receive zero (not None) gradients
    anchor = getattr(self, "pose_net", None) or getattr(self, "root_net", None) or self.backbone
    zero = sum(p.sum() for p in anchor.parameters()) * 0.0 + heatmaps3[0].sum() * 0.0

    losses = {}
    t1 = torch.stack([t.to(device) for t in targets_2d1]) if targets_2d1 is not None else None
    t2 = torch.stack([t.to(device) for t in targets_2d2]) if targets_2d2 is not None else None
    if t1 is not None and t2 is not None:
        t3 = torch.stack([t.to(device) for t in targets_2d3])
        losses["loss_2d"] = (F.mse_loss(t1, torch.stack(heatmaps1)) + F.mse_loss(t2, torch.stack(heatmaps2))
                             + F.mse_loss(t3, torch.stack(heatmaps3))) / 3.0
    else:
        losses["loss_2d"] = zero
    if self.train_only_2d:
        return None, heatmaps3, None, losses

    if self.use_root_gt:
        grid_centers = _inference.gt_grid_centers(meta3, B, K, device)
    elif self.freeze_rootnet:
        grid_centers = self.root_net(heatmaps3, meta3, flip_xcoords=meta3[0]["hflip"])[3]
    elif self.rootnet_train_synth:
        # synthetic-root supervision on all three sets + consistency of the real volumes with set 3's (:312-330)
        main1, syn1, tgt1, _ = self.root_net(heatmaps1, meta1, flip_xcoords=meta1[0]["hflip"])
        main2, syn2, tgt2, _ = self.root_net(heatmaps2, meta2, flip_xcoords=meta2[0]["hflip"])
        main3, syn3, tgt3, grid_centers = self.root_net(heatmaps3, meta3, flip_xcoords=meta3[0]["hflip"])
        losses["loss_root_syn"] = self.weight_root_syn * (F.mse_loss(syn1, tgt1) + F.mse_loss(syn2, tgt2)
                                                          + F.mse_loss(syn3, tgt3))
        if self.root_reg_loss:
            main3 = main3.detach()
            losses["loss_root_reg"] = self.weight_root_reg * (F.mse_loss(main1, main3) + F.mse_loss(main2, main3))
    else:
        cubes1 = self.root_net(heatmaps1, meta1, flip_xcoords=meta1[0]["hflip"])[0]
        cubes2 = self.root_net(heatmaps2, meta2, flip_xcoords=meta2[0]["hflip"])[0]
        grid_centers = self.root_net(heatmaps3, meta3, flip_xcoords=meta3[0]["hflip"])[3]
        losses["loss_root_reg"] = (F.mse_loss(cubes1, targets_3d1.to(device)) + F.mse_loss(cubes2, targets_3d2.to(device)))
    if self.train_only_rootnet:
        return None, heatmaps3, grid_centers, losses

    if epoch < self.init_train_epochs_rootnet:
        losses["loss_pose3d_ssv"] = zero
        return None, heatmaps3, grid_centers, losses

    flags = grid_centers[:, :, 3].detach().cpu()          # one host sync (the reference: one per slot and view set)
    sets = [(heatmaps1, meta1)] if self.single_aug_training_posenet else [(heatmaps1, meta1), (heatmaps2, meta2)]
    preds = []
    for hms, meta in sets:
        pred = torch.zeros(B, K, J, 5, device=device)
        pred[:, :, :, 3:] = grid_centers[:, :, 3:].reshape(B, -1, 1, 2)
        if pose_regression_net.SLOT_BATCH:
            # all slots in one pass, every slot's rows one BatchNorm statistic group (= the reference's per-slot calls)
            got = self.pose_net.regress_slots(hms, meta, grid_centers, flags, flip_xcoords=meta[0]["hflip"])
            if got is not None:
                joints, slots, samples = got
                idx = (torch.tensor(samples, device=device), torch.tensor(slots, device=device))
                xyz = torch.zeros(B, K, J, 3, device=device).index_put(idx, joints)
                pred = torch.cat([xyz, pred[..., 3:]], dim=-1)
        else:
            for n in range(K):
                if bool((flags[:, n] >= 0).any()):
                    pred[:, n, :, 0:3] = self.pose_net(hms, meta, grid_centers[:, n], flip_xcoords=meta[0]["hflip"])
        preds.append(pred)
    pred_out = preds[-1].detach().clone()
    n_valid = [int((flags[b] >= 0).sum()) for b in range(B)]
    people = [[pred[b, :n_valid[b], :, :3] for b in range(B)] for pred in preds]
    cams = [m["camera"] for m in meta1]                     # the un-augmented cameras are shared by both sets
    xx, yy = self.hm_xx.to(device), self.hm_yy.to(device)

    if self.single_aug_training_posenet:
        if n_valid[0] > 0:
            kps11 = [project_to_views(people[0], cam, meta1[0]["trans"]) for cam in cams]
            rendered = render_gaussians(kps11, xx, yy)
            losses["loss_pose3d_ssv"] = F.mse_loss(t1, rendered) if t1 is not None else zero
        else:
            losses["loss_pose3d_ssv"] = zero
        return pred_out, heatmaps3, grid_centers, losses

    use_l1 = self.USE_L1 and epoch >= self.L1_EPOCH
    if n_valid[0] > 0:
        kps12 = [project_to_views(people[0], cam, meta2[0]["trans"]) for cam in cams]    # set-1 poses in set 2's views
        kps21 = [project_to_views(people[1], cam, meta1[0]["trans"]) for cam in cams]    # set-2 poses in set 1's views
        hm21, hm12 = render_gaussians(kps21, xx, yy), render_gaussians(kps12, xx, yy)
        loss1 = loss2 = torch.zeros((), device=device)
        if t1 is not None:
            loss1 = (F.mse_loss(t1, hm21, reduction="none") * attn1).mean() if self.WITH_ATTN else F.mse_loss(t1, hm21)
        if t2 is not None:
            loss2 = (F.mse_loss(t2, hm12, reduction="none") * attn2).mean() if self.WITH_ATTN else F.mse_loss(t2, hm12)
        losses["loss_pose3d_ssv"] = loss1 + loss2
        if self.WITH_ATTN:
            losses["loss_attn_ssv"] = (F.mse_loss(attn1, torch.ones_like(attn1))
                                       + F.mse_loss(attn2, torch.ones_like(attn2))) * self.attn_weight
        if use_l1:
            losses["loss_pose3d_l1_ssv"] = (hungarian_l1(kps12, meta2, self.width, self.height, self.L1_ATTN)
                                            + hungarian_l1(kps21, meta1, self.width, self.height, self.L1_ATTN)) * self.L1_WEIGHT
    else:
        if self.WITH_ATTN:
            losses["loss_attn_ssv"] = (F.mse_loss(attn1, torch.ones_like(attn1))
                                       + F.mse_loss(attn2, torch.ones_like(attn2))) * 0.0
        if use_l1:
            losses["loss_pose3d_l1_ssv"] = zero
        losses["loss_pose3d_ssv"] = zero
    return pred_out, heatmaps3, grid_centers, losses
