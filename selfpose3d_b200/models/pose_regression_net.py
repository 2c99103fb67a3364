"""PoseRegressionNet -- per-person cube: un-project -> V2VNet -> soft-argmax joints.

Interface mirror of the reference's ``lib/models/pose_regression_net.py:13-53``
(``SoftArgmaxLayer``, ``PoseRegressionNet``); state-dict keys ``v2v_net.*`` identical.
``forward`` keeps the reference contract (one proposal slot across the batch, rows with
``flag < 0`` skipped and left zero); ``regress`` is the batched entry that takes any number of
cube centres at once (all proposal slots of all samples), which is how the top-level model
calls it.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import ops
from .project_layer import ProjectLayer
from .v2v_net import V2VNet


# Training forwards run all proposal slots of a step through the pose net in ONE pass (grouped BatchNorm statistics keep
# the reference's per-slot batches); SP3D_SLOT_BATCH=0 keeps one pass per slot (A/B checks).
SLOT_BATCH = os.environ.get("SP3D_SLOT_BATCH", "1") != "0"


class SoftArgmaxLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.beta = cfg.NETWORK.BETA

    def forward(self, x, grids):
        """``x [B,C,X,Y,Z]`` (or ``[B,C,N]``), ``grids [B,N,3]`` -> ``[B,C,3]`` (reference :19-28).

        The kernel rebuilds voxel coordinates from per-axis vectors, so the explicit ``grids``
        tensor is decomposed back into (axis values, zero centre); it must be the separable
        x-major / z-fastest grid ``ProjectLayer`` produces."""
        B, C = int(x.shape[0]), int(x.shape[1])
        if x.dim() != 5:
            raise ValueError("SoftArgmaxLayer expects [B,C,X,Y,Z] cubes")
        X, Y, Z = [int(s) for s in x.shape[2:]]
        x = x.float().contiguous()
        out = torch.empty(B, C, 3, device=x.device, dtype=torch.float32)
        zero = torch.zeros(1, 3, device=x.device, dtype=torch.float32)
        g = grids.float().reshape(B, X, Y, Z, 3)
        for i in range(B):   # per-sample axes (API path only; the batched path never builds `grids`)
            lin = (g[i, :, 0, 0, 0].contiguous(), g[i, 0, :, 0, 1].contiguous(), g[i, 0, 0, :, 2].contiguous())
            out[i] = ops.softargmax_axes(x[i:i + 1], (C * X * Y * Z, X * Y * Z, 1), 1, C, (X, Y, Z), zero, lin,
                                         self.beta)[0]
        return out


class PoseRegressionNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.grid_size = cfg.PICT_STRUCT.GRID_SIZE
        self.cube_size = cfg.PICT_STRUCT.CUBE_SIZE
        self.num_joints = cfg.NETWORK.NUM_JOINTS

        self.project_layer = ProjectLayer(cfg)
        self.v2v_net = V2VNet(cfg.NETWORK.NUM_JOINTS, cfg.NETWORK.NUM_JOINTS)
        self.soft_argmax_layer = SoftArgmaxLayer(cfg)

    def regress(self, all_heatmaps, cams, centers, cube_sample, chunk=None, group_counts=None):
        """Joints of ``n`` person cubes: ``centers [n,>=3]`` (all valid), ``cube_sample [n]`` int32
        sample index of each cube -> ``[n, J, 3]`` world mm.  Cubes are processed ``chunk`` at a
        time to bound activation memory (a 64^3 cube needs ~0.4 GB of float32 activations).
        ``group_counts`` (training mode): the cubes are consecutive groups of that many items, each normalised with
        its own batch statistics (``autograd.grouped_batches``) -- one pass for all proposal slots of a step."""
        n = int(centers.shape[0])
        J = self.num_joints
        if self.training:
            return self._regress_train(all_heatmaps, cams, centers, cube_sample, group_counts)
        if chunk is None:
            import os
            # bf16 activations of one 64^3 cube through V2VNet peak at ~0.2 GB (float32: ~0.4 GB)
            # (float32 activations on the split-operand tensor-core path: 40 cubes = ~20 GB incl. the bf16 term copies)
            # (term-pair activations of the 3-pair tensor-core mode: 2 x bf16 = ~0.4 GB per cube, 80 cubes = ~32 GB)
            default = ("80" if ops.volume_dtype() == torch.bfloat16 or ops.use_split()
                       else ("16" if ops.float32_conv() == "simt" else "40"))
            chunk = int(os.environ.get("SP3D_CUBE_CHUNK", default))
        out = torch.empty(n, J, 3, device=centers.device, dtype=torch.float32)
        X, Y, Z = [int(s) for s in self.cube_size]
        bf16 = ops.volume_dtype() == torch.bfloat16
        hms_f16 = None
        if bf16 and 1 < J <= 16:   # throughput un-projection reads fp16 channel-last maps: convert once per forward
            from .project_layer import _common_strides
            hms, st = _common_strides(all_heatmaps)
            hms_f16 = ops.heatmaps_to_f16(hms, st, J)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            cubes, _ = self.project_layer.project_cl(all_heatmaps, cams, centers[s:e], False, self.grid_size,
                                                     self.cube_size, cube_sample=cube_sample[s:e],
                                                     dtype="split" if ops.use_split() else ops.volume_dtype(),
                                                     c_pitch=ops.round_up(J, 16) if bf16 else None,
                                                     hms_f16=hms_f16)
            if bf16 and J <= 15 and max(X, Y, Z) <= 256:
                # output layer + soft-argmax in one kernel (csrc/conv_tc.cu, fused head)
                head = ops.SoftargmaxHead(e - s, J, (X, Y, Z), centers[s:e], self.grid_size, self.soft_argmax_layer.beta)
                out[s:e] = self.v2v_net.forward_softargmax_cl(cubes, head)
                continue
            y = self.v2v_net.forward_cl(cubes)
            pitch = int(y.shape[-1])
            out[s:e] = ops.softargmax(y, (X * Y * Z * pitch, 1, pitch), e - s, J, (X, Y, Z), centers[s:e],
                                      self.grid_size, self.soft_argmax_layer.beta)
        return out

    def _regress_train(self, all_heatmaps, cams, centers, cube_sample, group_counts=None):
        """``.train()`` mode: the same three steps under autograd (``selfpose3d_b200.autograd``) -- gradients reach the
        heat-maps (and through them the backbone) and the V2VNet parameters, as in the reference's training forward
        (``lib/models/multi_person_posenet_ssv.py:330-407`` calls this module under autograd).  float32 only."""
        from .. import autograd as ag
        if ops.volume_dtype() != torch.float32:
            raise ValueError("the training path runs on float32 volumes (ops.set_volume_dtype(torch.float32))")
        J = self.num_joints
        hms = [h.float().contiguous() for h in all_heatmaps]
        spec = ([float(v) for v in self.grid_size], [int(v) for v in self.cube_size], self.project_layer.img_size,
                self.project_layer.heatmap_size, J, ops.round_up(J, 4))
        cubes = ag.Unproject.apply(cams, centers, cube_sample, spec, *hms)
        if group_counts is not None and sum(group_counts) != int(centers.shape[0]):
            raise ValueError("group_counts must cover the %d cubes" % int(centers.shape[0]))
        with ag.grouped_batches(group_counts if group_counts is not None else [int(centers.shape[0])], cubes.device):
            y = self.v2v_net.forward_cl(cubes)
        return ag.SoftArgmax.apply(y, centers, (J, spec[1], spec[0], float(self.soft_argmax_layer.beta)))

    def regress_slots(self, all_heatmaps, meta, grid_centers, flags, flip_xcoords=None):
        """Training forward of ALL proposal slots in one pass: what the reference gets from
        ``for n in range(num_cand): pose_net(all_heatmaps, meta, grid_centers[:, n])`` (one call per slot with at least one
        valid row, V2VNet on that slot's valid rows: ``lib/models/multi_person_posenet.py:88-99``), with every slot's
        rows as one BatchNorm statistic group.  ``grid_centers [B,K,>=4]``, ``flags``: its ``[..., 3]`` on the host.
        Returns ``(joints [n,J,3], slot [n], sample [n])`` in slot-major order (host lists), or ``None`` when no slot
        is valid."""
        device = all_heatmaps[0].device
        B, K = int(grid_centers.shape[0]), int(grid_centers.shape[1])
        slots, samples, counts = [], [], []
        for n in range(K):
            rows = [i for i in range(B) if float(flags[i, n]) >= 0]
            if rows:
                counts.append(len(rows))
                slots += [n] * len(rows)
                samples += rows
        if not counts:
            return None
        cams = ops.pack_cameras(meta, self.project_layer.img_size, flip_xcoords).to(device, non_blocking=True)
        si = torch.tensor(samples, device=device, dtype=torch.long)
        ni = torch.tensor(slots, device=device, dtype=torch.long)
        centers = grid_centers.to(device=device, dtype=torch.float32)[si, ni].contiguous()
        joints = self.regress(all_heatmaps, cams, centers, si.to(torch.int32), group_counts=counts)
        return joints, slots, samples

    def forward(self, all_heatmaps, meta, grid_centers, flip_xcoords=None):
        device = all_heatmaps[0].device
        B = int(all_heatmaps[0].shape[0])
        pred = torch.zeros(B, self.num_joints, 3, device=device)
        index = torch.nonzero(grid_centers[:, 3] >= 0).flatten()      # host sync, as the reference's boolean mask
        if index.numel() == 0:
            return pred
        cams = ops.pack_cameras(meta, self.project_layer.img_size, flip_xcoords).to(device, non_blocking=True)
        centers = grid_centers.to(device=device, dtype=torch.float32)[index].contiguous()
        pred[index] = self.regress(all_heatmaps, cams, centers, index.to(torch.int32))
        return pred
