"""Oracle: the whole inference path (images/heat-maps -> proposals -> 3-D poses) on CPU.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Also the timed CPU baseline
("port") of ``bench.py``.

Restates ``MultiPersonPoseNetSSV.do_inference``
(``lib/models/multi_person_posenet_ssv.py:105-153``),
``CuboidProposalNetSoft.get_grid_centres`` (``cuboid_proposal_net_soft.py:129-149``)
and ``PoseRegressionNet.forward`` (``pose_regression_net.py:41-53``) over the
oracle pieces.  ``unproject_torch`` is the same un-projection as
``oracle.geometry.unproject`` written with the multi-threaded torch CPU calls the
reference itself makes (``mm``, ``F.grid_sample``), so that the CPU baseline is
timed the way the reference would run.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import geometry, nets, volume_ops


def _cam_at(cam_arrays, c, i):
    return {k: np.asarray(v[c][i]) for k, v in cam_arrays.items()}


def unproject_torch(heatmaps, cam_arrays, centers, scales, rotations, image_size, heatmap_size,
                    grid_size, grid_center, cube_size, flip=None, dtype=torch.float32):
    """project_layer.py:42-102 with torch CPU ops.  ``heatmaps`` list[V] of ``[B,C,h,w]``.  ``dtype`` float64: the same
    operations in double precision (the "truth" used to budget tolerances; the reference itself is float32)."""
    V = len(heatmaps)
    B, C = heatmaps[0].shape[:2]
    X, Y, Z = [int(s) for s in cube_size]
    N = X * Y * Z
    w, h = float(heatmap_size[0]), float(heatmap_size[1])
    W, H = float(image_size[0]), float(image_size[1])
    gc_all = torch.as_tensor(np.asarray(grid_center), dtype=torch.float32).to(dtype)
    cubes = torch.zeros(B, C, N, dtype=dtype)
    grids = torch.zeros(B, N, 3, dtype=dtype)
    heatmaps = [hm.to(dtype) for hm in heatmaps]
    for i in range(B):
        if gc_all.shape[1] != 3 and not gc_all[i, 3] >= 0:
            continue
        gc = gc_all[0] if gc_all.shape[0] == 1 else gc_all[i]
        g1 = [torch.linspace(-grid_size[a] / 2, grid_size[a] / 2, cube_size[a]).to(dtype) + gc[a] for a in range(3)]
        gx, gy, gz = torch.meshgrid(g1[0], g1[1], g1[2], indexing="ij")
        grid = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)
        grids[i] = grid
        num = torch.zeros(C, N, dtype=dtype)
        den = torch.zeros(N, dtype=dtype)
        for c in range(V):
            cam = _cam_at(cam_arrays, c, i)
            # (camera parameters are cast to float32 first, as the reference does: cameras.py:14-23)
            R = torch.as_tensor(cam["R"], dtype=torch.float32).to(dtype)
            T = torch.as_tensor(cam["T"], dtype=torch.float32).reshape(3, 1).to(dtype)
            k = torch.as_tensor(cam["k"], dtype=torch.float32).reshape(3).to(dtype)
            p = torch.as_tensor(cam["p"], dtype=torch.float32).reshape(2).to(dtype)
            xcam = torch.mm(R, grid.t() - T)
            y = xcam[:2] / (xcam[2] + 1e-5)
            r2 = torch.clamp((y ** 2).sum(0), max=1e10)
            radial = 1 + k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3
            tan = p[0] * y[1] + p[1] * y[0]
            corr = radial + 2 * tan
            u = y[0] * corr + p[1] * r2
            v = y[1] * corr + p[0] * r2
            px = float(cam["fx"]) * u + float(cam["cx"])
            py = float(cam["fy"]) * v + float(cam["cy"])
            width, height = 2 * float(centers[c][i][0]), 2 * float(centers[c][i][1])
            m = ((px >= 0) & (py >= 0) & (px < width) & (py < height)).to(dtype)
            hi = max(width, height)
            px = px.clamp(-1.0, hi)
            py = py.clamp(-1.0, hi)
            A = torch.as_tensor(geometry.get_affine_transform(
                centers[c][i], scales[c][i], rotations[c][i], image_size), dtype=torch.float32).to(dtype)
            qx = A[0, 0] * px + A[0, 1] * py + A[0, 2]
            qy = A[1, 0] * px + A[1, 1] * py + A[1, 2]
            if flip is not None and bool(flip[i]):
                qx = W - qx
            sx = (qx * w / W / (w - 1) * 2.0 - 1.0).clamp(-1.1, 1.1)
            sy = (qy * h / H / (h - 1) * 2.0 - 1.0).clamp(-1.1, 1.1)
            sg = torch.stack([sx, sy], dim=1).view(1, 1, N, 2)
            s = F.grid_sample(heatmaps[c][i:i + 1], sg, align_corners=True)[0, :, 0]
            num += s * m[None]
            den += m
        out = num / (den + 1e-6)[None]
        out[out != out] = 0.0
        cubes[i] = out.clamp(0.0, 1.0)
    return cubes.view(B, C, X, Y, Z), grids


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def inference(sd, cfg, cam_arrays, centers, scales, rotations, images=None, heatmaps=None,
              root_channel_only=True, dtype=torch.float32, grid_centers=None):
    """do_inference (multi_person_posenet_ssv.py:105-153).

    ``sd``: full model state dict (``backbone.*``, ``root_net.v2v_net.*``,
    ``pose_net.v2v_net.*``).  ``cfg``: dict with ``image_size``, ``heatmap_size``,
    ``space_size``, ``space_center``, ``initial_cube_size``, ``grid_size``,
    ``cube_size``, ``max_people``, ``threshold``, ``beta``, ``root_idx``.
    Returns ``pred [B,K,J,5]``, heat-maps ``list[V]``, ``grid_centers [B,K,5]``,
    ``root_cubes [B,X,Y,Z]`` (torch tensors).

    ``dtype`` float64 evaluates every step in double precision (tolerance budgeting: how far the float32 reference
    itself is from exact arithmetic).  ``grid_centers``: take these proposals instead of the root net's own (so that
    a float64 run regresses the SAME person cubes as a float32 run whose near-tied top-K order may differ).
    """
    if heatmaps is None:
        bsd = _sub(sd, "backbone.")
        heatmaps = [nets.pose_resnet_forward(v.to(dtype), bsd, dtype=dtype) for v in images]          # :108-110
    B, J = heatmaps[0].shape[:2]
    K = int(cfg["max_people"])
    hm_root = [h[:, cfg["root_idx"]][:, None].contiguous() for h in heatmaps] if root_channel_only else heatmaps
    init_cubes, _ = unproject_torch(hm_root, cam_arrays, centers, scales, rotations, cfg["image_size"],
                                    cfg["heatmap_size"], cfg["space_size"], [cfg["space_center"]],
                                    cfg["initial_cube_size"], dtype=dtype)     # cuboid_proposal_net_soft.py:137-144
    root_cubes = nets.v2v_forward(init_cubes, _sub(sd, "root_net.v2v_net."), dtype=dtype)[:, 0]
    if grid_centers is None:
        gc = torch.from_numpy(volume_ops.proposal_layer(
            root_cubes.float().numpy(), cfg["space_size"], cfg["space_center"], cfg["initial_cube_size"], K,
            cfg["threshold"]))
    else:
        gc = torch.as_tensor(grid_centers).float().clone()
    pred = torch.zeros(B, K, J, 5, dtype=dtype)
    pred[:, :, :, 3:] = gc[:, :, 3:].reshape(B, -1, 1, 2).to(dtype)            # :139-140
    psd = _sub(sd, "pose_net.v2v_net.")
    for n in range(K):                                                         # :143-148
        index = gc[:, n, 3] >= 0
        if int(index.sum()) == 0:
            continue
        cubes, grids = unproject_torch(heatmaps, cam_arrays, centers, scales, rotations, cfg["image_size"],
                                       cfg["heatmap_size"], cfg["grid_size"], gc[:, n].numpy(),
                                       cfg["cube_size"], dtype=dtype)
        valid = nets.v2v_forward(cubes[index], psd, dtype=dtype)               # pose_regression_net.py:49-51
        p = torch.softmax(float(cfg["beta"]) * valid.reshape(valid.shape[0], J, -1, 1), dim=2)
        pred[index, n, :, 0:3] = (p * grids[index].unsqueeze(1)).sum(dim=2)
    return pred, heatmaps, gc, root_cubes
