"""Training targets rendered on the GPU.

The reference's DataLoader workers build, per item and with numpy, the 2-D Gaussian target heat-maps of every view
(``JointsDataset.generate_target_heatmap``, ``lib/dataset/JointsDataset.py:237-302``) and the 3-D root target volume
(``generate_3d_target``, ``:304-341``); once the GPU path is fast that CPU work bounds the input pipeline.  These two
functions take the collated ``joints`` / ``roots`` of a whole batch and render the same targets in one launch each
(``sp3d_target_heatmaps`` / ``sp3d_target_volume``)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .ops import _require_cuda, _stream

_WINDOWS = {}


def gaussian_window(sigma):
    """The reference's own Gaussian window (``:275-281``), evaluated with numpy in float32 exactly as it does."""
    key = float(sigma)
    if key not in _WINDOWS:
        tmp_size = sigma * 3
        size = 2 * tmp_size + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, np.newaxis]
        x0 = y0 = size // 2
        _WINDOWS[key] = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2)).astype(np.float32)
    return _WINDOWS[key]


def heatmap_targets(joints, joints_vis, n_people, image_size, heatmap_size, sigma=3):
    """``joints [n, P, J, >=2]`` float64 network-input pixels, ``joints_vis [n, P, J, >=1]``, ``n_people [n]`` (CUDA)
    -> ``(target [n, J, h, w], target_weight [n, J, 1])`` float32, the values of ``generate_target_heatmap`` for every
    item (an item = one view of one sample).  ``image_size`` / ``heatmap_size`` are ``[w, h]``."""
    _require_cuda(joints, joints_vis, n_people)
    joints, joints_vis = joints.double().contiguous(), joints_vis.double().contiguous()
    n, P, J = [int(v) for v in joints.shape[:3]]
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    win = gaussian_window(sigma)
    window = torch.from_numpy(win).to(joints.device)
    target = torch.empty(n, J, h, w, device=joints.device, dtype=torch.float32)
    weight = torch.empty(n, J, device=joints.device, dtype=torch.float32)
    counts = n_people.to(device=joints.device, dtype=torch.int32).contiguous()
    a = _lib.TargetHeatmapsArgs()
    a.joints, a.joints_vis, a.n_people = joints.data_ptr(), joints_vis.data_ptr(), counts.data_ptr()
    a.n_items, a.P, a.J, a.jstride, a.vstride = n, P, J, int(joints.shape[3]), int(joints_vis.shape[3])
    a.h, a.w = h, w
    stride = np.asarray(image_size, dtype=np.float64) / np.asarray(heatmap_size, dtype=np.float64)
    a.stride_x, a.stride_y = float(stride[0]), float(stride[1])
    a.window, a.radius = window.data_ptr(), int(win.shape[0] // 2)
    a.target, a.target_weight = target.data_ptr(), weight.data_ptr()
    _lib.call("sp3d_target_heatmaps", a, _stream(), kind="targets", work=target.numel() * 4)
    return target, weight[..., None]


def root_targets(roots, n_people, space_size, space_center, cube_size, sigma=200.0):
    """``roots [n, P, 3]`` float64 world mm, ``n_people [n]`` (CUDA) -> ``[n, X, Y, Z]`` float32, the values of
    ``generate_3d_target`` (voxel-wise maximum of the people's 200 mm Gaussians inside their 3-sigma boxes)."""
    _require_cuda(roots, n_people)
    roots = roots.double().contiguous()
    n, P = int(roots.shape[0]), int(roots.shape[1])
    X, Y, Z = [int(v) for v in cube_size]
    dev = roots.device
    grids = [torch.from_numpy(np.linspace(-space_size[a] / 2, space_size[a] / 2, int(cube_size[a])) + space_center[a]).to(dev)
             for a in range(3)]
    target = torch.empty(n, X, Y, Z, device=dev, dtype=torch.float32)
    counts = n_people.to(device=dev, dtype=torch.int32).contiguous()
    a = _lib.TargetVolumeArgs()
    a.roots, a.n_people, a.n_items, a.P = roots.data_ptr(), counts.data_ptr(), n, P
    a.grid_x, a.grid_y, a.grid_z = grids[0].data_ptr(), grids[1].data_ptr(), grids[2].data_ptr()
    a.X, a.Y, a.Z, a.sigma = X, Y, Z, float(sigma)
    a.target = target.data_ptr()
    _lib.call("sp3d_target_volume", a, _stream(), kind="targets", work=target.numel() * 4)
    return target
