"""Oracle: gradients of the voxel pose path on CPU -- TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference has no hand-written backward: training runs ``torch.autograd`` through the same module calls as
inference (``lib/models/multi_person_posenet_ssv.py:222-501``).  The oracle therefore IS autograd over the oracle's
torch restatements of the forward (float32 or float64), pinned against gradients recorded from the unmodified
reference modules (``tests/golden/make_golden_backward.py`` -> ``tests/golden/backward.npz``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import pipeline


def unproject_grad(heatmaps, cam_arrays, centers, scales, rotations, image_size, heatmap_size, grid_size, grid_center,
                   cube_size, grad_cubes, flip=None):
    """dL/dheatmaps of ``ProjectLayer.forward`` (``lib/models/project_layer.py:42-106``) for ``L = sum(cubes *
    grad_cubes)``: autograd through ``pipeline.unproject_torch``.  ``heatmaps`` list[V] of ``[B,C,h,w]``."""
    hms = [h.detach().clone().requires_grad_(True) for h in heatmaps]
    cubes, _ = pipeline.unproject_torch(hms, cam_arrays, centers, scales, rotations, image_size, heatmap_size,
                                        grid_size, grid_center, cube_size, flip=flip)
    (cubes * grad_cubes).sum().backward()
    return [h.grad if h.grad is not None else torch.zeros_like(h) for h in hms], cubes.detach()


def softargmax_forward(x, grids, beta):
    """``SoftArgmaxLayer.forward`` (``lib/models/pose_regression_net.py:19-28``)."""
    b, c = x.shape[:2]
    p = F.softmax(beta * x.reshape(b, c, -1, 1), dim=2)
    return torch.sum(p * grids.unsqueeze(1), dim=2)


def softargmax_grad(x, grids, beta, grad_out):
    x = x.detach().clone().requires_grad_(True)
    out = softargmax_forward(x, grids, beta)
    (out * grad_out).sum().backward()
    return x.grad, out.detach()


def basic3d_train(x, weight, bias, gamma, beta, grad_y, k, eps=1e-5):
    """``Basic3DBlock`` (``lib/models/v2v_net.py:10-20``) in TRAINING mode (batch statistics) forward + backward for
    ``L = sum(y * grad_y)`` -> ``(y, grad_x, grad_weight, grad_bias, grad_gamma, grad_beta, batch_mean, batch_var)``."""
    x = x.detach().clone().requires_grad_(True)
    ps = [t.detach().clone().requires_grad_(True) for t in (weight, bias, gamma, beta)]
    z = F.conv3d(x, ps[0], ps[1], padding=(k - 1) // 2)
    y = F.relu(F.batch_norm(z, None, None, ps[2], ps[3], training=True, eps=eps))
    (y * grad_y).sum().backward()
    dims = (0, 2, 3, 4)
    return (y.detach(), x.grad) + tuple(p.grad for p in ps) + (z.detach().mean(dims), z.detach().var(dims, unbiased=False))
