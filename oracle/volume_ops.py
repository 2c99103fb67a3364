"""Oracle: 3-D NMS / top-k proposals and soft-argmax (numpy).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Restates ``lib/core/proposal.py:18-48`` (``get_index`` / ``max_pool`` / ``nms``),
``lib/models/cuboid_proposal_net_soft.py:46-68`` (``ProposalLayerSoft``;
``cuboid_proposal_net.py:42-83`` is the same arithmetic on the no-GT branch) and
``lib/models/pose_regression_net.py:19-28`` (``SoftArgmaxLayer``).
"""
from __future__ import annotations

import numpy as np


def max_pool3(x):
    """``F.max_pool3d(k=3, s=1, p=1)`` on ``[B,X,Y,Z]`` (-inf padding) -- proposal.py:28-30."""
    B, X, Y, Z = x.shape
    p = np.full((B, X + 2, Y + 2, Z + 2), -np.inf, dtype=x.dtype)
    p[:, 1:-1, 1:-1, 1:-1] = x
    out = np.full_like(x, -np.inf)
    for dx in range(3):
        for dy in range(3):
            for dz in range(3):
                out = np.maximum(out, p[:, dx:dx + X, dy:dy + Y, dz:dz + Z])
    return out


def nms(root_cubes, max_num):
    """proposal.py:35-48.  Non-maxima are zeroed (not -inf), then top-k over the
    flattened volume.  Tie policy (implementation-defined in torch): highest
    value first, then lowest flat index.  Returns ``(values [B,K], index [B,K,3])``."""
    x = np.asarray(root_cubes, dtype=np.float32)
    B = x.shape[0]
    keep = (x == max_pool3(x)).astype(np.float32)
    flat = (keep * x).reshape(B, -1)
    order = np.lexsort((np.broadcast_to(np.arange(flat.shape[1]), flat.shape), -flat), axis=1)[:, :max_num]
    vals = np.take_along_axis(flat, order, axis=1)
    shape = x.shape[1:]
    ix = order // (shape[1] * shape[2])                       # get_index :18-25
    iy = (order % (shape[1] * shape[2])) // shape[2]
    iz = order % shape[2]
    return vals, np.stack([ix, iy, iz], axis=2)


def get_real_loc(index, space_size, space_center, cube_size, f64=False):
    """cuboid_proposal_net_soft.py:46-52: ``idx / (n-1) * size + centre - size / 2`` in that
    operation order (not bit-equal to the ``linspace`` voxel coordinate).

    ``idx / (n-1)`` is always float32.  ``size`` / ``centre`` are ``torch.tensor(cfg...)``
    (:21-23): float32 when the config holds python lists (every YAML-loaded config), float64
    when it holds the numpy defaults of ``config.py:224-227`` -- then torch promotes the rest
    of the expression to float64 (``f64=True``) and the float32 cast happens on assignment
    into ``grid_centers`` (:60)."""
    f = np.float32
    q = (index.astype(f) / (np.asarray(cube_size, dtype=f) - f(1)).astype(f)).astype(f)
    t = np.float64 if f64 else np.float32
    size = np.asarray(space_size, dtype=t)
    cen = np.asarray(space_center, dtype=t)
    loc = (q.astype(t) * size).astype(t)
    loc = (loc + cen).astype(t)
    return (loc - (size / t(2.0)).astype(t)).astype(t).astype(f)


def proposal_layer(root_cubes, space_size, space_center, cube_size, max_people, threshold, f64=False):
    """cuboid_proposal_net_soft.py:54-68 -> ``grid_centers [B,K,5]`` = (x, y, z, flag, score),
    ``flag = (score > threshold) - 1``."""
    vals, idx = nms(root_cubes, max_people)
    B = vals.shape[0]
    gc = np.zeros((B, max_people, 5), dtype=np.float32)
    gc[:, :, 0:3] = get_real_loc(idx, space_size, space_center, cube_size, f64=f64)
    gc[:, :, 4] = vals
    gc[:, :, 3] = (vals > np.float32(threshold)).astype(np.float32) - 1.0
    return gc


def soft_argmax(x, grids, beta, dtype=np.float32):
    """pose_regression_net.py:19-28: ``sum_v softmax(beta * x)_v * grid_v``.

    ``x [B,C,...voxels]``, ``grids [B,N,3]`` -> ``[B,C,3]``.  With
    ``dtype=float64`` this is the up-cast truth for tolerance budgeting."""
    x = np.asarray(x)
    B, C = x.shape[:2]
    z = (x.reshape(B, C, -1).astype(dtype) * dtype(beta)).astype(dtype)
    z = z - z.max(axis=2, keepdims=True)
    e = np.exp(z).astype(dtype)
    p = (e / e.sum(axis=2, keepdims=True, dtype=dtype)).astype(dtype)
    g = np.asarray(grids).astype(dtype)
    return np.einsum("bcn,bnk->bck", p, g).astype(dtype)
