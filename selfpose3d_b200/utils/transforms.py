"""Host-side 2x3 affine helpers of the un-projection path.

Mirror of the three functions the hot path uses from the reference's
``lib/utils/transforms.py``: ``get_affine_transform`` (:61-103),
``affine_transform_pts_cuda`` (:119-123) and ``get_scale`` (:151-162).  They
prepare the original-image -> network-input affine that the fused
un-projection kernel applies per view; no per-voxel work happens here.

Written without OpenCV: the reference obtains the matrix from
``cv2.getAffineTransform`` on three float32 point pairs; the same 6 unknowns
are eliminated here in float64 in OpenCV's pivot/update order, which makes the
matrix bit-identical to the reference's (checked in tests against goldens).
"""
from __future__ import annotations

import numpy as np
import torch


def _as_np(v):
    if isinstance(v, torch.Tensor):
        return np.array(v.detach().cpu())
    return v


def get_dir(src_point, rot_rad):
    """Rotate ``src_point`` by ``rot_rad`` (reference transforms.py:131-138)."""
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return [src_point[0] * cs - src_point[1] * sn,
            src_point[0] * sn + src_point[1] * cs]


def get_3rd_point(a, b):
    """Third point of the right-angle triple (reference transforms.py:126-128)."""
    direct = a - b
    return np.array(b) + np.array([-direct[1], direct[0]], dtype=np.float32)


def _solve_affine(src, dst):
    """2x3 float64 matrix mapping three float32 ``src`` points to ``dst``.

    Row-interleaved 6x6 system (x-row, y-row per point), partial pivoting,
    ``row_j += (-a_ji / a_ii) * row_i`` updates, then back substitution.
    """
    rows = []
    for (x, y), (u, v) in zip(src.tolist(), dst.tolist()):
        rows.append([x, y, 1.0, 0.0, 0.0, 0.0, u])
        rows.append([0.0, 0.0, 0.0, x, y, 1.0, v])
    n = 6
    for col in range(n):
        best = max(range(col, n), key=lambda r: (abs(rows[r][col]), -r))
        if best != col:
            rows[col], rows[best] = rows[best], rows[col]
        inv = -1.0 / rows[col][col]
        for r in range(col + 1, n):
            f = rows[r][col] * inv
            for c in range(col + 1, n + 1):
                rows[r][c] += f * rows[col][c]
    sol = [0.0] * n
    for r in range(n - 1, -1, -1):
        acc = rows[r][n]
        for c in range(r + 1, n):
            acc -= rows[r][c] * sol[c]
        sol[r] = acc / rows[r][r]
    return np.array(sol, dtype=np.float64).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size,
                         shift=np.array([0, 0], dtype=np.float32), inv=0):
    """Original-image -> network-input affine (reference transforms.py:61-103).

    ``scale`` is in units of 200 px.  The source/destination triples are kept
    in float32 exactly as the reference builds them; the result is float64
    ``[2, 3]``.
    """
    scale = _as_np(scale)
    center = _as_np(center)
    rot = _as_np(rot)
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    scale = np.asarray(scale)

    scale_tmp = scale * 200.0
    src_w, src_h = scale_tmp[0], scale_tmp[1]
    dst_w, dst_h = output_size[0], output_size[1]

    rot_rad = np.pi * rot / 180
    if src_w >= src_h:
        src_dir = get_dir([0, src_w * -0.5], rot_rad)
        dst_dir = np.array([0, dst_w * -0.5], np.float32)
    else:
        src_dir = get_dir([src_h * -0.5, 0], rot_rad)
        dst_dir = np.array([dst_h * -0.5, 0], np.float32)

    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2:, :] = get_3rd_point(src[0, :], src[1, :])
    dst[2:, :] = get_3rd_point(dst[0, :], dst[1, :])

    if inv:
        return _solve_affine(dst, src)
    return _solve_affine(src, dst)


def affine_transform_pts_cuda(pts, t):
    """Apply a 2x3 affine to ``[N, 2]`` points (reference transforms.py:119-123)."""
    x = pts[:, 0]
    y = pts[:, 1]
    ox = t[0, 0] * x + t[0, 1] * y + t[0, 2]
    oy = t[1, 0] * x + t[1, 1] * y + t[1, 2]
    return torch.stack([ox, oy], dim=1)


def get_scale(image_size, resized_size):
    """Letter-box scale in units of 200 px (reference transforms.py:151-162)."""
    w, h = image_size
    w_resized, h_resized = resized_size
    if w / w_resized < h / h_resized:
        w_pad = h / h_resized * w_resized
        h_pad = h
    else:
        w_pad = w
        h_pad = w / w_resized * h_resized
    return np.array([w_pad / 200.0, h_pad / 200.0], dtype=np.float32)
