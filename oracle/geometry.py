"""Oracle: camera projection, input affine and the multi-view un-projection (numpy).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Restates ``lib/utils/transforms.py:61-103`` (``get_affine_transform``),
``lib/utils/cameras.py:27-55`` (``project_point_radial``) and
``lib/models/project_layer.py:22-102`` (``compute_grid`` / ``get_voxel``) with
every intermediate rounded to ``dtype`` in the reference's operation order.
``dtype=np.float64`` gives the up-cast "truth" used to budget tolerances.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- affine
def get_affine_transform(center, scale, rot, output_size):
    """lib/utils/transforms.py:61-103 with ``shift = 0``, ``inv = 0``.

    The reference fills float32 source/destination triples and hands them to
    ``cv2.getAffineTransform`` (float64 solve of the 6 unknowns, restated in
    ``_cv_get_affine``).
    """
    center = np.asarray(center, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float32)
    scale_tmp = scale * np.float32(200.0)                    # float32, :74
    src_w, src_h = scale_tmp[0], scale_tmp[1]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * float(rot) / 180                       # :78
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    if src_w >= src_h:                                       # :79-84
        p = (0.0, float(src_w) * -0.5)
        dst_dir = np.array([0, dst_w * -0.5], np.float32)
    else:
        p = (float(src_h) * -0.5, 0.0)
        dst_dir = np.array([dst_h * -0.5, 0], np.float32)
    src_dir = np.array([p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs])   # get_dir :131-138

    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center                                          # :88
    src[1] = center + src_dir                                # :89
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    for arr in (src, dst):                                   # get_3rd_point :126-128
        d = arr[0] - arr[1]
        arr[2] = arr[1] + np.array([-d[1], d[0]], dtype=np.float32)

    return _cv_get_affine(src, dst)


def _cv_get_affine(src, dst):
    """``cv2.getAffineTransform`` (OpenCV 4.x ``imgwarp.cpp``): the 6x6 system
    ``[x y 1 0 0 0; 0 0 0 x y 1] m = [u; v]`` solved in float64 by OpenCV's own
    partial-pivot Gaussian elimination (``matrix_decomp.cpp`` ``LUImpl``), whose
    operation order is followed here so the result is bit-identical."""
    n = 6
    A = [[0.0] * n for _ in range(n)]
    b = [0.0] * n
    for i in range(3):
        A[2 * i][0] = A[2 * i + 1][3] = float(src[i][0])
        A[2 * i][1] = A[2 * i + 1][4] = float(src[i][1])
        A[2 * i][2] = A[2 * i + 1][5] = 1.0
        b[2 * i], b[2 * i + 1] = float(dst[i][0]), float(dst[i][1])
    for i in range(n):
        piv = i
        for j in range(i + 1, n):
            if abs(A[j][i]) > abs(A[piv][i]):
                piv = j
        if piv != i:
            A[i], A[piv] = A[piv], A[i]
            b[i], b[piv] = b[piv], b[i]
        d = -1.0 / A[i][i]
        for j in range(i + 1, n):
            alpha = A[j][i] * d
            for c in range(i + 1, n):
                A[j][c] += alpha * A[i][c]
            b[j] += alpha * b[i]
    for i in range(n - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, n):
            s -= A[i][c] * b[c]
        b[i] = s / A[i][i]
    return np.array(b, dtype=np.float64).reshape(2, 3)


def get_scale(image_size, resized_size):
    """lib/utils/transforms.py:151-162."""
    w, h = image_size
    wr, hr = resized_size
    if w / wr < h / hr:
        w_pad, h_pad = h / hr * wr, h
    else:
        w_pad, h_pad = w, w / wr * hr
    return np.array([w_pad / 200.0, h_pad / 200.0], dtype=np.float32)


# ----------------------------------------------------------------------------- camera
def project_point_radial(x, R, T, f, c, k, p, dtype=np.float32):
    """lib/utils/cameras.py:27-55.  ``x [N,3]`` -> ``[N,2]`` pixels, all in ``dtype``."""
    t = dtype
    x = np.asarray(x, dtype=t)
    R = np.asarray(R, dtype=t).reshape(3, 3)
    T = np.asarray(T, dtype=t).reshape(3)
    f = np.asarray(f, dtype=t).reshape(2)
    c = np.asarray(c, dtype=t).reshape(2)
    k = np.asarray(k, dtype=t).reshape(3)
    p = np.asarray(p, dtype=t).reshape(2)
    d = (x - T[None]).astype(t)                                          # :41 x^T - T
    xc = ((d[:, 0] * R[0, 0] + d[:, 1] * R[0, 1]).astype(t) + d[:, 2] * R[0, 2]).astype(t)
    yc = ((d[:, 0] * R[1, 0] + d[:, 1] * R[1, 1]).astype(t) + d[:, 2] * R[1, 2]).astype(t)
    zc = ((d[:, 0] * R[2, 0] + d[:, 1] * R[2, 1]).astype(t) + d[:, 2] * R[2, 2]).astype(t)
    zc = (zc + t(1e-5)).astype(t)                                        # :42
    y0 = (xc / zc).astype(t)
    y1 = (yc / zc).astype(t)
    r2 = np.minimum((y0 * y0 + y1 * y1).astype(t), t(1e10))              # :45-46
    r4 = (r2 * r2).astype(t)
    r6 = (r4 * r2).astype(t)
    radial = (t(1) + ((k[0] * r2 + k[1] * r4).astype(t) + k[2] * r6).astype(t)).astype(t)   # :47-48
    tan = (p[0] * y1 + p[1] * y0).astype(t)                              # :50
    corr = (radial + (t(2) * tan).astype(t)).astype(t)                   # :51
    u = (y0 * corr + p[1] * r2).astype(t)                                # :53
    v = (y1 * corr + p[0] * r2).astype(t)
    return np.stack([(f[0] * u + c[0]).astype(t), (f[1] * v + c[1]).astype(t)], axis=1)     # :54


# ----------------------------------------------------------------------------- grid
def compute_grid_1d(box_size, box_center, n_bins, linspace=None):
    """The three 1-D coordinate vectors of ``compute_grid`` (project_layer.py:28-35).

    ``torch.linspace`` on CPU is not reproducible by a closed form to the last
    ulp (SURVEY.md §7 hard part 4), so the float32 vectors are taken from
    ``torch.linspace`` itself unless ``linspace`` (a callable) is supplied.
    """
    if linspace is None:
        import torch

        def linspace(a, b, n):
            return torch.linspace(a, b, int(n)).numpy()
    out = []
    for a in range(3):
        g = linspace(-box_size[a] / 2, box_size[a] / 2, n_bins[a]).astype(np.float32)
        out.append((g + np.float32(box_center[a])).astype(np.float32))
    return out


def compute_grid(box_size, box_center, n_bins):
    """project_layer.py:22-40: ``[X*Y*Z, 3]`` float32, x slowest, z fastest."""
    gx, gy, gz = compute_grid_1d(box_size, box_center, n_bins)
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    return np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)


# ----------------------------------------------------------------------------- sampling
def bilinear_sample_zeros(img, fx, fy, dtype=np.float32):
    """``F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)``
    on un-normalised coordinates.  ``img [C,h,w]``, ``fx, fy [N]`` -> ``[C,N]``."""
    t = dtype
    C, h, w = img.shape
    x0 = np.floor(fx)
    y0 = np.floor(fy)
    wx1 = (fx - x0).astype(t)
    wy1 = (fy - y0).astype(t)
    wx0 = (t(1) - wx1).astype(t)
    wy0 = (t(1) - wy1).astype(t)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    out = np.zeros((C, fx.shape[0]), dtype=t)
    for (dy, dx, wgt) in ((0, 0, wx0 * wy0), (0, 1, wx1 * wy0), (1, 0, wx0 * wy1), (1, 1, wx1 * wy1)):
        xi = x0 + dx
        yi = y0 + dy
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        vals = img[:, np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)].astype(t)
        out = (out + vals * (wgt.astype(t) * ok)[None]).astype(t)
    return out


def view_sample_coords(grid, cam, center, trans, image_size, heatmap_size, flip, dtype=np.float32,
                       tensor_wh=None):
    """Per-voxel heat-map coordinates of one view (project_layer.py:76-90).

    Returns ``(fx, fy, mask, px, py)``: un-normalised sampling coordinates, the
    in-image mask on the un-clamped pixel coordinates (:78-79) and those pixel
    coordinates (for boundary-ambiguity analysis in tests).
    """
    t = dtype
    W, H = t(image_size[0]), t(image_size[1])
    w, h = t(heatmap_size[0]), t(heatmap_size[1])
    f = [cam["fx"], cam["fy"]]
    c = [cam["cx"], cam["cy"]]
    xy = project_point_radial(grid, cam["R"], cam["T"], f, c, cam["k"], cam["p"], dtype=t)
    px, py = xy[:, 0], xy[:, 1]
    width, height = t(2 * center[0]), t(2 * center[1])                    # :68
    mask = (px >= 0) & (py >= 0) & (px < width) & (py < height)           # :78-79
    hi = max(width, height)
    cx_ = np.clip(px, t(-1.0), hi)                                        # :80
    cy_ = np.clip(py, t(-1.0), hi)
    A = np.asarray(trans, dtype=t)                                        # :69-72 float32 affine
    qx = ((A[0, 0] * cx_ + A[0, 1] * cy_).astype(t) + A[0, 2]).astype(t)  # :81
    qy = ((A[1, 0] * cx_ + A[1, 1] * cy_).astype(t) + A[1, 2]).astype(t)
    if flip:
        qx = (W - qx).astype(t)                                           # :82-83
    u = ((qx * w).astype(t) / W).astype(t)                                # :84-86
    v = ((qy * h).astype(t) / H).astype(t)
    sx = np.clip((((u / (w - t(1))).astype(t) * t(2)).astype(t) - t(1)).astype(t), t(-1.1), t(1.1))   # :87-90
    sy = np.clip((((v / (h - t(1))).astype(t) * t(2)).astype(t) - t(1)).astype(t), t(-1.1), t(1.1))
    # align_corners=True un-normalise: grid_sample uses the TENSOR extent (== heatmap_size unless the
    # network input is not a multiple of 32, see tests/golden "inference_images")
    tw, th = (w, h) if tensor_wh is None else (t(tensor_wh[0]), t(tensor_wh[1]))
    fx = (((sx + t(1)) / t(2)).astype(t) * (tw - t(1))).astype(t)
    fy = (((sy + t(1)) / t(2)).astype(t) * (th - t(1))).astype(t)
    return fx, fy, mask, px, py


def unproject(heatmaps, cams, centers, scales, rotations, image_size, heatmap_size,
              grid_size, grid_center, cube_size, flip=None, dtype=np.float32, return_aux=False):
    """The un-projection ``ProjectLayer.get_voxel`` (project_layer.py:42-102).

    heatmaps  ``[V,B,C,h,w]``; cams ``[V][B]`` dicts (R,T,fx,fy,cx,cy,k,p);
    centers ``[V,B,2]``, scales ``[V,B,2]``, rotations ``[V,B]``;
    grid_center ``[1,3]`` (shared) or ``[B,5]`` (rows with ``[3] < 0`` skipped).
    Returns ``cubes [B,C,X,Y,Z]``, ``grids [B,N,3]`` (float32) and, with
    ``return_aux``, the minimum distance of any projected pixel to a mask
    boundary per voxel (``[B,N]``, for tests that exclude ambiguous voxels).
    """
    t = dtype
    heatmaps = np.asarray(heatmaps)
    V, B, C, h, w = heatmaps.shape
    X, Y, Z = [int(s) for s in cube_size]
    N = X * Y * Z
    grid_center = np.asarray(grid_center, dtype=np.float32)
    cubes = np.zeros((B, C, N), dtype=t)
    grids = np.zeros((B, N, 3), dtype=np.float32)
    margin = np.full((B, N), np.inf)
    for i in range(B):
        if grid_center.shape[1] != 3 and not grid_center[i, 3] >= 0:       # :54
            continue
        gc = grid_center[0] if grid_center.shape[0] == 1 else grid_center[i]
        grid = compute_grid(grid_size, gc[:3], (X, Y, Z))
        grids[i] = grid
        num = np.zeros((C, N), dtype=t)
        den = np.zeros(N, dtype=t)
        for c in range(V):
            trans = get_affine_transform(centers[c][i], scales[c][i], rotations[c][i], image_size)
            fl = bool(flip[i]) if flip is not None else False
            fx, fy, mask, px, py = view_sample_coords(
                grid.astype(t), cams[c][i], centers[c][i], trans.astype(np.float32), image_size,
                heatmap_size, fl, dtype=t, tensor_wh=(w, h))
            s = bilinear_sample_zeros(heatmaps[c, i], fx, fy, dtype=t)
            m = mask.astype(t)
            num = (num + s * m[None]).astype(t)                            # :93,96
            den = (den + m).astype(t)
            width, height = 2 * float(centers[c][i][0]), 2 * float(centers[c][i][1])
            dist = np.minimum(np.minimum(np.abs(px), np.abs(py)),
                              np.minimum(np.abs(px - width), np.abs(py - height)))
            margin[i] = np.minimum(margin[i], dist)
        out = (num / (den + t(1e-6))[None]).astype(t)                      # :96
        out[out != out] = 0                                                # :98
        cubes[i] = np.clip(out, 0, 1)                                      # :99
    cubes = cubes.reshape(B, C, X, Y, Z)
    if return_aux:
        return cubes, grids, margin
    return cubes, grids
