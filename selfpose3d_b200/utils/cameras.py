"""Pin-hole + radial/tangential camera model (host/torch form).

Interface mirror of the reference's ``lib/utils/cameras.py`` for the entry
points the hot path touches: ``unfold_camera_param`` (:13-24),
``project_point_radial`` (:27-55) and ``project_pose`` (:111-113).  Inside the
voxel path this arithmetic is fused into the un-projection kernel
(``csrc/unproject.cu``); the torch form here serves callers that project a
handful of points (synthetic data, the SSL re-projection) and is written
point-wise instead of with ``mm``/``einsum``/``ger``.
"""
from __future__ import annotations

import torch


def unfold_camera_param(camera, device=None):
    """Camera dict -> float32 tensors R[3,3] T[3,1] f[2,1] c[2,1] k[3,1] p[2,1]."""
    def f32(v):
        return torch.as_tensor(v, dtype=torch.float, device=device)

    R = f32(camera["R"])
    T = f32(camera["T"]).reshape(3, 1)
    f = torch.stack([f32(camera["fx"]).reshape(()), f32(camera["fy"]).reshape(())]).reshape(2, 1)
    c = torch.stack([f32(camera["cx"]).reshape(()), f32(camera["cy"]).reshape(())]).reshape(2, 1)
    k = f32(camera["k"]).reshape(3, 1)
    p = f32(camera["p"]).reshape(2, 1)
    return R, T, f, c, k, p


def project_point_radial(x, R, T, f, c, k, p):
    """World points ``[N,3]`` (mm) -> distorted pixel coordinates ``[N,2]``.

    ``T`` is the camera centre in world mm; camera frame is ``R (x - T)``.
    No behind-camera rejection (the reference has none): the perspective
    divide uses ``z + 1e-5`` and ``r^2`` is clamped at ``1e10``.
    """
    d = x - T.reshape(1, 3)
    xc = d[:, 0] * R[0, 0] + d[:, 1] * R[0, 1] + d[:, 2] * R[0, 2]
    yc = d[:, 0] * R[1, 0] + d[:, 1] * R[1, 1] + d[:, 2] * R[1, 2]
    zc = d[:, 0] * R[2, 0] + d[:, 1] * R[2, 1] + d[:, 2] * R[2, 2]
    zc = zc + 1e-5
    y0 = xc / zc
    y1 = yc / zc
    r2 = torch.clamp(y0 * y0 + y1 * y1, max=1e10)
    radial = 1 + (k[0] * r2 + k[1] * (r2 * r2) + k[2] * (r2 * r2 * r2))
    tan = p[0] * y1 + p[1] * y0
    corr = radial + 2 * tan
    u = y0 * corr + p[1] * r2
    v = y1 * corr + p[0] * r2
    return torch.stack([f[0] * u + c[0], f[1] * v + c[1]], dim=1)


def project_pose(x, camera):
    R, T, f, c, k, p = unfold_camera_param(camera, device=x.device)
    return project_point_radial(x, R, T, f, c, k, p)
