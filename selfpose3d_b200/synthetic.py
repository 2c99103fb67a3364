"""Seeded synthetic scenes for parity tests, smoke and bench.

No dataset or checkpoint exists in the container, so every measurement runs
on: a ring of Panoptic-like cameras, the collated ``meta`` structure the
reference datasets produce (``lib/dataset/JointsDataset.py:211-223``), Gaussian
heat-maps rendered at the projections of synthetic people, and deterministic
"trained-like" weights.  Only numpy's legacy ``RandomState`` is used, so the
same seed gives the same bytes on every machine.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

from .utils.transforms import get_scale


def ring_cameras(num_views=5, seed=0, image=(1920, 1080), target=(0.0, -500.0, 800.0)):
    """``num_views`` cameras on a ring (radius 4-5 m, height 2.5 m) looking at ``target``.

    World is z-up, millimetres.  Returns dicts with float64 ``R [3,3]``,
    ``T [3,1]`` (camera centre), scalars ``fx fy cx cy``, ``k [3,1]``, ``p [2,1]``.
    """
    rs = np.random.RandomState(seed)
    cams = []
    for v in range(num_views):
        ang = 2 * np.pi * (v + 0.25 * rs.rand()) / num_views
        rad = 4000.0 + 1000.0 * rs.rand()
        centre = np.array([target[0] + rad * np.cos(ang), target[1] + rad * np.sin(ang), 2500.0])
        fwd = np.asarray(target, dtype=np.float64) - centre
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], axis=0)
        cams.append({
            "R": R.astype(np.float64),
            "T": centre.reshape(3, 1).astype(np.float64),
            "fx": np.float64(1400.0 + 20.0 * rs.randn()),
            "fy": np.float64(1400.0 + 20.0 * rs.randn()),
            "cx": np.float64(image[0] / 2 + 5.0 * rs.randn()),
            "cy": np.float64(image[1] / 2 + 5.0 * rs.randn()),
            "k": np.array([[-0.25], [0.12], [-0.01]], dtype=np.float64) * (1 + 0.05 * rs.randn()),
            "p": np.array([[1e-3], [-5e-4]], dtype=np.float64),
        })
    return cams


def make_meta(cams, batch, image_size, orig=(1920, 1080), rotation=None, scale_mul=None,
              dtype=torch.float64):
    """Collated ``meta`` list (one dict per view) as ``default_collate`` would build it.

    ``rotation`` / ``scale_mul``: optional ``[V][B]`` nested sequences for the
    augmentation cases (degrees, multiplicative scale jitter).
    """
    base_scale = get_scale(orig, image_size)
    meta = []
    for v, cam in enumerate(cams):
        rot = torch.zeros(batch, dtype=torch.float64)
        scale = torch.from_numpy(np.tile(base_scale[None], (batch, 1)).astype(np.float32))
        if rotation is not None:
            rot = torch.as_tensor(rotation[v], dtype=torch.float64)
        if scale_mul is not None:
            scale = scale * torch.as_tensor(scale_mul[v], dtype=torch.float32)[:, None]
        camera = {}
        for key, val in cam.items():
            t = torch.as_tensor(np.asarray(val), dtype=dtype)
            camera[key] = t.unsqueeze(0).repeat(batch, *([1] * t.dim())).contiguous()
        meta.append({
            "center": torch.tensor([[orig[0] / 2.0, orig[1] / 2.0]] * batch, dtype=torch.float64),
            "scale": scale,
            "rotation": rot,
            "camera": camera,
        })
    return meta


def synthetic_people(batch, max_people=4, seed=0, num_joints=15,
                     space_center=(0.0, -500.0, 800.0)):
    """``[batch][K][J,3]`` joint positions (mm): roots in the inner 6x6 m, z about 0.9 m."""
    rs = np.random.RandomState(seed + 1000)
    people = []
    for _ in range(batch):
        k = rs.randint(1, max_people + 1)
        sample = []
        for _ in range(k):
            root = np.array([space_center[0] + rs.uniform(-3000, 3000),
                             space_center[1] + rs.uniform(-3000, 3000),
                             900.0 + rs.uniform(-100, 100)])
            sample.append(root[None] + rs.randn(num_joints, 3) * 250.0)
        people.append(sample)
    return people


def render_heatmaps(people, meta, image_size, heatmap_size, num_joints=15, sigma=3.0):
    """Gaussian heat-maps ``list[V] of [B,J,h,w]`` float32 at the projected joints."""
    from .utils.cameras import project_pose
    from .utils.transforms import get_affine_transform, affine_transform_pts_cuda

    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    W, H = float(image_size[0]), float(image_size[1])
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32),
                            torch.arange(w, dtype=torch.float32), indexing="ij")
    out = []
    for m in meta:
        B = m["center"].shape[0]
        hm = torch.zeros(B, num_joints, h, w)
        for i in range(B):
            cam = {k: v[i] for k, v in m["camera"].items()}
            trans = torch.as_tensor(
                get_affine_transform(m["center"][i], m["scale"][i], m["rotation"][i], image_size),
                dtype=torch.float32)
            for person in people[i]:
                pts = torch.as_tensor(person, dtype=torch.float32)
                xy = affine_transform_pts_cuda(project_pose(pts, cam), trans)
                u = xy[:, 0] * w / W
                v = xy[:, 1] * h / H
                g = torch.exp(-((xs[None] - u[:, None, None]) ** 2 + (ys[None] - v[:, None, None]) ** 2)
                              / (2 * sigma * sigma))
                hm[i] = torch.maximum(hm[i], g)
        out.append(hm.contiguous())
    return out


def _rs_for(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 32))


def trained_like_state_dict(module, seed=0, gain=1.0, bn_spread=0.25, out_gain=0.02, res_gamma=0.3):
    """Deterministic non-degenerate weights for any conv/BN module tree.

    Conv / transposed-conv weights are He-normal times ``gain``; biases small;
    BatchNorm affine and running statistics are spread around (1, 0, 0, 1) so
    that folding BN is actually exercised.  Keyed by parameter name, so the
    reference module and ours receive identical tensors.  The last BN of every
    residual branch is scaled by ``res_gamma`` (keeps activations O(1) through
    the un-normalised eval-mode stack) and the output / final 1x1 convs by
    ``out_gain`` so that heat-maps and voxel scores land in a trained-like
    O(0.1) range where soft-argmax (beta = 100) is sharp but not one-hot.
    """
    sd = {}
    for key, ref in module.state_dict().items():
        rs = _rs_for(key, seed)
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = torch.from_numpy(rs.uniform(1 - bn_spread, 1 + bn_spread, shape).astype(np.float32))
        elif key.endswith("running_mean"):
            sd[key] = torch.from_numpy((bn_spread * 0.4 * rs.randn(*shape)).astype(np.float32))
        elif ref.dim() >= 3:  # conv / transposed conv weight
            fan_in = int(np.prod(shape[1:]))
            if "upsample" in key or "deconv" in key:  # [Cin, Cout, k...]: each output sees Cin * k/stride taps
                fan_in = shape[0] * max(1, int(np.prod(shape[2:])) // (2 ** (len(shape) - 2)))
            std = gain * np.sqrt(2.0 / fan_in)
            if "output_layer" in key or "final_layer" in key:
                std *= out_gain
            sd[key] = torch.from_numpy((std * rs.randn(*shape)).astype(np.float32))
        elif key.endswith("weight"):  # BN gamma
            g = rs.uniform(1 - bn_spread, 1 + bn_spread, shape)
            if "bn3." in key or "res_branch.4." in key:
                g = g * res_gamma
            sd[key] = torch.from_numpy(g.astype(np.float32))
        else:  # BN beta / conv bias
            sd[key] = torch.from_numpy((0.05 * rs.randn(*shape)).astype(np.float32))
    return sd


def random_images(batch, num_views, image_size, seed=0):
    """``list[V] of [B,3,H,W]`` float32 N(0,1) images (``image_size`` is ``[w,h]``)."""
    rs = np.random.RandomState(seed + 77)
    W, H = int(image_size[0]), int(image_size[1])
    return [torch.from_numpy(rs.randn(batch, 3, H, W).astype(np.float32)) for _ in range(num_views)]


def ssl_training_case(image_size, heatmap_size, num_joints, batch, num_views, max_people, seed=77, image_seed=40,
                      augment=((12.0, 1.1, False), (-8.0, 0.9, True), (0.0, 1.0, False))):
    """Inputs of one self-supervised training step (``MultiPersonPoseNetSSV.forward(inference=False)``): three view
    sets ``(views, meta, pseudo heat-maps)`` -- set 1 and 2 with rotation / scale jitter (and an h-flip), set 3 plain --
    with the ``meta`` entries the SSL forward reads (``camera`` incl. ``f`` / ``c``, ``trans``, ``hflip``, pseudo 2-D
    poses ``joints`` / ``joints_vis``), as ``lib/dataset/JointsDatasetSSV.py:540-640`` collates them."""
    from .utils.transforms import get_affine_transform
    cams = ring_cameras(num_views, seed=0)
    rs = np.random.RandomState(seed)
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    sets = []
    for s, (rot, mul, flip) in enumerate(augment):
        meta = make_meta(cams, batch, image_size, rotation=[[rot] * batch] * num_views, scale_mul=[[mul] * batch] * num_views)
        for m in meta:
            cam = {k: v.float() for k, v in m["camera"].items()}
            cam["f"] = torch.stack([cam["fx"], cam["fy"]], -1).reshape(batch, 2, 1)
            cam["c"] = torch.stack([cam["cx"], cam["cy"]], -1).reshape(batch, 2, 1)
            m["camera"] = cam
            m["joints"] = torch.zeros(batch, max_people, num_joints, 2, dtype=torch.float64)
            m["joints"][:, :2] = torch.from_numpy(rs.uniform(5, 60, (batch, 2, num_joints, 2)))
            m["joints_vis"] = torch.ones(batch, max_people, num_joints, 2, dtype=torch.float64)
        trans = np.stack([get_affine_transform(meta[0]["center"][b].numpy(), meta[0]["scale"][b].numpy(),
                                               float(meta[0]["rotation"][b]), image_size) for b in range(batch)])
        meta[0]["trans"] = torch.from_numpy(trans.astype(np.float32))
        meta[0]["hflip"] = torch.tensor([flip] * batch)
        views = random_images(batch, num_views, image_size, seed=image_seed + s)
        targets = [torch.from_numpy(rs.rand(batch, num_joints, h, w).astype(np.float32)) for _ in range(num_views)]
        sets.append((views, meta, targets))
    return sets
