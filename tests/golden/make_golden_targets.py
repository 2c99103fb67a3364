#!/usr/bin/env python
"""Golden vectors of the data-side target generation: the UNMODIFIED reference's
``JointsDataset.generate_target_heatmap`` / ``generate_3d_target`` (``/root/reference/lib/dataset/JointsDataset.py``)
on seeded synthetic people -> ``tests/golden/targets.npz``.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_import  # noqa: E402
from integration import shims  # noqa: E402  (json_tricks / vedo stand-ins: import-time dependencies of lib/dataset)

shims.install()
ref_import.install()
from dataset.JointsDataset import JointsDataset  # noqa: E402

IMAGE, HEATMAP, J, P = (288, 384), (72, 96), 15, 6
SPACE_SIZE, SPACE_CENTER, CUBE = [8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0], [80, 80, 20]


def dataset():
    ds = object.__new__(JointsDataset)          # the methods below read these attributes only
    ds.num_joints, ds.target_type, ds.sigma, ds.use_different_joints_weight = J, "gaussian", 3, False
    ds.image_size, ds.heatmap_size = np.array(IMAGE), np.array(HEATMAP)
    ds.space_size, ds.space_center, ds.initial_cube_size = np.array(SPACE_SIZE), np.array(SPACE_CENTER), np.array(CUBE)
    ds.root_id = 2
    return ds


def cases(seed=0, items=6):
    rs = np.random.RandomState(seed)
    joints = np.zeros((items, P, J, 2))
    vis = np.zeros((items, P, J, 3))
    roots = np.zeros((items, P, 3))
    counts = rs.randint(0, P + 1, items)
    counts[0], counts[1] = 0, P
    for i in range(items):
        for p in range(counts[i]):
            c = rs.uniform([-40, -40], [IMAGE[0] + 40, IMAGE[1] + 40])      # some people partly outside the image
            joints[i, p] = c + rs.randn(J, 2) * 30
            v = (rs.rand(J) > 0.2).astype(np.float64)
            if p == 2:
                v[:] = 0                                                    # a person without any visible joint
            vis[i, p] = v[:, None]
            roots[i, p] = [rs.uniform(-4500, 4500), rs.uniform(-5000, 4000), rs.uniform(-400, 2000)]
    joints[1, 0, 0] = [11.999999999999998, 16.0]                       # truncation edge of int(x / 4.0)
    return joints, vis, roots, counts


def main():
    ds = dataset()
    joints, vis, roots, counts = cases()
    targets, weights, volumes = [], [], []
    for i in range(len(counts)):
        n = int(counts[i])
        t, w = ds.generate_target_heatmap([joints[i, p] for p in range(n)], [vis[i, p] for p in range(n)]) if n else (
            np.zeros((J, HEATMAP[1], HEATMAP[0]), np.float32), np.zeros((J, 1), np.float32))
        pose3d = np.zeros((n, J, 3))
        pose3d[:, 2] = roots[i, :n]
        targets.append(t)
        weights.append(w)
        volumes.append(ds.generate_3d_target(pose3d))
    np.savez_compressed(os.path.join(HERE, "targets.npz"), joints=joints, joints_vis=vis, roots=roots, counts=counts,
                        target=np.stack(targets), weight=np.stack(weights), volume=np.stack(volumes),
                        image_size=IMAGE, heatmap_size=HEATMAP, space_size=SPACE_SIZE, space_center=SPACE_CENTER, cube_size=CUBE)
    print("targets.npz %.1f KB" % (os.path.getsize(os.path.join(HERE, "targets.npz")) / 1024))


if __name__ == "__main__":
    main()
