"""Oracle: the reference's data-side target generation, restated in numpy.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows ``JointsDataset.compute_human_scale`` /
``generate_target_heatmap`` / ``generate_3d_target`` (``lib/dataset/JointsDataset.py:227-341``) line by line; pinned by
``tests/golden/targets.npz`` (generated from the unmodified reference by ``tests/golden/make_golden_targets.py``)."""
import numpy as np


def target_heatmap(joints, joints_vis, image_size, heatmap_size, sigma=3):
    """``joints``: list over people of ``[J, >=2]``, ``joints_vis`` likewise -> ``(target [J,h,w], weight [J,1])`` (:237-302)."""
    image_size, heatmap_size = np.asarray(image_size), np.asarray(heatmap_size)
    J = joints[0].shape[0] if len(joints) else joints_vis[0].shape[0]
    weight = np.zeros((J, 1), dtype=np.float32)
    for i in range(J):                                                      # :245-249
        for n in range(len(joints)):
            if joints_vis[n][i, 0] == 1:
                weight[i, 0] = 1
    target = np.zeros((J, heatmap_size[1], heatmap_size[0]), dtype=np.float32)
    feat_stride = image_size / heatmap_size                                 # :258
    tmp_size = sigma * 3
    for n in range(len(joints)):
        if np.sum(joints_vis[n][:, 0] == 1) == 0:                           # compute_human_scale == 0 (:228-230,261)
            continue
        for j in range(J):
            mu_x = int(joints[n][j][0] / feat_stride[0])                    # :268-269
            mu_y = int(joints[n][j][1] / feat_stride[1])
            ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
            br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
            if joints_vis[n][j, 0] == 0 or ul[0] >= heatmap_size[0] or ul[1] >= heatmap_size[1] or br[0] < 0 or br[1] < 0:
                continue
            size = 2 * tmp_size + 1
            x = np.arange(0, size, 1, np.float32)
            y = x[:, np.newaxis]
            x0 = y0 = size // 2
            g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
            g_x = max(0, -ul[0]), min(br[0], heatmap_size[0]) - ul[0]
            g_y = max(0, -ul[1]), min(br[1], heatmap_size[1]) - ul[1]
            img_x = max(0, ul[0]), min(br[0], heatmap_size[0])
            img_y = max(0, ul[1]), min(br[1], heatmap_size[1])
            target[j][img_y[0]:img_y[1], img_x[0]:img_x[1]] = np.maximum(
                target[j][img_y[0]:img_y[1], img_x[0]:img_x[1]], g[g_y[0]:g_y[1], g_x[0]:g_x[1]])
        target = np.clip(target, 0, 1)
    return target, weight


def target_volume(roots, space_size, space_center, cube_size, sigma=200.0):
    """``roots``: ``[n_people, 3]`` -> ``[X,Y,Z]`` float32 (:304-341, integer ``root_id`` branch)."""
    g1 = [np.linspace(-space_size[a] / 2, space_size[a] / 2, cube_size[a]) + space_center[a] for a in range(3)]
    target = np.zeros((cube_size[0], cube_size[1], cube_size[2]), dtype=np.float32)
    for mu in roots:
        lo = [np.searchsorted(g1[a], mu[a] - 3 * sigma) for a in range(3)]
        hi = [np.searchsorted(g1[a], mu[a] + 3 * sigma, "right") for a in range(3)]
        if any(lo[a] >= hi[a] for a in range(3)):
            continue
        gx, gy, gz = np.meshgrid(g1[0][lo[0]:hi[0]], g1[1][lo[1]:hi[1]], g1[2][lo[2]:hi[2]], indexing="ij")
        g = np.exp(-((gx - mu[0]) ** 2 + (gy - mu[1]) ** 2 + (gz - mu[2]) ** 2) / (2 * sigma ** 2))
        sl = (slice(lo[0], hi[0]), slice(lo[1], hi[1]), slice(lo[2], hi[2]))
        target[sl] = np.maximum(target[sl], g)
    return np.clip(target, 0, 1)
