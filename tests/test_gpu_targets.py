"""GPU: training targets rendered on the device (``selfpose3d_b200.targets``: ``sp3d_target_heatmaps`` /
``sp3d_target_volume``) against the vectors recorded from the unmodified reference's
``JointsDataset.generate_target_heatmap`` / ``generate_3d_target`` (tests/golden/targets.npz) -- bit-exact for the 2-D
heat-maps (the Gaussian window is the reference's own numpy table), 1e-6 for the float64 3-D Gaussian -- and against
the numpy oracle on a larger seeded batch."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

from oracle import targets as otargets  # noqa: E402
from selfpose3d_b200 import targets  # noqa: E402

DEV = "cuda:0"


def test_target_heatmaps_and_volumes_match_reference_golden(golden):
    g = golden("targets")
    counts = torch.from_numpy(g["counts"].astype(np.int32)).to(DEV)
    t, w = targets.heatmap_targets(torch.from_numpy(g["joints"]).to(DEV), torch.from_numpy(g["joints_vis"]).to(DEV), counts,
                                   g["image_size"], g["heatmap_size"], sigma=3)
    assert np.array_equal(t.cpu().numpy(), g["target"])
    assert np.array_equal(w.cpu().numpy(), g["weight"])
    v = targets.root_targets(torch.from_numpy(g["roots"]).to(DEV), counts, g["space_size"], g["space_center"],
                             [int(c) for c in g["cube_size"]])
    np.testing.assert_allclose(v.cpu().numpy(), g["volume"], rtol=0, atol=1e-6)


def test_targets_match_the_oracle_on_a_batch_of_views():
    rs = np.random.RandomState(3)
    n, P, J, image, hm = 40, 10, 15, (960, 512), (240, 128)          # 8 samples x 5 views, Panoptic-shaped maps
    counts = rs.randint(0, P + 1, n)
    joints = rs.uniform(-50, 1010, (n, P, J, 2))
    vis = (rs.rand(n, P, J, 2) > 0.15).astype(np.float64)
    vis[..., 1] = vis[..., 0]
    roots = np.stack([rs.uniform(-4200, 4200, (n, P)), rs.uniform(-4700, 3700, (n, P)), rs.uniform(-300, 1900, (n, P))], -1)
    cnt = torch.from_numpy(counts.astype(np.int32)).to(DEV)
    t, w = targets.heatmap_targets(torch.from_numpy(joints).to(DEV), torch.from_numpy(vis).to(DEV), cnt, image, hm)
    v = targets.root_targets(torch.from_numpy(roots).to(DEV), cnt, [8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0], [80, 80, 20])
    t, w, v = t.cpu().numpy(), w.cpu().numpy(), v.cpu().numpy()
    for i in range(n):
        k = int(counts[i])
        if k:
            want_t, want_w = otargets.target_heatmap([joints[i, p] for p in range(k)], [vis[i, p] for p in range(k)], image, hm)
            assert np.array_equal(t[i], want_t) and np.array_equal(w[i], want_w), i
        else:
            assert not t[i].any() and not w[i].any()
        want_v = otargets.target_volume(roots[i, :k], [8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0], [80, 80, 20])
        np.testing.assert_allclose(v[i], want_v, rtol=0, atol=1e-6)
