"""Pin the CPU oracle (``oracle/``) against golden vectors produced by the unmodified
reference (``tests/golden/make_golden.py``).  CPU only."""
import numpy as np
import torch

from conftest import cams_from_arrays, cam_arrays
from oracle import geometry, nets, pipeline, volume_ops
from selfpose3d_b200 import synthetic


def test_affine_matches_reference(golden):
    g = golden("affine")
    for case, want in zip(g["cases"], g["trans"]):
        got = geometry.get_affine_transform(case[0:2], case[2:4].astype(np.float32), case[4], case[5:7])
        assert np.array_equal(got, want)  # bit-identical float64 matrix (cv2's elimination order restated)


def test_get_scale_branches():
    np.testing.assert_allclose(geometry.get_scale((1920, 1080), (960, 512)), [10.125, 5.4], rtol=1e-7)
    np.testing.assert_allclose(geometry.get_scale((1920, 1080), (288, 384)), [9.6, 12.8], rtol=1e-7)


def test_project_point_radial_matches_reference(golden):
    g = golden("project_pose")
    for v in range(g["pixels"].shape[0]):
        got = geometry.project_point_radial(
            g["points"], g["cam_R"][v], g["cam_T"][v], [g["cam_fx"][v], g["cam_fy"][v]],
            [g["cam_cx"][v], g["cam_cy"][v]], g["cam_k"][v], g["cam_p"][v])
        # pixel coordinates O(1e3): float32 ulp is 1.2e-4 px; mm/einsum summation order differs
        np.testing.assert_allclose(got, g["pixels"][v], rtol=3e-6, atol=2e-3)  # rtol: points near/behind the camera plane project to 1e10


def _unproject_case(g, dtype):
    cams = cams_from_arrays(g)
    return geometry.unproject(
        g["heatmaps"], cams, g["center"], g["scale"], g["rotation"], g["image_size"], g["heatmap_size"],
        g["grid_size"], g["grid_center"], g["cube_size"], flip=g.get("flip"), dtype=dtype, return_aux=True)


def _check_cubes(got, want, margin, atol=1e-4):
    """Tolerance: the north star's 1e-4 on values in [0, 1] (a float32 ulp of the O(1e3) pixel
    coordinate is 6e-5 px; times the heat-map slope that is ~1e-5 in value).  Voxels whose projection is within 1e-2 px of an image border in some view may flip
    their in-image mask under a 1-ulp coordinate difference; they are counted, not compared."""
    B, C = want.shape[:2]
    ambiguous = (margin < 1e-2)
    diff = np.abs(got.reshape(B, C, -1) - want.reshape(B, C, -1))
    diff = np.where(ambiguous[:, None], 0.0, diff)
    assert ambiguous.mean() < 1e-3
    assert diff.max() <= atol, diff.max()


def test_unproject_root_matches_reference(golden):
    g = golden("project_layer_root")
    cubes, grids, margin = _unproject_case(g, np.float32)
    assert np.array_equal(grids, g["grids"])  # bit-identical voxel coordinates
    _check_cubes(cubes, g["cubes"], margin)
    cubes64, _, _ = _unproject_case(g, np.float64)
    _check_cubes(cubes64, g["cubes"], margin)


def test_unproject_pose_matches_reference(golden):
    g = golden("project_layer_pose")
    cubes, grids, margin = _unproject_case(g, np.float32)
    assert np.array_equal(grids, g["grids"])
    assert not cubes[1].any() and not grids[1].any()  # invalid proposal row stays zero
    _check_cubes(cubes, g["cubes"], margin)


def test_unproject_torch_port_matches_reference(golden):
    for name in ("project_layer_root", "project_layer_pose"):
        g = golden(name)
        hms = [torch.from_numpy(h) for h in g["heatmaps"]]
        cubes, grids = pipeline.unproject_torch(
            hms, cam_arrays(g), g["center"], g["scale"], g["rotation"], g["image_size"], g["heatmap_size"],
            g["grid_size"], g["grid_center"], g["cube_size"], flip=g.get("flip"))
        assert np.array_equal(grids.numpy(), g["grids"])
        np.testing.assert_allclose(cubes.numpy(), g["cubes"], rtol=0, atol=1e-4)


def test_nms_and_proposals_match_reference(golden):
    g = golden("proposal")
    vals, idx = volume_ops.nms(g["root_cubes"], int(g["max_people"]))
    assert np.array_equal(vals, g["topk_values"])
    assert np.array_equal(idx, g["topk_index"])
    gc = volume_ops.proposal_layer(g["root_cubes"], g["space_size"], g["space_center"], g["cube_size"],
                                   int(g["max_people"]), float(g["threshold"]))
    assert np.array_equal(gc, g["grid_centers"])
    assert (gc[2, :, 3] == -1).all() and (gc[0, :, 3] == 0).all()
    gc64 = volume_ops.proposal_layer(g["root_cubes"], g["space_size"], g["space_center"], g["cube_size"],
                                     int(g["max_people"]), float(g["threshold"]), f64=True)
    assert np.array_equal(gc64, golden("proposal_f64")["grid_centers"])
    assert not np.array_equal(gc64, gc)  # the two config dtypes really differ in the last ulp


def test_soft_argmax_matches_reference(golden):
    g = golden("softargmax")
    got = volume_ops.soft_argmax(g["x"], g["grids"], float(g["beta"]))
    np.testing.assert_allclose(got, g["out"], rtol=0, atol=2e-3)  # mm; coordinates O(1e3)
    truth = volume_ops.soft_argmax(g["x"], g["grids"], float(g["beta"]), dtype=np.float64)
    # the reference's own fp32 error against the fp64 up-cast, for the tolerance budget
    assert np.abs(g["out"] - truth).max() < 2e-3


def test_v2v_blocks_match_reference(golden):
    g = golden("v2v_blocks")
    x = torch.from_numpy(g["x"])
    fns = {"basic7": nets.basic3d, "basic3": nets.basic3d, "res_4_8": nets.res3d, "res_4_4": nets.res3d,
           "up_4_8": nets.upsample3d}
    for name, fn in fns.items():
        pre = "w_%s_" % name
        sd = {"m." + k[len(pre):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre)}
        y = fn(x, sd, "m")
        np.testing.assert_allclose(y.numpy(), g["y_" + name], rtol=1e-5, atol=1e-6)


def _ref_shaped_v2v_sd(cin, cout, seed):
    from selfpose3d_b200.models.v2v_net import V2VNet
    return synthetic.trained_like_state_dict(V2VNet(cin, cout), seed=seed)


def test_v2v_net_matches_reference(golden):
    g = golden("v2v_net")
    y1 = nets.v2v_forward(torch.from_numpy(g["x1"]), _ref_shaped_v2v_sd(1, 1, int(g["seed1"])))
    np.testing.assert_allclose(y1.numpy(), g["y1"], rtol=1e-4, atol=1e-6)
    y3 = nets.v2v_forward(torch.from_numpy(g["x3"]), _ref_shaped_v2v_sd(3, 3, int(g["seed3"])))
    np.testing.assert_allclose(y3.numpy(), g["y3"], rtol=1e-4, atol=1e-6)


def test_pose_resnet_matches_reference(golden):
    from selfpose3d_b200.config import default_config
    from selfpose3d_b200.models import pose_resnet
    g = golden("pose_resnet50")
    net = pose_resnet.get_pose_net(default_config(), is_train=False)
    sd = synthetic.trained_like_state_dict(net, seed=int(g["seed"]))
    y = nets.pose_resnet_forward(torch.from_numpy(g["x"]), sd)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=1e-5)


def _small_cfg(g):
    return dict(image_size=g["image_size"], heatmap_size=g["heatmap_size"], space_size=g["space_size"],
                space_center=g["space_center"], initial_cube_size=g["initial_cube_size"],
                grid_size=g["grid_size"], cube_size=g["cube_size"], max_people=int(g["max_people"]),
                threshold=float(g["threshold"]), beta=100.0, root_idx=2)


def _small_model_sd(seed, num_joints):
    from selfpose3d_b200.config import default_config
    from selfpose3d_b200.models import multi_person_posenet_ssv
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = num_joints
    cfg.NETWORK.ROOTNET_ROOTHM = True
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    return synthetic.trained_like_state_dict(model, seed=seed)


def test_inference_pipeline_matches_reference(golden):
    g = golden("inference_small")
    sd = _small_model_sd(int(g["seed"]), int(g["num_joints"]))
    hms = [torch.from_numpy(h) for h in g["heatmaps"]]
    pred, _, gc, root_cubes = pipeline.inference(sd, _small_cfg(g), cam_arrays(g), g["center"], g["scale"],
                                                 g["rotation"], heatmaps=hms)
    np.testing.assert_allclose(root_cubes.numpy(), g["root_cubes"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gc.numpy(), g["grid_centers"], rtol=1e-5, atol=1e-6)
    valid = g["pred"][:, :, 0, 3] >= 0
    assert valid.any() and (~valid).any()
    np.testing.assert_allclose(pred.numpy()[valid], g["pred"][valid], rtol=0, atol=1e-2)  # mm


def test_inference_from_images_matches_reference(golden):
    g0 = golden("inference_small")
    g = golden("inference_images")
    sd = _small_model_sd(int(g["seed"]), int(g0["num_joints"]))
    cfg = _small_cfg(g0)
    cfg["threshold"] = float(g["threshold"])
    ca = {k: v[:, :1] for k, v in cam_arrays(g0).items()}
    imgs = [torch.from_numpy(x) for x in g["images"]]
    pred, hms, gc, _ = pipeline.inference(sd, cfg, ca, g0["center"][:, :1], g0["scale"][:, :1],
                                          g0["rotation"][:, :1], images=imgs)
    np.testing.assert_allclose(torch.stack(hms).numpy(), g["heatmaps"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gc.numpy(), g["grid_centers"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=0, atol=1e-2)


# ---------------------------------------------------------------------------------------------- gradients
def test_backward_oracle_matches_reference_gradients(golden):
    """oracle/backward.py (autograd over the oracle's torch restatements) against gradients recorded from the
    unmodified reference modules (tests/golden/make_golden_backward.py)."""
    from oracle import backward
    gb = golden("backward")
    g = golden("project_layer_pose")
    hms = [torch.from_numpy(h) for h in g["heatmaps"]]
    grads, cubes = backward.unproject_grad(
        hms, cam_arrays(g), g["center"], g["scale"], g["rotation"], g["image_size"], g["heatmap_size"],
        g["grid_size"], g["grid_center"], g["cube_size"], torch.from_numpy(gb["pl_grad_cubes"]), flip=g.get("flip"))
    want = gb["pl_grad_heatmaps"]
    got = np.stack([t.numpy() for t in grads])
    assert np.abs(want).max() > 0 and not got[:, 1].any()        # the invalid proposal row receives no gradient
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4 * np.abs(want).max())

    s = golden("softargmax")
    gx, _ = backward.softargmax_grad(torch.from_numpy(s["x"]), torch.from_numpy(s["grids"]), float(s["beta"]),
                                     torch.from_numpy(gb["sa_grad_out"]))
    np.testing.assert_allclose(gx.numpy(), gb["sa_grad_x"], rtol=0, atol=1e-5 * np.abs(gb["sa_grad_x"]).max())

    t = {k: torch.from_numpy(gb["b3_" + k]) for k in ("x", "w", "b", "gamma", "beta", "grad_y")}
    y, dx, dw, db, dg, dbeta, _, _ = backward.basic3d_train(t["x"], t["w"], t["b"], t["gamma"], t["beta"], t["grad_y"], 3)
    for got_, name in ((y, "y"), (dx, "grad_x"), (dw, "grad_w"), (dg, "grad_gamma"), (dbeta, "grad_beta")):
        want_ = gb["b3_" + name]
        np.testing.assert_allclose(got_.numpy(), want_, rtol=0, atol=2e-5 * max(np.abs(want_).max(), 1e-3), err_msg=name)
    # (the conv bias gradient is annihilated by the batch normalisation: both sides are rounding noise)
    assert np.abs(db.numpy()).max() < 1e-3 and np.abs(gb["b3_grad_b"]).max() < 1e-3


def _v2v_train_oracle(gb):
    """Autograd through oracle.nets.v2v_forward(training=True) on the inputs of the recorded reference step."""
    from selfpose3d_b200.models import v2v_net
    net = v2v_net.V2VNet(3, 3)
    sd0 = synthetic.trained_like_state_dict(net, seed=int(gb["v2v_seed"]))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in sd0.items()}
    x = torch.from_numpy(gb["v2v_x"]).clone().requires_grad_(True)
    y = nets.v2v_forward(x, sd, training=True)
    (y * torch.from_numpy(gb["v2v_grad_y"])).sum().backward()
    return y.detach(), x.grad, sd


def test_v2v_training_oracle_matches_reference_step(golden):
    gb = golden("backward")
    y, gx, sd = _v2v_train_oracle(gb)
    np.testing.assert_allclose(y.numpy(), gb["v2v_y"], rtol=0, atol=1e-5 * np.abs(gb["v2v_y"]).max())
    np.testing.assert_allclose(gx.numpy(), gb["v2v_grad_x"], rtol=0, atol=1e-4 * np.abs(gb["v2v_grad_x"]).max())
    for name, norm, tot in zip(gb["v2v_param_names"], gb["v2v_param_grad_norm"], gb["v2v_param_grad_sum"]):
        g = sd[str(name)].grad.double()
        assert abs(float(g.norm()) - norm) <= 1e-3 * max(norm, 1e-3), name
        assert abs(float(g.sum()) - tot) <= 1e-3 * max(norm, 1e-3) * max(1.0, np.sqrt(g.numel())), name


def test_target_generation_oracle_matches_reference_golden(golden):
    """oracle/targets.py (JointsDataset.generate_target_heatmap / generate_3d_target restated) against the vectors
    recorded from the unmodified reference (tests/golden/make_golden_targets.py)."""
    from oracle import targets
    g = golden("targets")
    for i, n in enumerate(g["counts"]):
        n = int(n)
        if n:
            t, w = targets.target_heatmap([g["joints"][i, p] for p in range(n)], [g["joints_vis"][i, p] for p in range(n)],
                                          g["image_size"], g["heatmap_size"], sigma=3)
            assert np.array_equal(t, g["target"][i]) and np.array_equal(w, g["weight"][i])
        else:
            assert not g["target"][i].any() and not g["weight"][i].any()
        v = targets.target_volume(g["roots"][i, :n], g["space_size"], g["space_center"], [int(c) for c in g["cube_size"]])
        assert np.array_equal(v, g["volume"][i])
    assert g["target"].max() == 1.0 and g["volume"].max() > 0.5
