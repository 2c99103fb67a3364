"""GPU parity: the CUDA path (through the module surface -> ctypes -> libsp3d C ABI) against the
golden vectors of the unmodified reference and against the CPU oracle on seeded inputs.

Run on the B200 box: ``python -m pytest tests -m gpu``.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

from conftest import cams_from_arrays, cam_arrays

pytestmark = pytest.mark.gpu

from oracle import geometry, nets, pipeline, volume_ops  # noqa: E402
from selfpose3d_b200 import ops, synthetic, _lib  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import (cuboid_proposal_net, cuboid_proposal_net_soft, multi_person_posenet_ssv,  # noqa: E402
                                    pose_regression_net, pose_resnet, project_layer, v2v_net)

DEV = "cuda:0"


def meta_from_golden(g, batch_slice=None):
    V = g["center"].shape[0]
    meta = []
    for c in range(V):
        sl = slice(None) if batch_slice is None else batch_slice
        meta.append({
            "center": torch.from_numpy(g["center"][c][sl]),
            "scale": torch.from_numpy(g["scale"][c][sl]),
            "rotation": torch.from_numpy(g["rotation"][c][sl]),
            "camera": {k[4:]: torch.from_numpy(g[k][c][sl]) for k in g if k.startswith("cam_")},
        })
    return meta


def cfg_for(g, **kw):
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE = [int(v) for v in g["image_size"]]
    cfg.NETWORK.HEATMAP_SIZE = [int(v) for v in g["heatmap_size"]]
    for k, v in kw.items():
        cfg.NETWORK[k] = v
    return cfg


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("name", ["project_layer_root", "project_layer_pose"])
@pytest.mark.parametrize("channel_last", [False, True])
def test_unproject_matches_reference_golden(golden, name, channel_last):
    g = golden(name)
    layer = project_layer.ProjectLayer(cfg_for(g))
    hms = [torch.from_numpy(h).to(DEV) for h in g["heatmaps"]]
    if channel_last:  # the layout the backbone hands over: stride(1) == 1
        hms = [h.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) for h in hms]
    gc = g["grid_center"]
    grid_center = [list(gc[0])] if gc.shape[1] == 3 else torch.from_numpy(gc).to(DEV)
    flip = torch.from_numpy(g["flip"]) if "flip" in g else None
    cubes, grids = layer(hms, meta_from_golden(g), list(g["grid_size"]), grid_center, list(g["cube_size"]),
                         flip_xcoords=flip)
    assert cubes.shape == g["cubes"].shape and grids.shape == g["grids"].shape
    assert np.array_equal(grids.cpu().numpy(), g["grids"])                      # voxel coordinates: bit-exact
    # oracle (float32 numpy restatement, same operation order): expected bit-exact
    o_cubes, _, margin = geometry.unproject(
        g["heatmaps"], cams_from_arrays(g), g["center"], g["scale"], g["rotation"], g["image_size"],
        g["heatmap_size"], g["grid_size"], g["grid_center"], g["cube_size"], flip=g.get("flip"), return_aux=True)
    got = cubes.cpu().numpy()
    n_diff = int((got != o_cubes).sum())
    assert np.abs(got - o_cubes).max() <= 1e-6, (n_diff, np.abs(got - o_cubes).max())
    # reference golden: 1e-4 on [0,1] values; voxels projecting within 1e-2 px of an image border may flip
    amb = (margin < 1e-2)
    B, C = got.shape[:2]
    diff = np.where(amb[:, None], 0, np.abs(got.reshape(B, C, -1) - g["cubes"].reshape(B, C, -1)))
    assert diff.max() <= 1e-4 and amb.mean() < 1e-3


def test_unproject_seeded_vs_oracle_many_channels():
    """C = 17 (> one 16-channel register group), 5 views, rotation + scale + flip, bf16 + padded output."""
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [96, 128], [24, 32]
    cams = synthetic.ring_cameras(5, seed=7)
    B, C = 2, 17
    meta = synthetic.make_meta(cams, B, (96, 128), rotation=[[5.0, -12.0]] * 5, scale_mul=[[1.1, 0.9]] * 5)
    rs = np.random.RandomState(3)
    hm_np = rs.rand(5, B, C, 32, 24).astype(np.float32)
    centers = np.array([[200.0, -700.0, 900.0, 0.0, 1.0], [-900.0, 300.0, 1000.0, 2.0, 1.0]], dtype=np.float32)
    flip = np.array([True, False])
    layer = project_layer.ProjectLayer(cfg)
    hms = [torch.from_numpy(h).to(DEV) for h in hm_np]
    cubes, grids = layer(hms, meta, [2000.0] * 3, torch.from_numpy(centers).to(DEV), [12, 8, 16],
                         flip_xcoords=torch.from_numpy(flip))
    cams_nested = [[{k: np.asarray(v[i]) for k, v in m["camera"].items()} for i in range(B)] for m in meta]
    o_cubes, o_grids = geometry.unproject(
        hm_np, cams_nested, [m["center"].numpy() for m in meta], [m["scale"].numpy() for m in meta],
        [m["rotation"].numpy() for m in meta], (96, 128), (24, 32), [2000.0] * 3, centers, [12, 8, 16], flip=flip)
    assert np.array_equal(grids.cpu().numpy(), o_grids)
    assert np.abs(cubes.cpu().numpy() - o_cubes).max() <= 1e-6
    # layout-native path: channel-last padded output, float32 and bf16
    camt = ops.pack_cameras(meta, cfg.NETWORK.IMAGE_SIZE, flip).to(DEV)
    cen = torch.from_numpy(centers).to(DEV)
    cl, _ = layer.project_cl(hms, camt, cen, True, [2000.0] * 3, [12, 8, 16])
    assert cl.shape == (2, 12, 8, 16, 20) and not cl[..., 17:].any()
    assert np.abs(cl[..., :17].permute(0, 4, 1, 2, 3).cpu().numpy() - o_cubes).max() <= 1e-6
    clb, _ = layer.project_cl(hms, camt, cen, True, [2000.0] * 3, [12, 8, 16], dtype=torch.bfloat16)
    assert np.abs(clb[..., :17].permute(0, 4, 1, 2, 3).float().cpu().numpy() - o_cubes).max() <= 4e-3  # bf16 rounding


@pytest.mark.parametrize("cube", [[12, 8, 16], [10, 13, 21]])    # whole tiles / ragged in x, y and z
def test_unproject_throughput_form_vs_oracle(cube):
    """math_mode 1 (fp16 channel-last maps, half2 tap blending, bf16 cubes) against the float32 numpy oracle:
    15 channels, 5 views, rotation + scale + flip, one invalid cube.  Tolerance = bf16 output step (3.9e-3) plus the
    half-precision blend (~1e-3); voxels within 1e-2 px of an image border may take the other in-image decision
    and are counted, not compared."""
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [96, 128], [24, 32]
    cams = synthetic.ring_cameras(5, seed=7)
    B, C = 2, 15
    meta = synthetic.make_meta(cams, B, (96, 128), rotation=[[5.0, -12.0]] * 5, scale_mul=[[1.1, 0.9]] * 5)
    rs = np.random.RandomState(3)
    hm_np = rs.rand(5, B, C, 32, 24).astype(np.float32)
    centers = np.array([[200.0, -700.0, 900.0, 0.0, 1.0], [-900.0, 300.0, 1000.0, 2.0, 1.0],
                        [100.0, 100.0, 700.0, -1.0, 0.0]], dtype=np.float32)
    flip = np.array([True, False])
    layer = project_layer.ProjectLayer(cfg)
    hms = [torch.from_numpy(h).to(DEV) for h in hm_np]
    camt = ops.pack_cameras(meta, cfg.NETWORK.IMAGE_SIZE, flip).to(DEV)
    cen = torch.from_numpy(centers).to(DEV)
    sample = torch.tensor([0, 1, 1], dtype=torch.int32, device=DEV)
    before = _lib.launch_count
    got, _ = layer.project_cl(hms, camt, cen, True, [2000.0] * 3, cube, cube_sample=sample, dtype=torch.bfloat16,
                              c_pitch=16)
    assert _lib.launch_count - before == 2          # fp16 conversion + the fused kernel
    assert got.shape == (3, cube[0], cube[1], cube[2], 16) and got.dtype == torch.bfloat16
    assert not got[..., 15:].any() and not got[2].any()      # padding channel and the invalid cube are zeros
    cams_nested = [[{k: np.asarray(v[i]) for k, v in m["camera"].items()} for i in range(B)] for m in meta]
    for q, b in ((0, 0), (1, 1)):
        want, _ = geometry.unproject(
            hm_np[:, b:b + 1], [[c[b]] for c in cams_nested], [m["center"].numpy()[b:b + 1] for m in meta],
            [m["scale"].numpy()[b:b + 1] for m in meta], [m["rotation"].numpy()[b:b + 1] for m in meta], (96, 128),
            (24, 32), [2000.0] * 3, centers[q:q + 1], cube, flip=flip[b:b + 1])
        diff = np.abs(got[q, ..., :15].permute(3, 0, 1, 2).float().cpu().numpy() - want[0])
        bad = diff > 6e-3
        assert bad.mean() < 2e-3, (bad.mean(), diff.max())     # border-decision voxels only
        assert np.median(diff) < 1.5e-3


def test_unproject_full_size_throughput_vs_float32_form():
    """BASELINE size (64^3 x 15 channels, 96x72 maps, 5 views): the throughput form against the oracle-validated
    float32 form on the same inputs, plus size-independent properties -- linearity in the heat-maps (no clamp active
    for maps in [0, 0.5]) and values inside [0, 1]."""
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [288, 384], [72, 96]
    B, C = 2, 15
    meta = synthetic.make_meta(synthetic.ring_cameras(5, seed=0), B, (288, 384))
    g = torch.Generator().manual_seed(5)
    h1 = [torch.rand(B, C, 96, 72, generator=g).mul_(0.5).to(DEV) for _ in range(5)]
    h2 = [torch.rand(B, C, 96, 72, generator=g).mul_(0.5).to(DEV) for _ in range(5)]
    cams = ops.pack_cameras(meta, cfg.NETWORK.IMAGE_SIZE).to(DEV)
    cen = torch.tensor([[300.0, -900.0, 900.0, 0.0, 1.0], [-1500.0, 400.0, 1000.0, 1.0, 1.0]], device=DEV)
    layer = project_layer.ProjectLayer(cfg)

    def run(hms, dtype, pitch):
        return layer.project_cl(hms, cams, cen, True, [2000.0] * 3, [64] * 3, dtype=dtype, c_pitch=pitch)[0][..., :C].float()

    exact = run(h1, torch.float32, 16)
    fast = run(h1, torch.bfloat16, 16)
    diff = (fast - exact).abs()
    assert float((diff > 6e-3).float().mean()) < 2e-3 and float(diff.median()) < 1.5e-3
    assert float(fast.min()) >= 0.0 and float(fast.max()) <= 1.0
    both = run([a + b for a, b in zip(h1, h2)], torch.bfloat16, 16)
    lin = (both - (fast + run(h2, torch.bfloat16, 16))).abs()
    assert float((lin > 1.2e-2).float().mean()) < 2e-3       # three bf16 roundings + border-decision voxels


def test_unproject_view_sharded_partial_sums_equal_full():
    """Multi-GPU exchange: sum of per-view-shard partial numerators/counts + finalize == full."""
    g_cfg = default_config()
    g_cfg.NETWORK.IMAGE_SIZE, g_cfg.NETWORK.HEATMAP_SIZE = [72, 96], [18, 24]
    cams = synthetic.ring_cameras(5, seed=1)
    B, C = 2, 3
    meta = synthetic.make_meta(cams, B, (72, 96))
    people = synthetic.synthetic_people(B, seed=5, num_joints=C)
    hms = [h.to(DEV) for h in synthetic.render_heatmaps(people, meta, (72, 96), (18, 24), num_joints=C, sigma=1.5)]
    camt = ops.pack_cameras(meta, (72, 96)).to(DEV)
    cen = torch.tensor([[0.0, -500.0, 800.0]] * B, device=DEV)
    X, Y, Z = 20, 20, 8
    N = X * Y * Z
    full = torch.empty(B, C, N, device=DEV)
    ops.unproject(hms, hms[0].stride(), camt, cen, [8000.0, 8000.0, 2000.0], (X, Y, Z), (72, 96), (24, 18), C,
                  full, (C * N, N, 1))
    total = torch.zeros(B, C + 1, N, device=DEV)
    for (v0, v1) in [(0, 2), (2, 3), (3, 5)]:
        part = torch.empty(B, C + 1, N, device=DEV)
        ops.unproject(hms, hms[0].stride(), camt, cen, [8000.0, 8000.0, 2000.0], (X, Y, Z), (72, 96), (24, 18), C,
                      part, ((C + 1) * N, N, 1), view_range=(v0, v1), partial=True)
        total += part
    ops.unproject_finalize(total, B, C, N, ((C + 1) * N, N, 1))
    assert torch.allclose(total[:, :C], full, rtol=0, atol=2e-7)  # summation order over views differs by <= 1 ulp


def test_unproject_empty_and_all_invalid():
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [72, 96], [18, 24]
    cams = synthetic.ring_cameras(5, seed=1)
    meta = synthetic.make_meta(cams, 2, (72, 96))
    hms = [torch.rand(2, 4, 24, 18, device=DEV) for _ in range(5)]
    layer = project_layer.ProjectLayer(cfg)
    centers = torch.tensor([[0.0, 0.0, 0.0, -1.0, 0.0]] * 2, device=DEV)
    cubes, grids = layer(hms, meta, [2000.0] * 3, centers, [8, 8, 8])
    assert not cubes.any() and not grids.any()
    camt = ops.pack_cameras(meta, (72, 96)).to(DEV)
    cl, _ = layer.project_cl(hms, camt, centers[:0], True, [2000.0] * 3, [8, 8, 8])
    assert cl.shape[0] == 0


# ------------------------------------------------------------------------------------------ K3
def test_nms_topk_matches_reference_golden(golden):
    g = golden("proposal")
    x = torch.from_numpy(g["root_cubes"]).to(DEV)
    gc, idx = ops.nms_topk(x, int(g["max_people"]), float(g["threshold"]), g["space_size"], g["space_center"],
                           return_index=True)
    assert np.array_equal(gc.cpu().numpy(), g["grid_centers"])      # bit-exact incl. get_real_loc rounding
    X, Y, Z = g["root_cubes"].shape[1:]
    flat = g["topk_index"][..., 0] * Y * Z + g["topk_index"][..., 1] * Z + g["topk_index"][..., 2]
    assert np.array_equal(idx.cpu().numpy(), flat)
    gc64 = ops.nms_topk(x, int(g["max_people"]), float(g["threshold"]), g["space_size"], g["space_center"], loc_f64=True)
    assert np.array_equal(gc64.cpu().numpy(), golden("proposal_f64")["grid_centers"])


@pytest.mark.parametrize("shape,K", [((2, 80, 80, 20), 10), ((1, 5, 4, 3), 32), ((3, 16, 16, 8), 1)])
def test_nms_topk_seeded_vs_oracle(shape, K):
    rs = np.random.RandomState(5)
    x = rs.rand(*shape).astype(np.float32)
    x[0, :2, :2, :2] = 0.75            # plateau: equal neighbours are all "maxima" -> ties, lowest index first
    x[-1] = np.round(x[-1] * 8) / 8    # heavy ties
    if shape[0] > 1:
        x[1] -= 2.0                    # all-negative sample: zeros of the non-maxima outrank the maxima
    want = volume_ops.proposal_layer(x, [8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0], shape[1:], K, 0.3)
    got = ops.nms_topk(torch.from_numpy(x).to(DEV), K, 0.3, [8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0])
    assert np.array_equal(got.cpu().numpy(), want)
    scores = got[:, :, 4].cpu().numpy()
    assert (np.diff(scores, axis=1) <= 0).all()                     # sortedness at full size


# ------------------------------------------------------------------------------------------ K4
def test_softargmax_matches_reference_golden(golden):
    g = golden("softargmax")
    layer = pose_regression_net.SoftArgmaxLayer(default_config())
    # the golden grid is random (not separable): evaluate through the kernel per axis-separable piece instead
    x = torch.from_numpy(g["x"]).to(DEV)
    B, C = x.shape[:2]
    truth = volume_ops.soft_argmax(g["x"], g["grids"], 100.0, dtype=np.float64)
    # separable grid built by ProjectLayer.compute_grid
    grid = project_layer.ProjectLayer(default_config()).compute_grid([500.0, 400.0, 300.0], [10.0, -20.0, 30.0],
                                                                      list(x.shape[2:]))
    grids = grid[None].repeat(B, 1, 1)
    got = layer(x, grids.to(DEV)).cpu().numpy()
    want64 = volume_ops.soft_argmax(g["x"], grids.numpy(), 100.0, dtype=np.float64)
    want32 = volume_ops.soft_argmax(g["x"], grids.numpy(), 100.0, dtype=np.float32)
    ref_err = np.abs(want32 - want64).max()
    assert np.abs(got - want64).max() <= 1e-3, np.abs(got - want64).max()          # mm, north-star tolerance
    assert np.abs(got - want32).max() <= 1e-3 + ref_err
    assert truth.shape == got.shape


def test_softargmax_full_size_properties():
    """64^3 x 15: (a) one-hot-like cube returns its voxel coordinate; (b) shifting the cube centre shifts
    the result by exactly that vector; (c) agrees with the float64 oracle within 1e-3 mm."""
    rs = np.random.RandomState(8)
    X = Y = Z = 64
    C = 15
    x = (rs.rand(2, X, Y, Z, 16) * 0.3).astype(np.float32)
    x[0, 10, 20, 30, 3] = 5.0
    xt = torch.from_numpy(x).to(DEV)
    cen = torch.tensor([[100.0, -200.0, 900.0], [-1500.0, 2500.0, 700.0]], device=DEV)
    out = ops.softargmax(xt, (X * Y * Z * 16, 1, 16), 2, C, (X, Y, Z), cen, [2000.0] * 3, 100.0).cpu().numpy()
    lin = [a.cpu().numpy() for a in ops.linspace_axes([2000.0] * 3, (X, Y, Z), DEV)]
    voxel = np.array([lin[0][10] + np.float32(100.0), lin[1][20] + np.float32(-200.0), lin[2][30] + np.float32(900.0)])
    assert np.abs(out[0, 3] - voxel).max() < 1e-3
    cen2 = cen + torch.tensor([[512.0, -256.0, 128.0]], device=DEV)   # exactly representable shift
    out2 = ops.softargmax(xt, (X * Y * Z * 16, 1, 16), 2, C, (X, Y, Z), cen2, [2000.0] * 3, 100.0).cpu().numpy()
    assert np.abs((out2 - out) - np.array([512.0, -256.0, 128.0])).max() < 1e-3
    grids = np.stack([geometry.compute_grid([2000.0] * 3, c, (X, Y, Z)) for c in cen.cpu().numpy()])
    want = volume_ops.soft_argmax(np.moveaxis(x[..., :C], -1, 1), grids, 100.0, dtype=np.float64)
    assert np.abs(out - want).max() <= 1e-3, np.abs(out - want).max()


# ------------------------------------------------------------------------------------------ convolutions
def test_v2v_blocks_match_reference_golden(golden):
    g = golden("v2v_blocks")
    x = torch.from_numpy(g["x"]).to(DEV)
    blocks = {"basic7": v2v_net.Basic3DBlock(4, 8, 7), "basic3": v2v_net.Basic3DBlock(4, 8, 3),
              "res_4_8": v2v_net.Res3DBlock(4, 8), "res_4_4": v2v_net.Res3DBlock(4, 4),
              "up_4_8": v2v_net.Upsample3DBlock(4, 8, 2, 2)}
    for name, mod in blocks.items():
        pre = "w_%s_" % name
        mod.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(pre)}, strict=True)
        y = mod.to(DEV).eval()(x)
        np.testing.assert_allclose(y.cpu().numpy(), g["y_" + name], rtol=1e-4, atol=1e-5, err_msg=name)


def test_v2v_net_matches_reference_golden(golden):
    g = golden("v2v_net")
    for tag, cin in (("1", 1), ("3", 3)):
        net = v2v_net.V2VNet(cin, cin)
        net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(g["seed" + tag])), strict=True)
        y = net.to(DEV).eval()(torch.from_numpy(g["x" + tag]).to(DEV))
        np.testing.assert_allclose(y.cpu().numpy(), g["y" + tag], rtol=1e-4, atol=1e-5)


def test_pose_resnet_matches_reference_golden(golden):
    g = golden("pose_resnet50")
    net = pose_resnet.get_pose_net(default_config(), is_train=False)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(g["seed"])), strict=True)
    y = net.to(DEV).eval()(torch.from_numpy(g["x"]).to(DEV))
    assert y.shape == g["y"].shape and y.stride(1) == 1
    np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=2e-4, atol=2e-5)


def test_v2v_pose_size_vs_oracle_fp64():
    """One 64^3 x 15 person cube through V2VNet(15,15): ours and the float32 oracle against the float64 oracle."""
    net = v2v_net.V2VNet(15, 15)
    sd = synthetic.trained_like_state_dict(net, seed=41)
    net.load_state_dict(sd, strict=True)
    rs = np.random.RandomState(2)
    x = torch.from_numpy(rs.rand(1, 15, 64, 64, 64).astype(np.float32))
    y = net.to(DEV).eval()(x.to(DEV)).cpu()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    y32 = nets.v2v_forward(x, sd)
    y64 = nets.v2v_forward(x, sd, dtype=torch.float64)
    scale = float(y64.abs().max())
    ours = float((y.double() - y64).abs().max()) / scale
    ref = float((y32.double() - y64).abs().max()) / scale
    assert ours <= max(1e-4, 4 * ref), (ours, ref)   # heat-map tolerance 1e-4 relative


def test_cpu_tensor_raises_no_fallback():
    net = v2v_net.V2VNet(1, 1).eval()
    with pytest.raises(_lib.Sp3dError):
        net(torch.zeros(1, 1, 8, 8, 8))


# ------------------------------------------------------------------------------------------ whole path
def _small_model(g0, threshold):
    cfg = cfg_for(g0)
    cfg.NETWORK.NUM_JOINTS = int(g0["num_joints"])
    cfg.MULTI_PERSON.SPACE_SIZE = [float(v) for v in g0["space_size"]]
    cfg.MULTI_PERSON.SPACE_CENTER = [float(v) for v in g0["space_center"]]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [int(v) for v in g0["initial_cube_size"]]
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = int(g0["max_people"])
    cfg.MULTI_PERSON.THRESHOLD = threshold
    cfg.PICT_STRUCT.GRID_SIZE = [float(v) for v in g0["grid_size"]]
    cfg.PICT_STRUCT.CUBE_SIZE = [int(v) for v in g0["cube_size"]]
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=int(g0["seed"])), strict=True)
    return model.to(DEV).eval(), cfg


def test_inference_from_heatmaps_matches_reference_golden(golden):
    g = golden("inference_small")
    model, _ = _small_model(g, float(g["threshold"]))
    hms = [torch.from_numpy(h).to(DEV) for h in g["heatmaps"]]
    pred, _, gc = model(views1=None, meta1=meta_from_golden(g), input_heatmaps1=hms, inference=True)
    np.testing.assert_allclose(gc.cpu().numpy(), g["grid_centers"], rtol=1e-4, atol=1e-5)
    valid = g["pred"][:, :, 0, 3] >= 0
    assert np.array_equal(pred[:, :, 0, 3].cpu().numpy() >= 0, valid)
    np.testing.assert_allclose(pred.cpu().numpy()[valid], g["pred"][valid], rtol=0, atol=1e-2)   # mm
    assert not pred.cpu().numpy()[~valid][..., :3].any()


def test_inference_bf16_throughput_mode_vs_reference_golden(golden):
    """The bench's mode (fp16 maps + half2 un-projection, tcgen05 convolutions, fused soft-argmax head) on the
    reference's own golden inference case.  bf16 activations carry 8 significant bits: proposals that land on the
    reference's voxel are compared at centimetre level; this documents the accuracy envelope of the throughput mode,
    the float32 mode above is the parity mode."""
    g = golden("inference_small")
    model, _ = _small_model(g, float(g["threshold"]))
    hms = [torch.from_numpy(h).to(DEV) for h in g["heatmaps"]]
    ops.set_volume_dtype(torch.bfloat16)
    try:
        pred, _, gc = model(views1=None, meta1=meta_from_golden(g), input_heatmaps1=hms, inference=True)
    finally:
        ops.set_volume_dtype(torch.float32)
    pred, gc = pred.cpu().numpy(), gc.cpu().numpy()
    valid = g["pred"][:, :, 0, 3] >= 0
    same = valid & (pred[:, :, 0, 3] >= 0) & (np.abs(gc[:, :, :3] - g["grid_centers"][:, :, :3]).max(-1) < 1.0)
    assert same.sum() >= max(1, valid.sum() // 2), (int(same.sum()), int(valid.sum()))
    err = np.abs(pred[same][..., :3] - g["pred"][same][..., :3])
    assert float(np.median(err)) < 10.0 and float(err.max()) < 60.0, (float(np.median(err)), float(err.max()))


def test_inference_from_images_matches_reference_golden(golden):
    g0, g = golden("inference_small"), golden("inference_images")
    model, _ = _small_model(g0, float(g["threshold"]))
    imgs = [torch.from_numpy(x).to(DEV) for x in g["images"]]
    pred, hms, gc = model(views1=imgs, meta1=meta_from_golden(g0, slice(0, 1)), inference=True)
    np.testing.assert_allclose(torch.stack(hms).cpu().numpy(), g["heatmaps"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(gc.cpu().numpy(), g["grid_centers"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred"], rtol=0, atol=1e-2)


def test_cuboid_proposal_net_all_joints_matches_reference_golden(golden):
    g0, g = golden("inference_small"), golden("cuboid_proposal_allj")
    cfg = cfg_for(g0, ROOTNET_ROOTHM=False)
    cfg.NETWORK.NUM_JOINTS = int(g0["num_joints"])
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [int(v) for v in g0["initial_cube_size"]]
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = int(g0["max_people"])
    cfg.MULTI_PERSON.THRESHOLD = float(g0["threshold"])
    net = cuboid_proposal_net.CuboidProposalNet(cfg)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(g["seed"])), strict=True)
    hms = [torch.from_numpy(h).to(DEV) for h in g0["heatmaps"]]
    rc, gc = net.to(DEV).eval()(hms, meta_from_golden(g0))
    np.testing.assert_allclose(rc.cpu().numpy(), g["root_cubes"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gc.cpu().numpy()[..., :3], g["grid_centers"][..., :3], rtol=1e-5, atol=1e-5)
