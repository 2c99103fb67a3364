"""CPU check of the supervised training forward (``models/multi_person_posenet.py`` ``_forward_train``, the training
branches of ``CuboidProposalNet`` / ``PoseRegressionNet``) against ONE TRAINING STEP RECORDED FROM THE UNMODIFIED
REFERENCE (``tests/golden/make_golden_backward.py`` -> ``backward.npz``, keys ``sup_*``): proposals, matched
ground-truth ids, joints, both 3-D losses, the gradient of every heat-map and the gradient norm / sum of all 180
parameters.  Every kernel entry point is replaced by a torch / numpy emulation of its documented semantics
(include/sp3d.h), so this pins the host-side logic; the kernels are checked on the GPU
(tests/test_gpu_backward.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import volume_ops
from selfpose3d_b200 import autograd as ag
from selfpose3d_b200 import ops, synthetic
from selfpose3d_b200.config import default_config
from selfpose3d_b200.models import multi_person_posenet
from test_autograd_cpu import emulated  # noqa: F401  (fixture: convolution / BatchNorm / pooling kernels emulated)


def emul_unproject(cams, centers, cube_sample, spec, *hms):
    """sp3d_unproject_fwd (include/sp3d.h; lib/models/project_layer.py:42-102) from the packed camera table, with
    torch ops so that autograd provides sp3d_unproject_bwd's result."""
    grid_size, cube_size, img_size, hm_cfg_wh, C, pitch = spec
    X, Y, Z = [int(v) for v in cube_size]
    N = X * Y * Z
    W, H = float(img_size[0]), float(img_size[1])
    wc, hc = float(hm_cfg_wh[0]), float(hm_cfg_wh[1])
    lin = [torch.linspace(-grid_size[a] / 2, grid_size[a] / 2, cube_size[a]) for a in range(3)]
    cubes = []
    for q in range(int(centers.shape[0])):
        b = int(cube_sample[q]) if cube_sample is not None else q
        gx, gy, gz = torch.meshgrid(*[lin[a] + centers[q, a] for a in range(3)], indexing="ij")
        grid = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)
        num, den = torch.zeros(C, N), torch.zeros(N)
        for v in range(len(hms)):
            cam = cams[b, v]
            R, T, k, p, A = cam[0:9].view(3, 3), cam[9:12], cam[16:19], cam[19:21], cam[21:27].view(2, 3)
            xcam = torch.mm(R, grid.t() - T[:, None])
            y = xcam[:2] / (xcam[2] + 1e-5)
            r2 = torch.clamp((y ** 2).sum(0), max=1e10)
            corr = 1 + k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3 + 2 * (p[0] * y[1] + p[1] * y[0])
            px = cam[12] * (y[0] * corr + p[1] * r2) + cam[14]
            py = cam[13] * (y[1] * corr + p[0] * r2) + cam[15]
            width, height = float(cam[27]), float(cam[28])
            m = ((px >= 0) & (py >= 0) & (px < width) & (py < height)).float()
            px, py = px.clamp(-1.0, max(width, height)), py.clamp(-1.0, max(width, height))
            qx = A[0, 0] * px + A[0, 1] * py + A[0, 2]
            qy = A[1, 0] * px + A[1, 1] * py + A[1, 2]
            if float(cam[29]) != 0:
                qx = W - qx
            sx = (qx * wc / W / (wc - 1) * 2.0 - 1.0).clamp(-1.1, 1.1)
            sy = (qy * hc / H / (hc - 1) * 2.0 - 1.0).clamp(-1.1, 1.1)
            s = F.grid_sample(hms[v][b:b + 1], torch.stack([sx, sy], dim=1).view(1, 1, N, 2), align_corners=True)[0, :, 0]
            num, den = num + s * m[None], den + m
        o = num / (den + 1e-6)[None]
        o = torch.where(o != o, torch.zeros_like(o), o).clamp(0.0, 1.0)
        cubes.append(F.pad(o.t().reshape(X, Y, Z, C), (0, pitch - C)))
    return torch.stack(cubes)


def emul_softargmax(y, centers, spec):
    """sp3d_softargmax3d_fwd: sum_v softmax(beta x)_v * fl(lin + centre) (lib/models/pose_regression_net.py:19-28)."""
    C, cube_size, grid_size, beta = spec
    n = int(y.shape[0])
    lin = [torch.linspace(-grid_size[a] / 2, grid_size[a] / 2, cube_size[a]) for a in range(3)]
    out = []
    for q in range(n):
        gx, gy, gz = torch.meshgrid(*[lin[a] + centers[q, a] for a in range(3)], indexing="ij")
        grid = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)            # [N, 3]
        p = F.softmax(beta * y[q, ..., :C].reshape(-1, C).t(), dim=1)                           # [C, N]
        out.append(p @ grid)
    return torch.stack(out)


def emul_nms_topk(root_cubes, max_people, threshold, space_size, space_center, loc_f64=False, return_index=False):
    gc = volume_ops.proposal_layer(root_cubes.detach().numpy(), space_size, space_center, list(root_cubes.shape[1:]),
                                   int(max_people), float(threshold), f64=loc_f64)
    return torch.from_numpy(np.asarray(gc, dtype=np.float32))


class _Apply:
    def __init__(self, fn):
        self.apply = staticmethod(fn).__func__


def test_supervised_training_step_matches_reference(emulated, monkeypatch, golden):  # noqa: F811
    monkeypatch.setattr(ag, "Unproject", _Apply(emul_unproject))
    monkeypatch.setattr(ag, "SoftArgmax", _Apply(emul_softargmax))
    monkeypatch.setattr(ops, "nms_topk", emul_nms_topk)
    g, gb = golden("inference_small"), golden("backward")
    cfg = default_config()
    J = int(g["num_joints"])
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [int(v) for v in g["image_size"]], [int(v) for v in g["heatmap_size"]]
    cfg.NETWORK.NUM_JOINTS = J
    cfg.DATASET.ROOTIDX = cfg.DATASET.ROOTIDX_PSEUDO = 2
    cfg.NETWORK.ROOTNET_ROOTHM, cfg.NETWORK.USE_GT, cfg.NETWORK.TRAIN_ONLY_2D, cfg.NETWORK.BETA = True, False, False, 100.0
    cfg.MULTI_PERSON.SPACE_SIZE = [float(v) for v in g["space_size"]]
    cfg.MULTI_PERSON.SPACE_CENTER = [float(v) for v in g["space_center"]]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [int(v) for v in g["initial_cube_size"]]
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = int(g["max_people"])
    cfg.MULTI_PERSON.THRESHOLD = float(g["threshold"])
    cfg.PICT_STRUCT.GRID_SIZE = [float(v) for v in g["grid_size"]]
    cfg.PICT_STRUCT.CUBE_SIZE = [int(v) for v in g["cube_size"]]
    cfg.BACKBONE_MODEL = ""
    model = multi_person_posenet.get_multi_person_pose_net(cfg, is_train=True)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=int(gb["sup_seed"])), strict=True)
    model.train()
    V = g["heatmaps"].shape[0]
    meta = [{"center": torch.from_numpy(g["center"][v]), "scale": torch.from_numpy(g["scale"][v]),
             "rotation": torch.from_numpy(g["rotation"][v]),
             "camera": {k[4:]: torch.from_numpy(g[k][v]) for k in g if k.startswith("cam_")}} for v in range(V)]
    meta[0].update(roots_3d=torch.from_numpy(gb["sup_roots_3d"]), num_person=torch.from_numpy(gb["sup_num_person"]),
                   joints_3d=torch.from_numpy(gb["sup_joints_3d"]), joints_3d_vis=torch.from_numpy(gb["sup_joints_3d_vis"]))
    hms = [torch.from_numpy(g["heatmaps"][v]).clone().requires_grad_(True) for v in range(V)]
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)     # the module API moves inputs to the GPU
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    pred, _, gc, loss_2d, loss_3d, loss_cord = model(views=None, meta=meta, targets_3d=torch.from_numpy(gb["sup_targets_3d"]),
                                                     input_heatmaps=hms)
    (loss_3d + loss_cord).backward()

    np.testing.assert_allclose(gc.detach().numpy()[..., :3], gb["sup_grid_centers"][..., :3], rtol=1e-5, atol=1e-3)
    assert np.array_equal(gc.detach().numpy()[..., 3], gb["sup_grid_centers"][..., 3])          # matched ground-truth ids
    np.testing.assert_allclose(gc.detach().numpy()[..., 4], gb["sup_grid_centers"][..., 4], rtol=0, atol=1e-5)
    assert float(loss_2d) == 0.0
    assert abs(float(loss_3d) - float(gb["sup_loss_3d"])) <= 1e-5 * float(gb["sup_loss_3d"])
    assert abs(float(loss_cord) - float(gb["sup_loss_cord"])) <= 1e-4 * float(gb["sup_loss_cord"])
    # (the emulated convolutions accumulate in float64, the reference in float32: beta = 100 turns that into ~0.1 mm)
    np.testing.assert_allclose(pred.detach().numpy(), gb["sup_pred"], rtol=0, atol=0.3)          # mm
    gh = np.stack([h.grad.numpy() for h in hms])
    want = gb["sup_grad_heatmaps"]
    assert np.abs(gh - want).max() <= 2e-3 * np.abs(want).max(), np.abs(gh - want).max() / np.abs(want).max()
    params = dict(model.named_parameters())
    top = float(gb["sup_param_grad_norm"].max())
    checked = 0
    for name, norm, tot in zip(gb["sup_param_names"], gb["sup_param_grad_norm"], gb["sup_param_grad_sum"]):
        p = params[str(name)]
        gn = 0.0 if p.grad is None else float(p.grad.double().norm())
        if norm < 1e-5 * top:      # convolution biases in front of a batch normalisation: cancelled, rounding noise only
            assert gn < 1e-4 * top, (name, gn, norm)
            continue
        if str(name) == "pose_net.v2v_net.output_layer.bias":
            # the soft-argmax is invariant to a per-channel shift of its input: this gradient is exactly zero and what
            # either side reports is float32 noise of sum_v dx_v (beta = 100)
            assert gn < 1e-3 * top, (name, gn, norm)
            continue
        assert abs(gn - norm) <= 5e-3 * norm, (name, gn, norm)
        gs = float(p.grad.double().sum())
        assert abs(gs - tot) <= 5e-3 * norm * max(1.0, np.sqrt(p.numel())), (name, gs, tot)
        checked += 1
    assert checked >= 60, checked


# ---------------------------------------------------------------------------------------------- SSL step
def emul_unproject_op(heatmaps, hm_strides, cams, centers, grid_size, cube_size, image_size, heatmap_hw, channels, out,
                      out_strides, out_c_pad=0, check_flag=False, cubes_per_sample=1, cube_sample=None, grids=None,
                      view_range=None, partial=False, heatmap_cfg_wh=None, fast=False):
    """ops.unproject as the inference path calls it (channel-last float32 output, all views)."""
    assert not partial and not fast and grids is None and view_range is None and cubes_per_sample == 1 and not check_flag
    pitch = int(out.shape[-1])
    spec = ([float(v) for v in grid_size], [int(v) for v in cube_size], image_size,
            heatmap_cfg_wh if heatmap_cfg_wh is not None else (heatmap_hw[1], heatmap_hw[0]), int(channels), pitch)
    with torch.no_grad():
        out.copy_(emul_unproject(cams, centers, cube_sample, spec, *[h.float() for h in heatmaps]))


def test_ssl_training_step_matches_reference(emulated, monkeypatch, golden):  # noqa: F811
    """``MultiPersonPoseNetSSV.forward(inference=False)``: one self-supervised step (ResNet-50 backbone and ResNet-18
    attention net in .train(), frozen root net, pose net in .train(); re-projection, Gaussian rendering, attention and
    Hungarian-L1 losses) against the step recorded from the unmodified reference
    (tests/golden/make_golden_ssl.py -> ssl_step.npz): the four losses, the joints, the proposals and the gradient
    norms of all 331 trained parameters."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden_ssl as gen
    from test_autograd_cpu import _conv_wgrad_any, _maxpool_any, _maxpool_bwd_any
    from selfpose3d_b200 import grad_ops
    from selfpose3d_b200.models import multi_person_posenet_ssv
    monkeypatch.setattr(ag, "Unproject", _Apply(emul_unproject))
    monkeypatch.setattr(ag, "SoftArgmax", _Apply(emul_softargmax))
    monkeypatch.setattr(ag, "RenderGaussians", _Apply(emul_render))       # sp3d_gauss_render_fwd (+ autograd for _bwd)
    monkeypatch.setattr(ops, "nms_topk", emul_nms_topk)
    monkeypatch.setattr(ops, "unproject", emul_unproject_op)
    monkeypatch.setattr(ops, "maxpool", _maxpool_any)
    monkeypatch.setattr(grad_ops, "maxpool_bwd", _maxpool_bwd_any)
    monkeypatch.setattr(grad_ops, "conv_wgrad", _conv_wgrad_any)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    gs = golden("ssl_step")
    cfg = gen.configure(default_config())
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = list(gen.IMAGE), list(gen.HEATMAP)
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=int(gs["seed"])), strict=True)
    model.train()
    model.root_net.eval()
    (v1, m1, t1), (v2, m2, t2), (v3, m3, t3) = gen.ssl_case()
    pred, hm3, gc, losses = model(views1=v1, meta1=m1, targets_2d1=t1, views2=v2, meta2=m2, targets_2d2=t2,
                                  views3=v3, meta3=m3, targets_2d3=t3, inference=False, epoch=1)
    sum(losses.values()).backward()

    assert sorted(losses) == [str(n) for n in gs["loss_names"]]
    for name, want in zip(gs["loss_names"], gs["loss_values"]):
        assert abs(float(losses[str(name)]) - want) <= 2e-3 * abs(want), (name, float(losses[str(name)]), want)
    np.testing.assert_allclose(torch.stack(hm3).detach().numpy(), gs["heatmaps3"], rtol=0,
                               atol=1e-3 * np.abs(gs["heatmaps3"]).max())
    np.testing.assert_allclose(gc.detach().numpy(), gs["grid_centers"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(pred.numpy(), gs["pred"], rtol=0, atol=1.0)          # mm (float64-accumulating emulation)
    params = dict(model.named_parameters())
    top = float(gs["param_grad_norm"].max())
    checked = 0
    for name, norm in zip(gs["param_names"], gs["param_grad_norm"]):
        p = params[str(name)]
        if norm < 0:                                   # frozen root net: no gradient on either side
            assert p.grad is None or not p.grad.any(), name
            continue
        gn = float(p.grad.double().norm())
        if norm < 1e-5 * top or str(name) == "pose_net.v2v_net.output_layer.bias":   # cancelled: rounding noise only
            assert gn < 1e-3 * top, (name, gn, norm)
            continue
        assert abs(gn - norm) <= 5e-2 * norm, (name, gn, norm)
        checked += 1
    assert checked >= 150, checked


def test_synthetic_root_step_matches_reference(emulated, monkeypatch, golden):  # noqa: F811
    """``CuboidProposalNetSoft`` with ``ROOTNET_TRAIN_SYNTH`` in .train(): the random roots are drawn in the
    reference's order, so with the same torch seed the target volume, the synthetic volume, the real volume, the
    proposals and the gradients equal the step recorded from the reference (ssl_step.npz, keys ``syn_*``)."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_golden_ssl as gen
    from selfpose3d_b200.models import cuboid_proposal_net_soft
    monkeypatch.setattr(ag, "Unproject", _Apply(emul_unproject))
    monkeypatch.setattr(ops, "nms_topk", emul_nms_topk)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    gs = golden("ssl_step")
    cfg = gen.configure(default_config())
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = list(gen.IMAGE), list(gen.HEATMAP)
    cfg.NETWORK.ROOTNET_TRAIN_SYNTH = True
    net = cuboid_proposal_net_soft.CuboidProposalNetSoft(cfg)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(gs["syn_seed"])), strict=True)
    net.train()
    (_, meta, targets), _, _ = gen.ssl_case()
    torch.manual_seed(gen.SYNTH_TORCH_SEED)
    main, syn, target, gc = net(targets, meta, flip_xcoords=meta[0]["hflip"])
    (100.0 * F.mse_loss(syn, target)).backward()
    np.testing.assert_allclose(target.numpy(), gs["syn_target"], rtol=0, atol=1e-6)
    for got, key in ((main, "syn_main"), (syn, "syn_cubes")):
        np.testing.assert_allclose(got.detach().numpy(), gs[key], rtol=0, atol=1e-4 * np.abs(gs[key]).max(), err_msg=key)
    np.testing.assert_allclose(gc.detach().numpy(), gs["syn_grid_centers"], rtol=1e-4, atol=1e-3)
    params = dict(net.named_parameters())
    top = float(gs["syn_param_grad_norm"].max())
    checked = 0
    for name, norm in zip(gs["syn_param_names"], gs["syn_param_grad_norm"]):
        gn = float(params[str(name)].grad.double().norm())
        if norm < 1e-5 * top:
            assert gn < 1e-3 * top, (name, gn, norm)
            continue
        assert abs(gn - norm) <= 1e-2 * norm, (name, gn, norm)
        checked += 1
    assert checked >= 40, checked


def test_bench_training_step_case(emulated, monkeypatch):  # noqa: F811
    """bench.py's training-step side measurement at toy size with emulated kernels: the ground truth derived from the
    training-mode proposals matches every slot, and the step leaves gradients on root-net and pose-net parameters
    (none on the frozen backbone's)."""
    import bench
    from test_autograd_cpu import _maxpool_any
    monkeypatch.setattr(ag, "Unproject", _Apply(emul_unproject))
    monkeypatch.setattr(ag, "SoftArgmax", _Apply(emul_softargmax))
    monkeypatch.setattr(ops, "nms_topk", emul_nms_topk)
    monkeypatch.setattr(ops, "maxpool", _maxpool_any)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 3
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [64, 96], [16, 24]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE, cfg.MULTI_PERSON.MAX_PEOPLE_NUM, cfg.MULTI_PERSON.THRESHOLD = [8, 8, 4], 2, -1e9
    cfg.PICT_STRUCT.CUBE_SIZE = [8, 8, 8]
    model, step = bench.build_training_step(cfg, "cpu", [64, 96], 3)
    grid = step()
    assert int((grid[:, :, 3] >= 0).sum()) == 2
    got = {n.split(".")[0] for n, p in model.named_parameters() if p.grad is not None and bool(p.grad.any())}
    assert got == {"root_net", "pose_net"}, got


def emul_softargmax_op(x, strides, n_cubes, channels, cube_size, centers, grid_size, beta, check_flag=False, lin=None):
    """ops.softargmax as the inference path calls it (channel-last float32 volume ``[n, X, Y, Z, pitch]``)."""
    assert lin is None and not check_flag
    return emul_softargmax(x, centers, (int(channels), [int(v) for v in cube_size], [float(v) for v in grid_size], float(beta)))


def apply_emulation_in_this_process():
    """The same kernel emulations as the fixtures above, applied for good (spawned `gloo` workers of
    tests/test_dist_gloo.py, which have no pytest fixtures)."""
    import test_autograd_cpu as T
    from selfpose3d_b200 import grad_ops
    ops.set_float32_conv("simt")     # the emulated launcher takes the float32 FMA kernel's arguments
    ops.conv_launch, ops.to_channel_last, ops.to_channel_first = T.emulate_conv_launch, T._cl, T._cf
    ops.maxpool, ops.nms_topk = T._maxpool_any, emul_nms_topk
    ops.unproject, ops.softargmax = emul_unproject_op, emul_softargmax_op
    grad_ops._f32 = lambda *a: None
    for name, fn in (("maxpool_bwd", T._maxpool_bwd_any), ("bn_stats", T._bn_stats), ("bn_apply", T._bn_apply),
                     ("bn_bwd", T._bn_bwd), ("relu_bwd", T._relu_bwd), ("conv_wgrad", T._conv_wgrad_any)):
        setattr(grad_ops, name, fn)
    ag.Unproject, ag.SoftArgmax = _Apply(emul_unproject), _Apply(emul_softargmax)
    ag.RenderGaussians = _Apply(emul_render)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)


def emul_render(kps, n_people, hw, inv_scale, sigma):
    """sp3d_gauss_render_fwd (include/sp3d.h): clip(sum_{p < n[b]} exp(-((x - kx s)/sigma)^2/2 - ((y - ky s)/sigma)^2/2), 0, 1)."""
    V, B, P, J, _ = kps.shape
    yy, xx = torch.meshgrid(torch.arange(hw[0], dtype=torch.float32), torch.arange(hw[1], dtype=torch.float32), indexing="ij")
    k = kps * inv_scale
    g = torch.exp(-(((xx - k[..., 0, None, None]) / sigma) ** 2) / 2 - (((yy - k[..., 1, None, None]) / sigma) ** 2) / 2)
    mask = (torch.arange(P)[None, :] < n_people[:, None]).float()                       # [B, P]
    return torch.clip((g * mask[None, :, :, None, None, None]).sum(2), 0.0, 1.0)


def test_render_kernel_wrapper_equals_the_tensor_expression(monkeypatch):
    """The padded ``[V,B,P,J,2]`` wrapper around the rendering kernel (opt-in, SP3D_RENDER_KERNEL=1) against the
    list-based tensor expression of ``_ssl_train.render_gaussians``, values and gradients, ragged people counts."""
    from selfpose3d_b200.models import _ssl_train
    monkeypatch.setattr(ag, "RenderGaussians", _Apply(emul_render))
    torch.manual_seed(3)
    V, B, J, h, w = 3, 3, 4, 12, 9
    counts = [2, 0, 3]
    base = [[(torch.rand(n, J, 2) * torch.tensor([4.0 * w, 4.0 * h])).requires_grad_(True) for n in counts] for _ in range(V)]
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    xx, yy = xx.view(1, 1, h, w), yy.view(1, 1, h, w)
    G = torch.randn(V, B, J, h, w)
    want = _ssl_train.render_gaussians(base, xx, yy)
    (want * G).sum().backward()
    grads = [[kp.grad.clone() for kp in v] for v in base]
    for v in base:
        for kp in v:
            kp.grad = None
    monkeypatch.setenv("SP3D_RENDER_KERNEL", "1")
    got = _ssl_train.render_gaussians(base, xx, yy)
    (got * G).sum().backward()
    assert torch.allclose(got, want, atol=1e-6)
    for v in range(V):
        for b in range(B):
            if counts[b]:
                assert torch.allclose(base[v][b].grad, grads[v][b], rtol=1e-4, atol=1e-7)


def test_slot_batched_pose_net_equals_one_pass_per_slot(emulated, monkeypatch):  # noqa: F811
    """``PoseRegressionNet.regress_slots`` (all proposal slots in one pass, grouped BatchNorm statistics) against the
    reference's structure -- one training forward per slot on that slot's valid rows: joints, heat-map gradients,
    parameter gradients and the running statistics after the step."""
    import copy
    from selfpose3d_b200 import autograd as ag
    from selfpose3d_b200.models import pose_regression_net
    monkeypatch.setattr(ag.Unproject, "apply", staticmethod(emul_unproject))
    monkeypatch.setattr(ag.SoftArgmax, "apply", staticmethod(emul_softargmax))
    ops.set_volume_dtype(torch.float32)
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 3
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [64, 48], [16, 12]
    cfg.PICT_STRUCT.CUBE_SIZE = [8, 8, 8]
    torch.manual_seed(3)
    net_a = pose_regression_net.PoseRegressionNet(cfg).train()
    net_b = copy.deepcopy(net_a)
    B, V, K = 3, 2, 3
    meta = synthetic.make_meta(synthetic.ring_cameras(V, seed=1), B, (64, 48))
    hms_a = [torch.rand(B, 3, 12, 16, requires_grad=True) for _ in range(V)]
    hms_b = [h.detach().clone().requires_grad_(True) for h in hms_a]
    gc = torch.zeros(B, K, 5)
    gc[..., :3] = torch.randn(B, K, 3) * 300
    gc[..., 3] = torch.tensor([[0, 1, -1], [0, -1, -1], [1, 0, -1]], dtype=torch.float32)     # slot 2: no valid row
    flags = gc[:, :, 3]
    # (a) one pass for all slots
    joints, slots, samples = net_a.regress_slots(hms_a, meta, gc, flags)
    assert slots == [0, 0, 0, 1, 1] and samples == [0, 1, 2, 0, 2]
    w = torch.randn_like(joints)
    (joints * w).sum().backward()
    # (b) the reference's loop
    out, o = [], 0
    for n in range(K):
        if bool((flags[:, n] >= 0).any()):
            single = net_b(hms_b, meta, gc[:, n])
            rows = [i for i in range(B) if flags[i, n] >= 0]
            out.append(single[rows])
    ref = torch.cat(out)
    (ref * w).sum().backward()
    assert float((joints - ref).detach().abs().max()) <= 1e-3 * max(1.0, float(ref.detach().abs().max()))
    for ha, hb in zip(hms_a, hms_b):
        assert float((ha.grad - hb.grad).abs().max()) <= 1e-4 * float(hb.grad.abs().max())
    # (a convolution bias in front of a batch-statistics BatchNorm has an exactly zero gradient: what both routes
    #  compute there is rounding noise, so the bar is relative to the largest gradient of the net)
    top = max(float(p.grad.abs().max()) for p in net_b.parameters())
    for (name, pa), (_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        scale = max(float(pb.grad.abs().max()), 1e-12)
        assert float((pa.grad - pb.grad).abs().max()) <= max(2e-4 * scale, 2e-5 * top), name
    for (name, ba), (_, bb) in zip(net_a.named_buffers(), net_b.named_buffers()):
        assert torch.allclose(ba.float(), bb.float(), rtol=1e-5, atol=1e-6), name
