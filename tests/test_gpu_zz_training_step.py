"""GPU: one supervised training step of ``MultiPersonPoseNet`` (root net + pose net in ``.train()``, heat-maps given)
through the kernels, against the step recorded from the unmodified reference (``backward.npz``, keys ``sup_*``).

The host-side logic of this step is pinned on CPU (tests/test_training_cpu.py, kernels emulated) and every kernel it
launches is checked on the GPU on its own (tests/test_gpu_backward.py).  These end-to-end compositions passed on the
B200 in the round-1 driver run (recorded as XPASS there); they are strict tests now."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

from selfpose3d_b200 import synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet  # noqa: E402

DEV = "cuda:0"
DEFAULT_F32_CONV = "simt"   # tests/conftest.py runs this module on the float32 FMA kernels


def test_supervised_training_step_matches_reference_on_gpu(golden):
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
    model = model.to(DEV).train()
    V = g["heatmaps"].shape[0]
    meta = [{"center": torch.from_numpy(g["center"][v]), "scale": torch.from_numpy(g["scale"][v]),
             "rotation": torch.from_numpy(g["rotation"][v]),
             "camera": {k[4:]: torch.from_numpy(g[k][v]) for k in g if k.startswith("cam_")}} for v in range(V)]
    meta[0].update(roots_3d=torch.from_numpy(gb["sup_roots_3d"]), num_person=torch.from_numpy(gb["sup_num_person"]),
                   joints_3d=torch.from_numpy(gb["sup_joints_3d"]), joints_3d_vis=torch.from_numpy(gb["sup_joints_3d_vis"]))
    hms = [torch.from_numpy(g["heatmaps"][v]).to(DEV).requires_grad_(True) for v in range(V)]
    pred, _, gc, loss_2d, loss_3d, loss_cord = model(views=None, meta=meta, targets_3d=torch.from_numpy(gb["sup_targets_3d"]),
                                                     input_heatmaps=hms)
    (loss_3d + loss_cord).backward()
    gc = gc.detach().cpu().numpy()
    np.testing.assert_allclose(gc[..., :3], gb["sup_grid_centers"][..., :3], rtol=1e-5, atol=1e-3)
    assert np.array_equal(gc[..., 3], gb["sup_grid_centers"][..., 3])
    np.testing.assert_allclose(gc[..., 4], gb["sup_grid_centers"][..., 4], rtol=0, atol=1e-5)
    assert abs(float(loss_3d) - float(gb["sup_loss_3d"])) <= 1e-5 * float(gb["sup_loss_3d"])
    assert abs(float(loss_cord) - float(gb["sup_loss_cord"])) <= 1e-4 * float(gb["sup_loss_cord"])
    np.testing.assert_allclose(pred.detach().cpu().numpy(), gb["sup_pred"], rtol=0, atol=0.3)
    gh = np.stack([h.grad.cpu().numpy() for h in hms])
    want = gb["sup_grad_heatmaps"]
    assert np.abs(gh - want).max() <= 2e-3 * np.abs(want).max()
    params = dict(model.named_parameters())
    top = float(gb["sup_param_grad_norm"].max())
    for name, norm in zip(gb["sup_param_names"], gb["sup_param_grad_norm"]):
        p = params[str(name)]
        gn = 0.0 if p.grad is None else float(p.grad.double().norm())
        if norm < 1e-5 * top or str(name) == "pose_net.v2v_net.output_layer.bias":   # cancelled: rounding noise only
            assert gn < 1e-3 * top, (name, gn, norm)
            continue
        assert abs(gn - norm) <= 5e-3 * norm, (name, gn, norm)


@pytest.mark.parametrize("case", ["3x3s2_even", "3x3s2_odd", "1x1s2", "7x7s2", "deconv4s2"])
def test_strided_conv_dgrad_vs_autograd(case):
    """Input gradients of strided / transposed 2-D convolutions (transposed convolution with an explicit output
    extent and tap-less phases; strided convolution) -- CPU twin: tests/test_conv_lowering_cpu.py."""
    import torch.nn as nn
    from selfpose3d_b200 import grad_ops, ops
    torch.manual_seed(21)
    transposed = False
    if case.startswith("3x3s2"):
        conv, x = nn.Conv2d(64, 128, 3, 2, 1), torch.randn(2, 64, *((10, 8) if case.endswith("even") else (9, 7)))
    elif case == "1x1s2":
        conv, x = nn.Conv2d(64, 128, 1, 2, 0), torch.randn(2, 64, 9, 8)
    elif case == "7x7s2":
        conv, x = nn.Conv2d(3, 64, 7, 2, 3), torch.randn(1, 3, 14, 12)
    else:
        conv, x, transposed = nn.ConvTranspose2d(256, 64, 4, 2, 1), torch.randn(2, 256, 5, 4), True
    conv = conv.double()
    x = x.double().requires_grad_(True)
    y = conv(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    pc = ops.PackedConv(conv.weight.detach().float().to(DEV), conv.bias.detach().float().to(DEV), None, conv.stride[0],
                        conv.padding[0], transposed=transposed, relu=0)
    xcl = ops.to_channel_last(x.detach().float().unsqueeze(2).to(DEV))
    gycl = ops.to_channel_last(gy.float().unsqueeze(2).to(DEV))
    gx = grad_ops.conv_dgrad(pc, gycl, out_pitch=int(xcl.shape[-1]), in_dims=tuple(xcl.shape[1:4]))
    got = ops.to_channel_first(gx, x.shape[1])[:, :, 0].cpu().double()
    assert float((got - x.grad).abs().max()) <= 2e-5 * float(x.grad.abs().max())


def test_pose_resnet_training_step_vs_oracle():
    """PoseResNet-50 in .train() on a 64 x 96 image batch through the kernels: output and the gradients of all
    parameters against float64 autograd through the oracle's restatement (L2 norm: single ReLU gates flip between
    float32 and float64 on feature maps this small) -- CPU twin: tests/test_autograd_cpu.py."""
    from oracle import nets
    from selfpose3d_b200.models import pose_resnet
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 5
    net = pose_resnet.get_pose_net(cfg, is_train=False)
    sd0 = synthetic.trained_like_state_dict(net, seed=90)
    net.load_state_dict(sd0, strict=True)
    torch.manual_seed(2)
    x = torch.randn(2, 3, 64, 96)
    gy = torch.randn(2, 5, 16, 24)

    def oracle(dtype):
        sd_ = {k: (v.to(dtype).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
               for k, v in sd0.items()}
        yo_ = nets.pose_resnet_forward(x.to(dtype), sd_, dtype=dtype, training=True)
        (yo_ * gy.to(dtype)).sum().backward()
        return yo_.detach(), sd_

    yo, sd = oracle(torch.float64)
    _, sd32 = oracle(torch.float32)
    net = net.to(DEV).train()
    y = net(x.to(DEV))
    (y * gy.to(DEV)).sum().backward()
    assert float((y.detach().cpu().double() - yo).abs().max()) <= 1e-3 * float(yo.abs().max())
    top = max(float(v.grad.abs().max()) for v in sd.values() if v.is_floating_point() and v.grad is not None)
    checked = 0
    for name, p in net.named_parameters():
        ref = sd[name].grad
        if float(ref.abs().max()) < 1e-7 * top:
            continue
        l2 = float((p.grad.cpu().double() - ref).norm() / ref.norm())
        assert l2 < max(3e-2, 3 * float((sd32[name].grad.double() - ref).norm() / ref.norm())), (name, l2)
        checked += 1
    assert checked >= 100


def test_ssl_training_step_matches_reference_on_gpu(golden):
    """``MultiPersonPoseNetSSV.forward(inference=False)`` through the kernels against the self-supervised step recorded
    from the unmodified reference (ssl_step.npz) -- CPU twin: tests/test_training_cpu.py."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_ssl as gen
    from selfpose3d_b200.models import multi_person_posenet_ssv
    gs = golden("ssl_step")
    cfg = gen.configure(default_config())
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = list(gen.IMAGE), list(gen.HEATMAP)
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=int(gs["seed"])), strict=True)
    model = model.to(DEV).train()
    model.root_net.eval()
    sets = [([v.to(DEV) for v in views], meta, [t.to(DEV) for t in targets]) for views, meta, targets in gen.ssl_case()]
    (v1, m1, t1), (v2, m2, t2), (v3, m3, t3) = sets
    pred, hm3, gc, losses = model(views1=v1, meta1=m1, targets_2d1=t1, views2=v2, meta2=m2, targets_2d2=t2,
                                  views3=v3, meta3=m3, targets_2d3=t3, inference=False, epoch=1)
    sum(losses.values()).backward()
    assert sorted(losses) == [str(n) for n in gs["loss_names"]]
    for name, want in zip(gs["loss_names"], gs["loss_values"]):
        assert abs(float(losses[str(name)]) - want) <= 2e-3 * abs(want), (name, float(losses[str(name)]), want)
    np.testing.assert_allclose(gc.detach().cpu().numpy(), gs["grid_centers"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(pred.cpu().numpy(), gs["pred"], rtol=0, atol=1.0)
    params = dict(model.named_parameters())
    top = float(gs["param_grad_norm"].max())
    for name, norm in zip(gs["param_names"], gs["param_grad_norm"]):
        p = params[str(name)]
        if norm < 0:
            assert p.grad is None or not p.grad.any(), name
            continue
        gn = float(p.grad.double().norm())
        if norm < 1e-5 * top or str(name) == "pose_net.v2v_net.output_layer.bias":
            assert gn < 1e-3 * top, (name, gn, norm)
            continue
        assert abs(gn - norm) <= 5e-2 * norm, (name, gn, norm)


def test_v2v_training_step_on_tensor_cores_matches_reference(golden):
    """The V2VNet training step with ``ops.set_float32_conv("bf16x3")``: forward convolutions and the input gradients
    whose adjoint shape has a compiled instantiation run on tcgen05 through split operands (weight gradients stay on
    the float32 FMA kernel), against the step recorded from the reference (same bounds as the float32 path, x2)."""
    from selfpose3d_b200 import ops
    from selfpose3d_b200.models import v2v_net
    gb = golden("backward")
    net = v2v_net.V2VNet(3, 3)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(gb["v2v_seed"])), strict=True)
    net = net.to(DEV).train()
    x = torch.from_numpy(gb["v2v_x"]).to(DEV).requires_grad_(True)
    ops.set_float32_conv("bf16x3")
    try:
        y = net(x)
        (y * torch.from_numpy(gb["v2v_grad_y"]).to(DEV)).sum().backward()
    finally:
        ops.set_float32_conv(DEFAULT_F32_CONV)
    scale = float(np.abs(gb["v2v_y"]).max())
    assert float(np.abs(y.detach().cpu().numpy() - gb["v2v_y"]).max()) <= 2e-4 * scale
    assert float(np.abs(x.grad.cpu().numpy() - gb["v2v_grad_x"]).max()) <= 2e-3 * float(np.abs(gb["v2v_grad_x"]).max())
    params = dict(net.named_parameters())
    for name, norm in zip(gb["v2v_param_names"], gb["v2v_param_grad_norm"]):
        g = params[str(name)].grad.double()
        assert abs(float(g.norm()) - norm) <= 4e-3 * max(norm, 1e-3), (name, float(g.norm()), norm)


def test_gauss_render_kernels_vs_autograd():
    """sp3d_gauss_render_fwd / bwd against the tensor expression of the reference (:410-448) and its autograd gradient:
    ragged people counts, joints inside and outside the map, sums above 1 (the clip gate)."""
    from selfpose3d_b200 import autograd as ag
    torch.manual_seed(5)
    V, B, P, J, h, w = 3, 2, 4, 15, 24, 18
    counts = torch.tensor([4, 2], dtype=torch.int32)
    kps = (torch.rand(V, B, P, J, 2, dtype=torch.float64) * torch.tensor([4.0 * w, 4.0 * h]) * 1.2 - 8.0)
    kps[:, 0, 1] = kps[:, 0, 0] + 0.5                       # two people on top of each other: the sum exceeds 1
    kps.requires_grad_(True)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    k = kps / 4.0
    g = torch.exp(-(((xx - k[..., 0, None, None]) / 3.0) ** 2) / 2 - (((yy - k[..., 1, None, None]) / 3.0) ** 2) / 2)
    mask = (torch.arange(P)[None, :] < counts[:, None]).double()
    want = torch.clip((g * mask[None, :, :, None, None, None]).sum(2), 0.0, 1.0)
    G = torch.randn(V, B, J, h, w, dtype=torch.float64)
    (want * G).sum().backward()
    kd = kps.detach().float().to(DEV).requires_grad_(True)
    got = ag.RenderGaussians.apply(kd, counts.to(DEV), (h, w), 0.25, 3.0)
    (got * G.float().to(DEV)).sum().backward()
    assert float((want >= 1).double().mean()) > 0.001
    assert float((got.cpu().double() - want).abs().max()) <= 1e-5
    ref = kps.grad * mask[None, :, :, None, None]
    assert float((kd.grad.cpu().double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


def test_synthetic_root_step_matches_reference_on_gpu(golden, monkeypatch):
    """``CuboidProposalNetSoft`` with ``ROOTNET_TRAIN_SYNTH`` in ``.train()`` (reference
    ``lib/models/cuboid_proposal_net_soft.py:151-241``) through the kernels: the random roots are drawn in the
    reference's order on the host RNG, so with the same torch seed the target volume, the synthetic and the real score
    volumes, the proposals and the parameter gradients equal the step recorded from the unmodified reference
    (ssl_step.npz, keys ``syn_*``) -- CPU twin: tests/test_training_cpu.py::test_synthetic_root_step_matches_reference."""
    import os
    import sys
    import torch.nn.functional as F
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_ssl as gen
    from selfpose3d_b200.models import cuboid_proposal_net_soft
    gs = golden("ssl_step")
    cfg = gen.configure(default_config())
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = list(gen.IMAGE), list(gen.HEATMAP)
    cfg.NETWORK.ROOTNET_TRAIN_SYNTH = True
    net = cuboid_proposal_net_soft.CuboidProposalNetSoft(cfg)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(gs["syn_seed"])), strict=True)
    net = net.to(DEV).train()
    (_, meta, targets), _, _ = gen.ssl_case()
    # the heat-map noise is drawn with randn_like on the maps' device; the recorded step drew it from the host
    # generator, so the test routes every randn_like through the host RNG (same seed -> same noise, same draw order)
    host_randn_like = torch.randn_like
    monkeypatch.setattr(torch, "randn_like", lambda t, **k: host_randn_like(t.cpu(), **k).to(t.device))
    torch.manual_seed(gen.SYNTH_TORCH_SEED)
    main, syn, target, gc = net([t.to(DEV) for t in targets], meta, flip_xcoords=meta[0]["hflip"])
    (100.0 * F.mse_loss(syn, target)).backward()
    np.testing.assert_allclose(target.cpu().numpy(), gs["syn_target"], rtol=0, atol=1e-6)
    for got, key in ((main, "syn_main"), (syn, "syn_cubes")):
        np.testing.assert_allclose(got.detach().cpu().numpy(), gs[key], rtol=0, atol=1e-4 * np.abs(gs[key]).max(), err_msg=key)
    np.testing.assert_allclose(gc.detach().cpu().numpy(), gs["syn_grid_centers"], rtol=1e-4, atol=1e-3)
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
