"""GPU tests of the backward operators (``include/sp3d.h`` "Backward operators", ``selfpose3d_b200/grad_ops.py``)
against the gradients recorded from the unmodified reference (``tests/golden/backward.npz``) and against CPU
autograd over the oracle's restatements (``oracle/backward.py``, torch.nn.functional on the host).

Tolerances: float32 scatter sums in a different order than ATen's -> 2e-5 of the gradient range; soft-argmax
(beta = 100 amplifies) 1e-4 of the range."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import cams_from_arrays, cam_arrays

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

from oracle import backward, geometry  # noqa: E402
from selfpose3d_b200 import grad_ops, ops, synthetic  # noqa: E402
from test_gpu_parity import meta_from_golden  # noqa: E402

DEV = "cuda:0"


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max()) / max(float(np.abs(want).max()), 1e-30)


def cl(x, pitch=None):       # [N,C,*sp] float32 CPU -> channel-last CUDA
    if x.dim() == 4:
        x = x.unsqueeze(2)
    return ops.to_channel_last(x.to(DEV).float(), c_pitch=pitch)


def cf(y, C, nd=3):          # channel-last CUDA -> [N,C,*sp] CPU
    out = ops.to_channel_first(y, C).cpu()
    return out[:, :, 0] if nd == 2 else out


# ------------------------------------------------------------------------------------------ un-projection
def _unproject_bwd_case(hm_np, meta, cam_arr, centers_np, flip, image_size, heatmap_size, grid_size, cube, grad_np,
                        channel_last):
    V, B, C, h, w = hm_np.shape
    hms = [torch.from_numpy(x).to(DEV) for x in hm_np]
    if channel_last:
        hms = [x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) for x in hms]
    grads = [torch.zeros_like(x) for x in hms]           # zeros_like keeps the (channel-last) strides
    assert all(g.stride() == x.stride() for g, x in zip(grads, hms))
    cams = ops.pack_cameras(meta, image_size, flip).to(DEV)
    cen = torch.from_numpy(centers_np).to(DEV)
    N = int(np.prod(cube))
    gcu = torch.from_numpy(grad_np.reshape(B, C, N)).to(DEV).contiguous()
    grad_ops.unproject_bwd(hms, hms[0].stride(), cams, cen, grid_size, cube, image_size, (h, w), C, gcu, (C * N, N, 1),
                           grads, check_flag=centers_np.shape[1] > 3)
    return np.stack([g.cpu().numpy() for g in grads])


@pytest.mark.parametrize("channel_last", [False, True])
def test_unproject_bwd_matches_reference_gradient(golden, channel_last):
    g, gb = golden("project_layer_pose"), golden("backward")
    # voxels that project within 1e-2 px of an image border may take the other side of the in-image test
    _, _, margin = geometry.unproject(g["heatmaps"], cams_from_arrays(g), g["center"], g["scale"], g["rotation"],
                                      g["image_size"], g["heatmap_size"], g["grid_size"], g["grid_center"], g["cube_size"],
                                      flip=g.get("flip"), return_aux=True)
    amb = margin < 1e-2
    grad = gb["pl_grad_cubes"].reshape(g["cubes"].shape[0], g["cubes"].shape[1], -1) * ~amb[:, None]
    got = _unproject_bwd_case(g["heatmaps"], meta_from_golden(g), cam_arrays(g), g["grid_center"], g["flip"],
                              [int(v) for v in g["image_size"]], [int(v) for v in g["heatmap_size"]],
                              [float(v) for v in g["grid_size"]], [int(v) for v in g["cube_size"]], grad, channel_last)
    want, _ = backward.unproject_grad([torch.from_numpy(x) for x in g["heatmaps"]], cam_arrays(g), g["center"], g["scale"],
                                      g["rotation"], g["image_size"], g["heatmap_size"], g["grid_size"], g["grid_center"],
                                      g["cube_size"], torch.from_numpy(grad.reshape(g["cubes"].shape)), flip=g.get("flip"))
    want = np.stack([t.numpy() for t in want])
    assert not got[:, 1].any()                                    # invalid proposal row: no gradient
    assert rel_err(got, want) <= 2e-5, rel_err(got, want)
    if not amb.any():                                             # then the recorded reference gradient applies as is
        assert rel_err(got, gb["pl_grad_heatmaps"]) <= 2e-5, rel_err(got, gb["pl_grad_heatmaps"])


def test_unproject_bwd_seeded_many_channels_vs_oracle():
    """C = 17 (two register groups), 5 views, rotation + scale + flip, values above 1 so that the clamp gate is hit."""
    cams = synthetic.ring_cameras(5, seed=7)
    B, C = 2, 17
    meta = synthetic.make_meta(cams, B, (96, 128), rotation=[[5.0, -12.0]] * 5, scale_mul=[[1.1, 0.9]] * 5)
    rs = np.random.RandomState(3)
    # per-channel offsets: a third of the channels average below 0, a third above 1 (both clamp gates close)
    offset = np.array([-0.5, 0.2, 0.9], dtype=np.float32)[np.arange(C) % 3]
    hm_np = (rs.rand(5, B, C, 32, 24).astype(np.float32) * 0.6 + offset[None, None, :, None, None])
    centers = np.array([[200.0, -700.0, 900.0, 0.0, 1.0], [-900.0, 300.0, 1000.0, 2.0, 1.0]], dtype=np.float32)
    flip = np.array([True, False])
    cube = [12, 8, 16]
    cam_arr = {k: np.stack([m["camera"][k].numpy() for m in meta]) for k in meta[0]["camera"]}
    cen_l, sc_l, rot_l = ([m[k].numpy() for m in meta] for k in ("center", "scale", "rotation"))
    cams_nested = [[{k: np.asarray(v[i]) for k, v in m["camera"].items()} for i in range(B)] for m in meta]
    o_cubes, _, margin = geometry.unproject(hm_np, cams_nested, cen_l, sc_l, rot_l, (96, 128), (24, 32), [2000.0] * 3,
                                            centers, cube, flip=flip, return_aux=True)
    # keep voxels away from image borders and from the clamp edges (gate decisions must not hinge on the last ulp)
    oc = o_cubes.reshape(B, C, -1)      # exactly 0 / 1 = clamped (gate closed on both sides); near 0 / 1 = undecidable
    pre_ok = (oc == 0) | (oc == 1) | ((np.abs(oc) > 1e-4) & (np.abs(oc - 1) > 1e-4))
    grad = rs.randn(B, C, int(np.prod(cube))).astype(np.float32) * (margin >= 1e-2)[:, None] * pre_ok
    got = _unproject_bwd_case(hm_np, meta, cam_arr, centers, flip, [96, 128], [24, 32], [2000.0] * 3, cube, grad, True)
    want, cubes = backward.unproject_grad([torch.from_numpy(x) for x in hm_np], cam_arr, cen_l, sc_l, rot_l, (96, 128),
                                          (24, 32), [2000.0] * 3, centers, cube,
                                          torch.from_numpy(grad.reshape(B, C, *cube)), flip=flip)
    # the clamp really gates part of the gradient, on both sides
    assert float((cubes >= 1).float().mean()) > 0.05 and float(((cubes <= 0) & (torch.from_numpy(grad.reshape(B, C, *cube)) != 0)).float().mean()) > 0.05
    assert rel_err(got, np.stack([t.numpy() for t in want])) <= 2e-5


# ------------------------------------------------------------------------------------------ soft-argmax
@pytest.mark.parametrize("shape,C,pitch", [((6, 5, 4), 3, 4), ((16, 12, 20), 15, 16), ((32, 32, 32), 15, 16), ((12, 10, 8), 5, 20)])
def test_softargmax_bwd_vs_autograd(golden, shape, C, pitch):
    rs = np.random.RandomState(12)
    n = 2
    x = (rs.rand(n, C, *shape) * 0.2).astype(np.float32)
    x[0, 0, 2, 3, 1] = 0.9
    x[1, C - 1, shape[0] - 1, 0, 3] = 0.5
    cen = np.array([[10.0, -20.0, 30.0], [-700.0, 400.0, 900.0]], dtype=np.float32)
    size = [500.0, 400.0, 300.0]
    grids = torch.from_numpy(np.stack([geometry.compute_grid(size, c, shape) for c in cen]))
    go = rs.randn(n, C, 3).astype(np.float32)
    want, out_ref = backward.softargmax_grad(torch.from_numpy(x).double(), grids.double(), 100.0, torch.from_numpy(go).double())
    xcl = ops.to_channel_last(torch.from_numpy(x).to(DEV), c_pitch=pitch)
    N = int(np.prod(shape))
    strides = (N * pitch, 1, pitch)
    cen_t = torch.from_numpy(cen).to(DEV)
    out = ops.softargmax(xcl, strides, n, C, shape, cen_t, size, 100.0)
    assert float((out.cpu().double() - out_ref).abs().max()) <= 1e-3
    gx = grad_ops.softargmax_bwd(xcl, strides, n, C, shape, cen_t, size, 100.0, out, torch.from_numpy(go).to(DEV))
    assert gx.shape == xcl.shape and not gx[..., C:].any()
    assert rel_err(cf(gx, C).numpy(), want.numpy()) <= 1e-4
    if shape == (6, 5, 4):     # the reference's recorded gradient has a non-separable random grid: checked through the
        gb = golden("backward")   # oracle on CPU (tests/test_oracle_golden.py); here only that shapes agree
        assert gb["sa_grad_x"].shape == x.shape


# ------------------------------------------------------------------------------------------ max pool
@pytest.mark.parametrize("nd,k,s,p,shape", [(3, 2, 2, 0, (8, 12, 16)), (3, 2, 2, 0, (6, 10, 4)), (2, 3, 2, 1, (17, 12)),
                                            (2, 3, 2, 1, (16, 11))])
def test_maxpool_bwd_vs_autograd(nd, k, s, p, shape):
    torch.manual_seed(nd * 10 + shape[0])
    C = 8
    x = torch.randn(2, C, *shape)
    x = torch.where(torch.rand_like(x) < 0.4, torch.zeros_like(x), x).clamp_min(0)     # ReLU-like: many tied zeros
    x.requires_grad_(True)
    y = F.max_pool3d(x, k, s, p) if nd == 3 else F.max_pool2d(x, k, s, p)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    kk, ss, pp = ([1] * (3 - nd) + [v] * nd for v in (k, s, p))
    pp = [0] * (3 - nd) + [p] * nd
    xcl = cl(x.detach())
    got = grad_ops.maxpool_bwd(xcl, C, kk, ss, pp, cl(gy))
    # pure routing; overlapping windows add up to 4 values in atomic order
    assert float((cf(got, C, nd) - x.grad).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------------------ convolutions
CONV_BWD_CASES = [  # nd, transposed, cin, cout, k, stride, pad, spatial
    (3, False, 15, 16, 7, 1, 3, (6, 9, 8)), (3, False, 16, 32, 3, 1, 1, (5, 8, 6)), (3, False, 32, 15, 1, 1, 0, (4, 6, 5)),
    (3, False, 1, 16, 7, 1, 3, (8, 8, 4)), (3, True, 64, 32, 2, 2, 0, (3, 4, 5)), (3, False, 128, 128, 3, 1, 1, (3, 4, 4)),
    (2, False, 64, 256, 1, 1, 0, (9, 7)), (2, False, 64, 64, 3, 1, 1, (9, 7)), (2, False, 3, 64, 7, 2, 3, (20, 14)),
    (2, False, 128, 128, 3, 2, 1, (10, 8)), (2, True, 256, 64, 4, 2, 1, (5, 4)),
]


@pytest.mark.parametrize("nd,transposed,cin,cout,k,stride,pad,sp", CONV_BWD_CASES)
def test_conv_wgrad_and_dgrad_vs_autograd(nd, transposed, cin, cout, k, stride, pad, sp):
    torch.manual_seed(cin * 7 + cout + k)
    Conv = {(3, False): nn.Conv3d, (3, True): nn.ConvTranspose3d, (2, False): nn.Conv2d, (2, True): nn.ConvTranspose2d}[
        (nd, transposed)]
    conv = Conv(cin, cout, k, stride, pad).double()
    x = torch.randn(2, cin, *sp, dtype=torch.float64, requires_grad=True)
    y = conv(x)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    cw = conv.weight.detach().float().to(DEV)
    pc = ops.PackedConv(cw, conv.bias.detach().float().to(DEV), None, stride, pad, transposed=transposed, relu=0)
    xcl, gycl = cl(x.detach().float()), cl(gy.float())
    # forward through the same geometry first (so a wrong gradient cannot hide behind a wrong forward)
    assert rel_err(cf(pc(xcl), cout, nd).numpy(), y.detach().numpy()) <= 2e-5
    gw, gbias = grad_ops.conv_wgrad(pc, xcl, gycl)
    assert gw.shape == conv.weight.shape
    assert rel_err(gw.cpu().numpy(), conv.weight.grad.numpy()) <= 2e-5, rel_err(gw.cpu().numpy(), conv.weight.grad.numpy())
    assert rel_err(gbias.cpu().numpy(), conv.bias.grad.numpy()) <= 2e-5
    same = (not transposed and stride == 1 and 2 * pad == k - 1) or (transposed and k == stride and pad == 0)
    if same:   # (the strided forms are exercised in tests/test_gpu_zz_training_step.py)
        gx = grad_ops.conv_dgrad(pc, gycl)
        assert rel_err(cf(gx, cin, nd).numpy(), x.grad.numpy()) <= 2e-5


# ------------------------------------------------------------------------------------------ BatchNorm (training)
@pytest.mark.parametrize("C,shape,with_relu", [(16, (2, 6, 5, 4), True), (37, (3, 4, 9, 7), False), (128, (1, 3, 5, 2), True)])
def test_batchnorm_training_ops_vs_autograd(C, shape, with_relu):
    torch.manual_seed(C)
    n, sp = shape[0], shape[1:]
    x = (torch.randn(n, C, *sp, dtype=torch.float64) * 2 + 0.7).requires_grad_(True)
    gamma = (torch.rand(C, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, dtype=torch.float64) * 0.2).requires_grad_(True)
    y = F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-5)
    if with_relu:
        y = F.relu(y)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    xcl = cl(x.detach().float())
    mean, var = grad_ops.bn_stats(xcl, C)
    dims = (0, 2, 3, 4)
    assert rel_err(mean.cpu().numpy(), x.detach().mean(dims).numpy()) <= 1e-6
    assert rel_err(var.cpu().numpy(), x.detach().var(dims, unbiased=False).numpy()) <= 1e-5
    g32, b32 = gamma.detach().float().to(DEV), beta.detach().float().to(DEV)
    scale = g32 * torch.rsqrt(var + 1e-5)
    ycl = grad_ops.bn_apply(xcl, C, scale, b32 - mean * scale, relu=1 if with_relu else 0)
    assert rel_err(cf(ycl, C).numpy(), y.detach().numpy()) <= 1e-5
    if ycl.shape[-1] > C:
        assert not ycl[..., C:].any()
    gx, gg, gb = grad_ops.bn_bwd(xcl, C, cl(gy.float()), mean, var, g32, 1e-5, y=ycl if with_relu else None)
    assert rel_err(cf(gx, C).numpy(), x.grad.numpy()) <= 2e-5
    assert rel_err(gg.cpu().numpy(), gamma.grad.numpy()) <= 2e-5
    assert rel_err(gb.cpu().numpy(), beta.grad.numpy()) <= 2e-5


@pytest.mark.parametrize("C,shape,counts", [(16, (6, 6, 4), [1, 3, 2]), (32, (4, 4, 8), [2, 2]), (15, (3, 5, 8), [1, 1, 1, 1]),
                                            (128, (2, 2, 2), [3, 1]), (1, (4, 4, 4), [2, 1])])
def test_grouped_batchnorm_equals_one_call_per_group(C, shape, counts):
    """Grouped statistics (include/sp3d.h: item_group / group_items): ONE launch set over all items must give what one
    call per group gives -- statistics, outputs, input gradients -- and the parameter gradients summed over groups."""
    torch.manual_seed(C + len(counts))
    n = sum(counts)
    pitch = ops.round_up(C, 4)
    x = (torch.randn(n, *shape, pitch, device=DEV) * 2 + 0.5) * torch.linspace(0.5, 2.0, n, device=DEV).view(n, 1, 1, 1, 1)
    gy = torch.randn_like(x)
    res = torch.randn_like(x)
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    groups = grad_ops.BnGroups(counts, DEV)
    mean, var = grad_ops.bn_stats(x, C, groups=groups)
    assert mean.shape == (len(counts), C)
    scale = gamma * torch.rsqrt(var + 1e-5)
    y = grad_ops.bn_apply(x, C, scale.contiguous(), (beta - mean * scale).contiguous(), relu=1, residual=res, groups=groups)
    gx, gg, gb = grad_ops.bn_bwd(x, C, gy, mean, var, gamma, 1e-5, y=y, groups=groups)
    o, gg_ref, gb_ref = 0, 0, 0
    for g, c in enumerate(counts):
        sl = slice(o, o + c)
        o += c
        m1, v1 = grad_ops.bn_stats(x[sl], C)
        assert torch.allclose(m1, mean[g], rtol=1e-6, atol=1e-7) and torch.allclose(v1, var[g], rtol=1e-5, atol=1e-7)
        s1 = gamma * torch.rsqrt(v1 + 1e-5)
        y1 = grad_ops.bn_apply(x[sl], C, s1, beta - m1 * s1, relu=1, residual=res[sl])
        assert torch.allclose(y1, y[sl], rtol=1e-5, atol=1e-5)
        gx1, gg1, gb1 = grad_ops.bn_bwd(x[sl], C, gy[sl], m1, v1, gamma, 1e-5, y=y1)
        assert rel_err(gx[sl].cpu().numpy(), gx1.cpu().numpy()) <= 2e-5
        gg_ref, gb_ref = gg_ref + gg1, gb_ref + gb1
    assert rel_err(gg.cpu().numpy(), gg_ref.cpu().numpy()) <= 2e-5 and rel_err(gb.cpu().numpy(), gb_ref.cpu().numpy()) <= 2e-5
    if pitch > C:
        assert not y[..., C:].any() and not gx[..., C:].any()
    # running statistics updated by the statistics launch = len(counts) successive momentum updates, unbiased variance
    rm, rv = torch.randn(C, device=DEV), torch.rand(C, device=DEV) + 0.5
    want_m, want_v = rm.clone(), rv.clone()
    versions = (rm._version, rv._version)
    grad_ops.bn_stats(x, C, groups=groups, running=(rm, rv, 0.1))
    per_item = x[0].numel() // pitch
    for g, c in enumerate(counts):
        ng = c * per_item
        want_m.mul_(0.9).add_(mean[g], alpha=0.1)
        want_v.mul_(0.9).add_(var[g] * (ng / (ng - 1)), alpha=0.1)
    assert torch.allclose(rm, want_m, rtol=1e-6, atol=1e-7) and torch.allclose(rv, want_v, rtol=1e-6, atol=1e-7)
    assert rm._version > versions[0] and rv._version > versions[1]     # caches keyed on the buffers' versions notice


def test_basic3d_block_training_step_matches_reference_gradient(golden):
    """conv 3^3 -> batch-statistics BatchNorm -> ReLU (lib/models/v2v_net.py:10-20 in .train()) forward and backward
    composed from the operators, against the gradients recorded from the reference module."""
    gb = golden("backward")
    t = {k: torch.from_numpy(gb["b3_" + k]) for k in ("x", "w", "b", "gamma", "beta", "grad_y")}
    C = 8
    pc = ops.PackedConv(t["w"].to(DEV), t["b"].to(DEV), None, 1, 1, relu=0)
    xcl = cl(t["x"])
    z = pc(xcl)
    mean, var = grad_ops.bn_stats(z, C)
    gamma, beta = t["gamma"].to(DEV), t["beta"].to(DEV)
    scale = gamma * torch.rsqrt(var + 1e-5)
    y = grad_ops.bn_apply(z, C, scale, beta - mean * scale, relu=1)
    assert rel_err(cf(y, C).numpy(), gb["b3_y"]) <= 2e-5
    gz, gg, gbeta = grad_ops.bn_bwd(z, C, cl(t["grad_y"]), mean, var, gamma, 1e-5, y=y)
    gw, _ = grad_ops.conv_wgrad(pc, xcl, gz)
    gx = grad_ops.conv_dgrad(pc, gz)
    assert rel_err(gg.cpu().numpy(), gb["b3_grad_gamma"]) <= 5e-5
    assert rel_err(gbeta.cpu().numpy(), gb["b3_grad_beta"]) <= 5e-5
    assert rel_err(gw.cpu().numpy(), gb["b3_grad_w"]) <= 5e-5
    assert rel_err(cf(gx, 4).numpy(), gb["b3_grad_x"]) <= 5e-5


# ------------------------------------------------------------------------------------------ whole-net training step
def test_v2v_net_training_step_matches_reference(golden):
    """V2VNet(3, 3).train(): forward (batch-statistics BatchNorm), running-statistics update and the gradients of the
    input and of all 156 parameters through the kernels, against the step recorded from the reference module."""
    from selfpose3d_b200.models import v2v_net
    gb = golden("backward")
    net = v2v_net.V2VNet(3, 3)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=int(gb["v2v_seed"])), strict=True)
    net = net.to(DEV).train()
    x = torch.from_numpy(gb["v2v_x"]).to(DEV).requires_grad_(True)
    y = net(x)
    (y * torch.from_numpy(gb["v2v_grad_y"]).to(DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), gb["v2v_y"]) <= 1e-4
    assert rel_err(x.grad.cpu().numpy(), gb["v2v_grad_x"]) <= 1e-3
    params = dict(net.named_parameters())
    for name, norm, tot in zip(gb["v2v_param_names"], gb["v2v_param_grad_norm"], gb["v2v_param_grad_sum"]):
        g = params[str(name)].grad
        assert g is not None, name
        g = g.double()
        assert abs(float(g.norm()) - norm) <= 2e-3 * max(norm, 1e-3), (name, float(g.norm()), norm)
        assert abs(float(g.sum()) - tot) <= 2e-3 * max(norm, 1e-3) * max(1.0, np.sqrt(g.numel())), name
    bn = net.front_layers[0].block[1]
    assert rel_err(bn.running_mean.cpu().numpy(), gb["v2v_bn0_running_mean"]) <= 1e-5
    assert rel_err(bn.running_var.cpu().numpy(), gb["v2v_bn0_running_var"]) <= 1e-5
    # .eval() afterwards: the fused inference path (no graph), with the updated running statistics
    net.eval()
    with torch.no_grad():
        assert net(x.detach()).shape == y.shape


def test_v2v_net_training_step_pose_size_vs_oracle():
    """One training step of V2VNet(15, 15) on a 32^3 cube (the per-person net at half extent, batch of one).  At this
    size float32 itself is the limit: ReLU gates and max-pool winners that are decided in the last bits differ between
    float32 and float64 (the float32 CPU oracle is 1.5e-2 / 4.7e-2 of the range away from the float64 one on dL/dx /
    dL/dparams), so the kernels are held to the float32 oracle's own distance to float64, not to an absolute bound."""
    from oracle import nets
    from selfpose3d_b200.models import v2v_net
    net = v2v_net.V2VNet(15, 15)
    sd0 = synthetic.trained_like_state_dict(net, seed=61)
    net.load_state_dict(sd0, strict=True)
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.rand(1, 15, 32, 32, 32).astype(np.float32))
    gy = torch.from_numpy(rs.randn(1, 15, 32, 32, 32).astype(np.float32))

    def oracle(dtype):
        sd = {k: (v.to(dtype).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in sd0.items()}
        xo = x.to(dtype).clone().requires_grad_(True)
        yo = nets.v2v_forward(xo, sd, dtype=dtype, training=True)
        (yo * gy.to(dtype)).sum().backward()
        return yo.detach(), xo.grad, sd

    y64, gx64, sd64 = oracle(torch.float64)
    y32, gx32, sd32 = oracle(torch.float32)
    net = net.to(DEV).train()
    xin = x.to(DEV).requires_grad_(True)
    y = net(xin)
    (y * gy.to(DEV)).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), y64.numpy()) <= 1e-4
    e_x, r_x = rel_err(xin.grad.cpu().numpy(), gx64.numpy()), rel_err(gx32.numpy(), gx64.numpy())
    assert e_x <= max(1e-3, 3 * r_x), (e_x, r_x)
    worst, worst_ref = 0.0, 0.0
    for name, p in net.named_parameters():
        ref = sd64[name].grad
        if float(ref.abs().max()) < 1e-9 * float(gy.abs().max()):
            continue                                   # conv biases in front of a batch normalisation cancel exactly
        worst = max(worst, rel_err(p.grad.cpu().numpy(), ref.numpy()))
        worst_ref = max(worst_ref, rel_err(sd32[name].grad.numpy(), ref.numpy()))
    print("V2VNet(15) 32^3 training step: dL/dx %.3g (float32 oracle %.3g), dL/dparams %.3g (%.3g)" % (e_x, r_x, worst, worst_ref))
    assert worst <= max(2e-3, 3 * worst_ref), (worst, worst_ref)


def test_pose_regression_net_training_step_vs_oracle(golden):
    """PoseRegressionNet.train() on the reference's golden per-person case (4 views, rotation / scale augmentation,
    h-flip, one invalid row): un-project -> V2VNet (batch statistics) -> soft-argmax under autograd; the gradients of
    the heat-maps and of the V2VNet parameters against float64 autograd through the oracle chain
    (oracle.pipeline.unproject_torch -> oracle.nets.v2v_forward(training=True) -> oracle.backward.softargmax_forward),
    with the float32 oracle's own distance to float64 as the yardstick (beta = 100 amplifies rounding)."""
    from oracle import nets, pipeline
    from selfpose3d_b200.models import pose_regression_net
    from test_gpu_parity import cfg_for
    g = golden("project_layer_pose")
    cfg = cfg_for(g)
    J = int(g["heatmaps"].shape[2])
    cfg.NETWORK.NUM_JOINTS = J
    cfg.PICT_STRUCT.GRID_SIZE = [float(v) for v in g["grid_size"]]
    cfg.PICT_STRUCT.CUBE_SIZE = [int(v) for v in g["cube_size"]]
    net = pose_regression_net.PoseRegressionNet(cfg)
    sd0 = synthetic.trained_like_state_dict(net, seed=71)
    net.load_state_dict(sd0, strict=True)
    rs = np.random.RandomState(17)
    B = g["heatmaps"].shape[1]
    G = torch.from_numpy(rs.randn(B, J, 3).astype(np.float32))
    gc = torch.from_numpy(g["grid_center"])
    valid = gc[:, 3] >= 0

    def oracle(dtype):
        hms = [torch.from_numpy(h).to(dtype).requires_grad_(True) for h in g["heatmaps"]]
        sd = {k[len("v2v_net."):]: (v.to(dtype).clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
                                    else v.clone()) for k, v in sd0.items() if k.startswith("v2v_net.")}
        cubes, grids = pipeline.unproject_torch([h.float() for h in hms], cam_arrays(g),
                                                g["center"], g["scale"], g["rotation"], g["image_size"], g["heatmap_size"],
                                                g["grid_size"], g["grid_center"], g["cube_size"], flip=g.get("flip"))
        y = nets.v2v_forward(cubes[valid].to(dtype), sd, dtype=dtype, training=True)
        pred = backward.softargmax_forward(y, grids[valid].to(dtype), 100.0)
        (pred * G[valid].to(dtype)).sum().backward()
        return pred.detach(), [h.grad for h in hms], {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}

    # unproject_torch works in float32 internally; its float64 call still gives float64 V2V / soft-argmax arithmetic
    p64, gh64, gp64 = oracle(torch.float64)
    p32, gh32, gp32 = oracle(torch.float32)

    net = net.to(DEV).train()
    hms = [torch.from_numpy(h).to(DEV).requires_grad_(True) for h in g["heatmaps"]]
    pred = net(hms, meta_from_golden(g), gc.to(DEV), flip_xcoords=torch.from_numpy(g["flip"]))
    (pred * G.to(DEV)).sum().backward()
    assert not pred[~valid].any()
    e_pred = float((pred[valid].detach().cpu().double() - p64).abs().max())
    r_pred = float((p32.double() - p64).abs().max())
    assert e_pred <= max(2e-2, 5 * r_pred), (e_pred, r_pred)                      # mm
    gh = np.stack([h.grad.cpu().numpy() for h in hms])
    e_h = rel_err(gh, np.stack([t.numpy() for t in gh64]))
    r_h = rel_err(np.stack([t.numpy() for t in gh32]), np.stack([t.numpy() for t in gh64]))
    assert e_h <= max(5e-3, 5 * r_h), (e_h, r_h)
    worst, worst_ref = 0.0, 0.0
    params = dict(net.v2v_net.named_parameters())
    for name, ref in gp64.items():
        if float(ref.abs().max()) < 1e-9 * max(1.0, float(max(t.abs().max() for t in gp64.values()))):
            continue
        worst = max(worst, rel_err(params[name].grad.cpu().numpy(), ref.numpy()))
        worst_ref = max(worst_ref, rel_err(gp32[name].numpy(), ref.numpy()))
    print("pose net training step: pred %.3g mm (f32 oracle %.3g), dL/dheatmaps %.3g (%.3g), dL/dparams %.3g (%.3g)"
          % (e_pred, r_pred, e_h, r_h, worst, worst_ref))
    assert worst <= max(1e-2, 5 * worst_ref), (worst, worst_ref)
