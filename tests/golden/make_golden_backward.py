#!/usr/bin/env python
"""Golden GRADIENTS from the unmodified reference (``/root/reference/lib``), CPU autograd, seeded synthetic inputs:
``ProjectLayer`` (the per-person case of ``project_layer_pose.npz``: rotation / scale augmentation, h-flip, one
invalid row), ``SoftArgmaxLayer`` and a training-mode ``Basic3DBlock``.  Build container only:
``python tests/golden/make_golden_backward.py`` -> ``tests/golden/backward.npz``."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_import  # noqa: E402

ref_import.install()

from core.config import config as ref_cfg  # noqa: E402
import models  # noqa: E402,F401
from models.project_layer import ProjectLayer  # noqa: E402
from models.v2v_net import Basic3DBlock, V2VNet  # noqa: E402
from models.pose_regression_net import SoftArgmaxLayer  # noqa: E402

torch.set_num_threads(8)


def supervised_step(rs):
    """One supervised training forward/backward of the reference ``MultiPersonPoseNet`` (root net + pose net in
    .train(), heat-maps given) on the frames of ``inference_small.npz``; ground-truth roots are placed next to the
    proposals the (randomly weighted) root net makes, so that the proposal matching (< 500 mm) finds people."""
    import copy
    from models import multi_person_posenet
    from selfpose3d_b200 import synthetic
    g = dict(np.load(os.path.join(HERE, "inference_small.npz")))
    J = int(g["num_joints"])
    c = ref_cfg
    c.NETWORK.IMAGE_SIZE, c.NETWORK.HEATMAP_SIZE = np.array(g["image_size"]), np.array(g["heatmap_size"])
    c.NETWORK.NUM_JOINTS = J
    c.DATASET.ROOTIDX = c.DATASET.ROOTIDX_PSEUDO = 2
    c.NETWORK.ROOTNET_ROOTHM, c.NETWORK.USE_GT, c.NETWORK.TRAIN_ONLY_2D, c.NETWORK.BETA = True, False, False, 100.0
    c.MULTI_PERSON.SPACE_SIZE = [float(v) for v in g["space_size"]]
    c.MULTI_PERSON.SPACE_CENTER = [float(v) for v in g["space_center"]]
    c.MULTI_PERSON.INITIAL_CUBE_SIZE = [int(v) for v in g["initial_cube_size"]]
    c.MULTI_PERSON.MAX_PEOPLE_NUM = int(g["max_people"])
    c.MULTI_PERSON.THRESHOLD = float(g["threshold"])
    c.PICT_STRUCT.GRID_SIZE = [float(v) for v in g["grid_size"]]
    c.PICT_STRUCT.CUBE_SIZE = [int(v) for v in g["cube_size"]]
    c.BACKBONE_MODEL, c.MODEL = "", "multi_person_posenet"
    V, B = g["heatmaps"].shape[:2]
    K = int(g["max_people"])
    meta = [{"center": torch.from_numpy(g["center"][v]), "scale": torch.from_numpy(g["scale"][v]),
             "rotation": torch.from_numpy(g["rotation"][v]),
             "camera": {k[4:]: torch.from_numpy(g[k][v]) for k in g if k.startswith("cam_")}} for v in range(V)]
    model = multi_person_posenet.get_multi_person_pose_net(c, is_train=True)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=81), strict=True)
    model.train()
    hms = [torch.from_numpy(g["heatmaps"][v]).clone().requires_grad_(True) for v in range(V)]
    with torch.no_grad():
        _, gc0 = copy.deepcopy(model).root_net([h.detach() for h in hms], meta)
    roots = torch.zeros(B, K, 3, dtype=torch.float64)
    roots[:, :2] = (gc0[:, :2, :3] + torch.tensor([120.0, -80.0, 60.0])).double()
    num_person = torch.tensor([2, 1][:B] + [1] * max(0, B - 2))
    joints = roots[:, :, None, :] + torch.from_numpy(rs.randn(B, K, J, 3) * 200.0)
    vis = torch.ones(B, K, J, 3, dtype=torch.float64)
    vis[0, 0, 1] = 0.0
    meta[0].update(roots_3d=roots, num_person=num_person, joints_3d=joints, joints_3d_vis=vis)
    targets_3d = torch.from_numpy(rs.rand(B, *c.MULTI_PERSON.INITIAL_CUBE_SIZE).astype(np.float32))
    pred, _, gc, loss_2d, loss_3d, loss_cord = model(views=None, meta=meta, targets_3d=targets_3d, input_heatmaps=hms)
    assert int((gc[:, :, 3] >= 0).sum()) >= 2, gc[:, :, 3]
    (loss_3d + loss_cord).backward()
    named = list(model.named_parameters())
    return dict(sup_seed=81, sup_roots_3d=roots.numpy(), sup_num_person=num_person.numpy(), sup_joints_3d=joints.numpy(),
                sup_joints_3d_vis=vis.numpy(), sup_targets_3d=targets_3d.numpy(), sup_pred=pred.detach().numpy(),
                sup_grid_centers=gc.detach().numpy(), sup_loss_3d=float(loss_3d), sup_loss_cord=float(loss_cord),
                sup_grad_heatmaps=np.stack([h.grad.numpy() for h in hms]),
                sup_param_names=np.array([n for n, _ in named]),
                sup_param_grad_norm=np.array([0.0 if p.grad is None else float(p.grad.double().norm()) for _, p in named]),
                sup_param_grad_sum=np.array([0.0 if p.grad is None else float(p.grad.double().sum()) for _, p in named]))


def main():
    out = {}
    # ---- ProjectLayer: inputs of the committed forward golden
    g = dict(np.load(os.path.join(HERE, "project_layer_pose.npz")))
    ref_cfg.NETWORK.IMAGE_SIZE = np.array(g["image_size"])
    ref_cfg.NETWORK.HEATMAP_SIZE = np.array(g["heatmap_size"])
    ref_cfg.NETWORK.NUM_JOINTS = int(g["heatmaps"].shape[2])
    V = g["heatmaps"].shape[0]
    meta = [{"center": torch.from_numpy(g["center"][c]), "scale": torch.from_numpy(g["scale"][c]),
             "rotation": torch.from_numpy(g["rotation"][c]),
             "camera": {k[4:]: torch.from_numpy(g[k][c]) for k in g if k.startswith("cam_")}} for c in range(V)]
    hms = [torch.from_numpy(g["heatmaps"][c]).clone().requires_grad_(True) for c in range(V)]
    cubes, _ = ProjectLayer(ref_cfg)(hms, meta, [float(v) for v in g["grid_size"]], torch.from_numpy(g["grid_center"]),
                                     [int(v) for v in g["cube_size"]], flip_xcoords=torch.from_numpy(g["flip"]))
    assert np.allclose(cubes.detach().numpy(), g["cubes"], atol=1e-6)
    rs = np.random.RandomState(31)
    gc = torch.from_numpy(rs.randn(*cubes.shape).astype(np.float32))
    (cubes * gc).sum().backward()
    out["pl_grad_cubes"] = gc.numpy()
    out["pl_grad_heatmaps"] = np.stack([h.grad.numpy() for h in hms])

    # ---- SoftArgmaxLayer: inputs of softargmax.npz
    s = dict(np.load(os.path.join(HERE, "softargmax.npz")))
    ref_cfg.NETWORK.BETA = float(s["beta"])
    x = torch.from_numpy(s["x"]).clone().requires_grad_(True)
    o = SoftArgmaxLayer(ref_cfg)(x, torch.from_numpy(s["grids"]))
    go = torch.from_numpy(rs.randn(*o.shape).astype(np.float32))
    (o * go).sum().backward()
    out["sa_grad_out"] = go.numpy()
    out["sa_grad_x"] = x.grad.numpy()

    # ---- Basic3DBlock(4, 8, 3) in training mode (batch statistics)
    torch.manual_seed(7)
    blk = Basic3DBlock(4, 8, 3).train()
    with torch.no_grad():
        blk.block[1].weight.uniform_(0.5, 1.5)
        blk.block[1].bias.normal_(0, 0.2)
    xb = torch.from_numpy(rs.rand(2, 4, 6, 5, 4).astype(np.float32)).requires_grad_(True)
    y = blk(xb)
    gy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
    (y * gy).sum().backward()
    out.update(b3_x=xb.detach().numpy(), b3_w=blk.block[0].weight.detach().numpy(), b3_b=blk.block[0].bias.detach().numpy(),
               b3_gamma=blk.block[1].weight.detach().numpy(), b3_beta=blk.block[1].bias.detach().numpy(),
               b3_grad_y=gy.numpy(), b3_y=y.detach().numpy(), b3_grad_x=xb.grad.numpy(),
               b3_grad_w=blk.block[0].weight.grad.numpy(), b3_grad_b=blk.block[0].bias.grad.numpy(),
               b3_grad_gamma=blk.block[1].weight.grad.numpy(), b3_grad_beta=blk.block[1].bias.grad.numpy())
    # ---- whole V2VNet(3, 3) in training mode: output, input gradient, per-parameter gradient norms / sums and the
    # updated running statistics of the first BatchNorm (weights regenerated from the seed by the tests)
    from selfpose3d_b200 import synthetic
    net = V2VNet(3, 3)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=53), strict=True)
    net.train()
    xv = torch.from_numpy(rs.rand(2, 3, 8, 8, 4).astype(np.float32)).requires_grad_(True)
    yv = net(xv)
    gv = torch.from_numpy(rs.randn(*yv.shape).astype(np.float32))
    (yv * gv).sum().backward()
    names = [n for n, _ in net.named_parameters()]
    out.update(v2v_seed=53, v2v_x=xv.detach().numpy(), v2v_grad_y=gv.numpy(), v2v_y=yv.detach().numpy(),
               v2v_grad_x=xv.grad.numpy(), v2v_param_names=np.array(names),
               v2v_param_grad_norm=np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()]),
               v2v_param_grad_sum=np.array([float(p.grad.double().sum()) for _, p in net.named_parameters()]),
               v2v_bn0_running_mean=net.front_layers[0].block[1].running_mean.numpy(),
               v2v_bn0_running_var=net.front_layers[0].block[1].running_var.numpy())
    out.update(supervised_step(rs))
    path = os.path.join(HERE, "backward.npz")
    np.savez_compressed(path, **out)
    print("backward.npz %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
