#!/usr/bin/env python
"""Generate the golden input/output vectors under ``tests/golden/`` by running the
UNMODIFIED reference (``/root/reference/lib``) on CPU with seeded synthetic inputs.

The reference has no tests or golden vectors of its own, so these
fixtures are what pins the oracle (``oracle/``) and, through it, the CUDA path.
Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Outputs are small ``.npz`` files.
"""
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
from models.v2v_net import V2VNet, Basic3DBlock, Res3DBlock, Upsample3DBlock  # noqa: E402
from models.cuboid_proposal_net_soft import ProposalLayerSoft, CuboidProposalNetSoft  # noqa: E402
from models.cuboid_proposal_net import CuboidProposalNet  # noqa: E402
from models.pose_regression_net import SoftArgmaxLayer, PoseRegressionNet  # noqa: E402
from models import pose_resnet, multi_person_posenet_ssv, multi_person_posenet  # noqa: E402
import utils.cameras as ref_cameras  # noqa: E402
import utils.transforms as ref_transforms  # noqa: E402
from core.proposal import nms as ref_nms  # noqa: E402

from selfpose3d_b200 import synthetic  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def meta_to_arrays(meta):
    """Flatten a meta list into arrays ``[V,B,...]`` for storage."""
    d = {
        "center": torch.stack([m["center"] for m in meta]),
        "scale": torch.stack([m["scale"] for m in meta]),
        "rotation": torch.stack([m["rotation"] for m in meta]),
    }
    for k in meta[0]["camera"]:
        d["cam_" + k] = torch.stack([m["camera"][k] for m in meta])
    return d


def set_geometry(image_size, heatmap_size, num_joints=15):
    ref_cfg.NETWORK.IMAGE_SIZE = np.array(image_size)
    ref_cfg.NETWORK.HEATMAP_SIZE = np.array(heatmap_size)
    ref_cfg.NETWORK.NUM_JOINTS = num_joints


# ---------------------------------------------------------------- A. affine helpers
def gold_affine():
    rs = np.random.RandomState(1)
    cases = []
    outs = []
    for (orig, size) in [((1920, 1080), (288, 384)), ((1920, 1080), (960, 512)), ((640, 480), (72, 96))]:
        base = ref_transforms.get_scale(orig, size)
        for rot in [0.0, 17.0, -45.0]:
            for mul in [1.0, 1.2, 0.65]:
                center = np.array([orig[0] / 2.0, orig[1] / 2.0]) + rs.randn(2) * 3
                scale = (base * mul).astype(np.float32)
                t = ref_transforms.get_affine_transform(center, scale, rot, size)
                cases.append([center[0], center[1], scale[0], scale[1], rot, size[0], size[1], orig[0], orig[1]])
                outs.append(t)
    save("affine", cases=np.array(cases, dtype=np.float64), trans=np.array(outs, dtype=np.float64))


# ---------------------------------------------------------------- B. camera projection
def gold_project_pose():
    cams = synthetic.ring_cameras(3, seed=3)
    rs = np.random.RandomState(2)
    pts = np.concatenate([rs.uniform(-4000, 4000, (96, 2)), rs.uniform(-200, 1800, (96, 1))], axis=1)
    pts[:, 1] -= 500
    x = torch.as_tensor(pts, dtype=torch.float32)
    outs = []
    for cam in cams:
        outs.append(ref_cameras.project_pose(x, {k: torch.as_tensor(v) for k, v in cam.items()}))
    arr = {"points": x, "pixels": torch.stack(outs)}
    for k in cams[0]:
        arr["cam_" + k] = np.stack([np.asarray(c[k]) for c in cams])
    save("project_pose", **arr)


# ---------------------------------------------------------------- C. ProjectLayer
def gold_project_layer():
    # c1: whole-space (root) geometry, portrait network input (the BASELINE 288x384 letter-box branch)
    set_geometry((72, 96), (18, 24))
    cams = synthetic.ring_cameras(5, seed=0)
    meta = synthetic.make_meta(cams, 2, (72, 96))
    people = synthetic.synthetic_people(2, seed=0, num_joints=2)
    hms = synthetic.render_heatmaps(people, meta, (72, 96), (18, 24), num_joints=2, sigma=1.5)
    layer = ProjectLayer(ref_cfg)
    cubes, grids = layer(hms, meta, [8000.0, 8000.0, 2000.0], [[0.0, -500.0, 800.0]], [10, 8, 6])
    save("project_layer_root", heatmaps=torch.stack(hms), cubes=cubes, grids=grids,
         image_size=[72, 96], heatmap_size=[18, 24], grid_size=[8000.0, 8000.0, 2000.0],
         grid_center=[[0.0, -500.0, 800.0]], cube_size=[10, 8, 6], **meta_to_arrays(meta))

    # c2: per-person cubes, landscape input (the shipped 960x512 branch), rotation/scale aug, h-flip, one invalid row
    set_geometry((120, 64), (30, 16))
    cams = synthetic.ring_cameras(4, seed=5)
    rot = [[0.0, 17.0, -30.0]] * 4
    mul = [[1.0, 1.2, 0.8]] * 4
    meta = synthetic.make_meta(cams, 3, (120, 64), rotation=rot, scale_mul=mul)
    people = synthetic.synthetic_people(3, seed=4, num_joints=3)
    hms = synthetic.render_heatmaps(people, meta, (120, 64), (30, 16), num_joints=3, sigma=1.5)
    rs = np.random.RandomState(9)
    hms = [h + torch.from_numpy(0.05 * rs.rand(*h.shape).astype(np.float32)) for h in hms]
    centers = torch.tensor([[people[0][0][0][0], people[0][0][0][1], people[0][0][0][2], 0.0, 0.9],
                            [0.0, 0.0, 0.0, -1.0, 0.1],
                            [people[2][0][0][0] + 300, people[2][0][0][1], people[2][0][0][2], 1.0, 0.7]],
                           dtype=torch.float32)
    flip = torch.tensor([False, False, True])
    cubes, grids = layer.__class__(ref_cfg)(hms, meta, [2000.0, 2000.0, 2000.0], centers, [8, 8, 8],
                                            flip_xcoords=flip)
    save("project_layer_pose", heatmaps=torch.stack(hms), cubes=cubes, grids=grids,
         image_size=[120, 64], heatmap_size=[30, 16], grid_size=[2000.0, 2000.0, 2000.0],
         grid_center=centers, cube_size=[8, 8, 8], flip=flip, **meta_to_arrays(meta))


# ---------------------------------------------------------------- D. NMS / proposals
def gold_proposal():
    # YAML-loaded configs carry python lists -> float32 tensors in ProposalLayer (the shipped path);
    # config.py defaults are float64 numpy arrays -> get_real_loc promotes to float64 ("proposal_f64").
    ref_cfg.MULTI_PERSON.SPACE_SIZE = [8000.0, 8000.0, 2000.0]
    ref_cfg.MULTI_PERSON.SPACE_CENTER = [0.0, -500.0, 800.0]
    ref_cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [12, 10, 6]
    ref_cfg.MULTI_PERSON.MAX_PEOPLE_NUM = 5
    ref_cfg.MULTI_PERSON.THRESHOLD = 0.3
    rs = np.random.RandomState(11)
    cubes = torch.from_numpy(rs.rand(3, 12, 10, 6).astype(np.float32))
    cubes[2] *= 0.25  # below threshold everywhere
    vals, idx = ref_nms(cubes, 5)
    layer = ProposalLayerSoft(ref_cfg)
    layer.eval()
    gc = layer(cubes, None, None)
    save("proposal", root_cubes=cubes, topk_values=vals, topk_index=idx, grid_centers=gc,
         space_size=[8000.0, 8000.0, 2000.0], space_center=[0.0, -500.0, 800.0],
         cube_size=[12, 10, 6], max_people=5, threshold=0.3)
    ref_cfg.MULTI_PERSON.SPACE_SIZE = np.array([8000.0, 8000.0, 2000.0])
    ref_cfg.MULTI_PERSON.SPACE_CENTER = np.array([0.0, -500.0, 800.0])
    layer = ProposalLayerSoft(ref_cfg)
    layer.eval()
    save("proposal_f64", grid_centers=layer(cubes, None, None))


# ---------------------------------------------------------------- E. soft-argmax
def gold_softargmax():
    ref_cfg.NETWORK.BETA = 100.0
    rs = np.random.RandomState(12)
    x = torch.from_numpy((rs.rand(2, 3, 6, 5, 4) * 0.2).astype(np.float32))
    x[0, 0, 2, 3, 1] = 0.9
    x[1, 2, 5, 0, 3] = 0.5
    grids = torch.from_numpy(rs.uniform(-1000, 1000, (2, 120, 3)).astype(np.float32))
    out = SoftArgmaxLayer(ref_cfg)(x, grids)
    save("softargmax", x=x, grids=grids, out=out, beta=100.0)


# ---------------------------------------------------------------- F. V2V blocks and nets
def _run_module(mod, x, seed):
    sd = synthetic.trained_like_state_dict(mod, seed=seed)
    mod.load_state_dict(sd, strict=True)
    mod.eval()
    with torch.no_grad():
        y = mod(x)
    return sd, y


def gold_v2v():
    rs = np.random.RandomState(13)
    x = torch.from_numpy(rs.rand(2, 4, 6, 5, 4).astype(np.float32))
    arrays = {"x": x}
    for name, mod in [("basic7", Basic3DBlock(4, 8, 7)), ("basic3", Basic3DBlock(4, 8, 3)),
                      ("res_4_8", Res3DBlock(4, 8)), ("res_4_4", Res3DBlock(4, 4)),
                      ("up_4_8", Upsample3DBlock(4, 8, 2, 2))]:
        sd, y = _run_module(mod, x, seed=20)
        arrays["y_" + name] = y
        for k, v in sd.items():
            arrays["w_%s_%s" % (name, k)] = v
    save("v2v_blocks", **arrays)

    # full nets: weights are regenerated from the seed by the test (too large to store)
    x1 = torch.from_numpy(rs.rand(2, 1, 8, 8, 4).astype(np.float32))
    _, y1 = _run_module(V2VNet(1, 1), x1, seed=21)
    x3 = torch.from_numpy(rs.rand(1, 3, 8, 8, 8).astype(np.float32))
    _, y3 = _run_module(V2VNet(3, 3), x3, seed=22)
    save("v2v_net", x1=x1, y1=y1, seed1=21, x3=x3, y3=y3, seed3=22)


# ---------------------------------------------------------------- G. PoseResNet-50
def gold_pose_resnet():
    set_geometry((48, 64), (12, 16), num_joints=15)
    ref_cfg.POSE_RESNET.NUM_LAYERS = 50
    net = pose_resnet.get_pose_net(ref_cfg, is_train=False)
    rs = np.random.RandomState(14)
    x = torch.from_numpy(rs.randn(2, 3, 64, 48).astype(np.float32))
    _, y = _run_module(net, x, seed=23)
    save("pose_resnet50", x=x, y=y, seed=23)


# ---------------------------------------------------------------- H. whole inference path
def gold_inference():
    J = 4
    set_geometry((72, 96), (18, 24), num_joints=J)
    ref_cfg.DATASET.ROOTIDX = 2
    ref_cfg.DATASET.ROOTIDX_PSEUDO = 2
    ref_cfg.NETWORK.ROOTNET_ROOTHM = True
    ref_cfg.NETWORK.ROOTNET_TRAIN_SYNTH = False
    ref_cfg.NETWORK.USE_GT = False
    ref_cfg.NETWORK.BETA = 100.0
    ref_cfg.WITH_ATTN = False
    ref_cfg.MULTI_PERSON.SPACE_SIZE = [8000.0, 8000.0, 2000.0]
    ref_cfg.MULTI_PERSON.SPACE_CENTER = [0.0, -500.0, 800.0]
    ref_cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [16, 16, 8]
    ref_cfg.MULTI_PERSON.MAX_PEOPLE_NUM = 3
    ref_cfg.PICT_STRUCT.GRID_SIZE = [2000.0, 2000.0, 2000.0]
    ref_cfg.PICT_STRUCT.CUBE_SIZE = [16, 16, 16]
    ref_cfg.BACKBONE_MODEL = "pose_resnet"
    ref_cfg.MODEL = "multi_person_posenet_ssv"

    cams = synthetic.ring_cameras(5, seed=0)
    meta = synthetic.make_meta(cams, 2, (72, 96))
    people = synthetic.synthetic_people(2, seed=2, num_joints=J)
    hms = synthetic.render_heatmaps(people, meta, (72, 96), (18, 24), num_joints=J, sigma=1.5)

    model = multi_person_posenet_ssv.get_multi_person_pose_net(ref_cfg, is_train=False)
    sd = synthetic.trained_like_state_dict(model, seed=30)
    model.load_state_dict(sd, strict=True)
    model.eval()
    with torch.no_grad():
        root_cubes, _, _, gc = model.root_net(hms, meta)
        # threshold between the 2nd and 3rd score of each sample so that valid and invalid slots both occur
        thr = float(((gc[:, 1, 4] + gc[:, 2, 4]) / 2).min())
        ref_cfg.MULTI_PERSON.THRESHOLD = thr
        model.root_net.proposal_layer.threshold = thr
        pred, _, gc = model(views1=None, meta1=meta, input_heatmaps1=hms, inference=True)
    save("inference_small", heatmaps=torch.stack(hms), root_cubes=root_cubes, grid_centers=gc, pred=pred,
         threshold=thr, seed=30, num_joints=J, image_size=[72, 96], heatmap_size=[18, 24],
         space_size=[8000.0, 8000.0, 2000.0], space_center=[0.0, -500.0, 800.0],
         initial_cube_size=[16, 16, 8], grid_size=[2000.0, 2000.0, 2000.0], cube_size=[16, 16, 16],
         max_people=3, **meta_to_arrays(meta))

    # images -> heat-maps -> pred through the backbone as well (tiny images)
    imgs = synthetic.random_images(1, 5, (72, 96), seed=3)
    meta1 = synthetic.make_meta(cams, 1, (72, 96))
    model.root_net.proposal_layer.threshold = -1.0  # every slot valid: heat-maps of noise images score low
    with torch.no_grad():
        pred2, hm2, gc2 = model(views1=imgs, meta1=meta1, inference=True)
    save("inference_images", images=torch.stack(imgs), heatmaps=torch.stack(hm2), grid_centers=gc2, pred=pred2,
         threshold=-1.0, seed=30)

    # supervised CuboidProposalNet with all J channels (ROOTNET_ROOTHM False), config-2 style
    ref_cfg.NETWORK.ROOTNET_ROOTHM = False
    net = CuboidProposalNet(ref_cfg)
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=31), strict=True)
    net.eval()
    with torch.no_grad():
        rc, gcs = net(hms, meta)
    save("cuboid_proposal_allj", root_cubes=rc, grid_centers=gcs, seed=31)


def gold_state_dict_keys():
    import json
    from easydict import EasyDict
    out = {}
    ref_cfg.NETWORK.NUM_JOINTS = 15
    ref_cfg.NETWORK.ROOTNET_ROOTHM = True
    ref_cfg.NETWORK.ROOTNET_TRAIN_SYNTH = False
    ref_cfg.WITH_ATTN = True
    ref_cfg.ATTN_NUM_LAYERS = 18
    ref_cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [80, 80, 20]
    ref_cfg.PICT_STRUCT.CUBE_SIZE = [64, 64, 64]
    m = multi_person_posenet_ssv.get_multi_person_pose_net(ref_cfg, is_train=False)
    out["multi_person_posenet_ssv_attn"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    ref_cfg.WITH_ATTN = False
    ref_cfg.NETWORK.ROOTNET_ROOTHM = False
    m = multi_person_posenet.get_multi_person_pose_net(ref_cfg, is_train=False)
    out["multi_person_posenet"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(out, f)
    print("state_dict_keys.json", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    gold_state_dict_keys()
    gold_affine()
    gold_project_pose()
    gold_project_layer()
    gold_proposal()
    gold_softargmax()
    gold_v2v()
    gold_pose_resnet()
    gold_inference()
