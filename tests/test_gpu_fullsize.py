"""Full-size parity: the BASELINE.json configurations end to end through the module API on the B200 against
``oracle/pipeline.py`` in float32 AND float64 (and against the staged unmodified reference itself, ``oracle/_ref``,
when it travelled to the box).

  config 1 -- one synthetic frame, 5 views of 3x384x288, PoseResNet-50 -> 80x80x20 root grid -> 10 proposals (all slots
              forced valid) -> 64^3 person cubes -> V2VNet -> soft-argmax  (BASELINE.json configs[0] / configs[2] geometry)
  config 2 -- Panoptic-shaped heat-maps [4, 15, 128, 240] x 5 views, ``CuboidProposalNet`` only, C = 15 (configs[1])

Bars (north star: heat-maps within 1e-4 rel-fp32, coordinates within 1e-3 mm): heat-maps and score volumes
max |ours - oracle_f32| <= 1e-4 of the range; proposals identical (mismatches counted and asserted zero); joints
|ours - f64| <= max(1.5 * |oracle_f32 - f64|, 1e-3 mm) -- i.e. our float32 result may be no further from exact
arithmetic than 1.5x what the reference's own float32 evaluation is (soft-argmax with beta = 100 amplifies float32
summation-order noise of the logits; both numbers are printed and appended to ``gpurun_out/fullsize_parity.txt``).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

from oracle import pipeline, ref_runner  # noqa: E402
from selfpose3d_b200 import ops, synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import cuboid_proposal_net, multi_person_posenet_ssv  # noqa: E402

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMAGE_SIZE, HEATMAP_SIZE, VIEWS, PROPOSALS = [288, 384], [72, 96], 5, 10


def note(line):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "fullsize_parity.txt"), "a") as f:
            f.write(line + "\n")
    print(line)


def _oracle_cfg(cfg):
    return dict(image_size=cfg.NETWORK.IMAGE_SIZE, heatmap_size=cfg.NETWORK.HEATMAP_SIZE,
                space_size=cfg.MULTI_PERSON.SPACE_SIZE, space_center=cfg.MULTI_PERSON.SPACE_CENTER,
                initial_cube_size=cfg.MULTI_PERSON.INITIAL_CUBE_SIZE, grid_size=cfg.PICT_STRUCT.GRID_SIZE,
                cube_size=cfg.PICT_STRUCT.CUBE_SIZE, max_people=cfg.MULTI_PERSON.MAX_PEOPLE_NUM,
                threshold=cfg.MULTI_PERSON.THRESHOLD, beta=cfg.NETWORK.BETA, root_idx=cfg.DATASET.ROOTIDX)


@pytest.fixture(scope="module")
def config1():
    """Model, inputs and the CPU answers (float32 oracle, float64 oracle on the float32 proposals, the reference)."""
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = list(IMAGE_SIZE), list(HEATMAP_SIZE)
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = PROPOSALS
    cfg.MULTI_PERSON.THRESHOLD = -1e9
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    sd = synthetic.trained_like_state_dict(model, seed=0)
    model.load_state_dict(sd, strict=True)
    cams = synthetic.ring_cameras(VIEWS, seed=0)
    meta = synthetic.make_meta(cams, 1, IMAGE_SIZE)
    images = synthetic.random_images(1, VIEWS, IMAGE_SIZE, seed=3)
    cam_arrays = {k: np.stack([m["camera"][k].numpy() for m in meta]) for k in meta[0]["camera"]}
    args = (sd, _oracle_cfg(cfg), cam_arrays, [m["center"].numpy() for m in meta], [m["scale"].numpy() for m in meta],
            [m["rotation"].numpy() for m in meta])
    with torch.no_grad():
        p32, h32, g32, r32 = pipeline.inference(*args, images=images)
        p64, h64, _, r64 = pipeline.inference(*args, images=images, dtype=torch.float64, grid_centers=g32)
    out = dict(cfg=cfg, model=model.to(DEV).eval(), meta=meta, images=images, p32=p32, h32=h32, g32=g32, r32=r32,
               p64=p64, h64=h64, r64=r64, ref=None)
    if ref_runner.available():
        ocfg = _oracle_cfg(cfg)
        ref_model, _ = ref_runner.build_model(state_dict=sd, num_joints=cfg.NETWORK.NUM_JOINTS,
                                              **{k: ocfg[k] for k in ocfg if k != "root_idx"}, root_idx=ocfg["root_idx"])
        out["ref"] = ref_runner.inference(ref_model, images, meta)
    return out


def test_oracle_port_equals_the_reference_at_full_size(config1):
    """Pins the oracle at BASELINE size: the port and the unmodified reference (both torch CPU float32) agree."""
    c = config1
    if c["ref"] is None:
        pytest.skip("oracle/_ref not staged on this box (sh oracle/make_ref.sh in the build container)")
    pred, hms, gc = c["ref"]
    hm_err = max(float((a - b).abs().max()) for a, b in zip(hms, c["h32"])) / max(float(b.abs().max()) for b in c["h32"])
    gc_err = float((gc[..., :3] - c["g32"][..., :3]).abs().max())
    j_err = float((pred[..., :3] - c["p32"][..., :3]).abs().max())
    note("config1 oracle port vs unmodified reference: heat-maps %.3g of range, proposals %.3g mm, joints %.3g mm"
         % (hm_err, gc_err, j_err))
    assert hm_err <= 1e-5 and gc_err <= 1e-3 and j_err <= 1e-2


@pytest.mark.parametrize("mode", ["bf16x3", "simt"])
def test_config1_full_path_vs_oracle_f32_and_f64(config1, mode):
    c = config1
    ops.set_volume_dtype(torch.float32)
    ops.set_float32_conv(mode)
    try:
        with torch.no_grad():
            pred, hms, gc = c["model"](views1=[im.to(DEV) for im in c["images"]], meta1=c["meta"], inference=True)
        pred, gc = pred.cpu(), gc.cpu()
        hms = [h.cpu() for h in hms]
    finally:
        ops.set_float32_conv("bf16x3")
    scale = max(float(h.abs().max()) for h in c["h32"])
    hm_err = max(float((a - b).abs().max()) for a, b in zip(hms, c["h32"])) / scale
    hm_ref = max(float((a.double() - b).abs().max()) for a, b in zip(c["h32"], c["h64"])) / scale
    hm_our = max(float((a.double() - b).abs().max()) for a, b in zip(hms, c["h64"])) / scale
    # proposals: every slot on the oracle's voxel (same order); scores within float32 noise
    mism = int(((gc[..., :3] - c["g32"][..., :3]).abs().amax(-1) > 1e-3).sum())
    score_err = float((gc[..., 4] - c["g32"][..., 4]).abs().max())
    same = (gc[..., :3] - c["g32"][..., :3]).abs().amax(-1) <= 1e-3
    j_our = float((pred[..., :3].double() - c["p64"][..., :3])[same].abs().max())
    j_ref = float((c["p32"][..., :3].double() - c["p64"][..., :3])[same].abs().max())
    j_vs32 = float((pred[..., :3] - c["p32"][..., :3])[same].abs().max())
    note("config1 %s: heat-maps |ours-f32| %.3g, |ours-f64| %.3g, |f32-f64| %.3g of range; proposals %d/%d mismatched, "
         "score diff %.3g; joints |ours-f64| %.4g mm, |oracle_f32-f64| %.4g mm, |ours-oracle_f32| %.4g mm"
         % (mode, hm_err, hm_our, hm_ref, mism, gc.shape[0] * gc.shape[1], score_err, j_our, j_ref, j_vs32))
    assert hm_err <= 1e-4, hm_err
    assert mism == 0, mism
    assert bool((gc[..., 3] >= 0).all()) and bool((pred[..., 3] >= 0).all())
    assert j_our <= max(1.5 * j_ref, 1e-3), (j_our, j_ref)


@pytest.fixture(scope="module")
def config2():
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [960, 512], [240, 128]
    cfg.NETWORK.ROOTNET_ROOTHM = False            # CuboidProposalNet of the shipped prn64_cpn80x80x20 config: all 15 joints
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = PROPOSALS
    cfg.MULTI_PERSON.THRESHOLD = -1e9
    B = 4
    net = cuboid_proposal_net.CuboidProposalNet(cfg)
    sd = synthetic.trained_like_state_dict(net, seed=7)
    net.load_state_dict(sd, strict=True)
    cams = synthetic.ring_cameras(VIEWS, seed=1)
    meta = synthetic.make_meta(cams, B, cfg.NETWORK.IMAGE_SIZE)
    people = synthetic.synthetic_people(B, seed=5, num_joints=15)
    hms = synthetic.render_heatmaps(people, meta, cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE, num_joints=15, sigma=3.0)
    cam_arrays = {k: np.stack([m["camera"][k].numpy() for m in meta]) for k in meta[0]["camera"]}
    geo = (cam_arrays, [m["center"].numpy() for m in meta], [m["scale"].numpy() for m in meta],
           [m["rotation"].numpy() for m in meta], cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE,
           cfg.MULTI_PERSON.SPACE_SIZE, [cfg.MULTI_PERSON.SPACE_CENTER], cfg.MULTI_PERSON.INITIAL_CUBE_SIZE)
    from oracle import nets, volume_ops
    vsd = {k[len("v2v_net."):]: v for k, v in sd.items() if k.startswith("v2v_net.")}
    with torch.no_grad():
        c32, _ = pipeline.unproject_torch(hms, *geo)
        r32 = nets.v2v_forward(c32, vsd)[:, 0]
        c64, _ = pipeline.unproject_torch(hms, *geo, dtype=torch.float64)
        r64 = nets.v2v_forward(c64, vsd, dtype=torch.float64)[:, 0]
    g32 = volume_ops.proposal_layer(r32.numpy(), cfg.MULTI_PERSON.SPACE_SIZE, cfg.MULTI_PERSON.SPACE_CENTER,
                                    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE, PROPOSALS, cfg.MULTI_PERSON.THRESHOLD)
    return dict(net=net.to(DEV).eval(), meta=meta, hms=hms, r32=r32, r64=r64, g32=torch.from_numpy(g32))


@pytest.mark.parametrize("mode", ["bf16x3", "simt"])
def test_config2_cuboid_proposal_net_vs_oracle(config2, mode):
    c = config2
    ops.set_volume_dtype(torch.float32)
    ops.set_float32_conv(mode)
    try:
        with torch.no_grad():
            rc, gc = c["net"]([h.to(DEV) for h in c["hms"]], c["meta"])
        rc, gc = rc.cpu(), gc.cpu()
    finally:
        ops.set_float32_conv("bf16x3")
    scale = float(c["r32"].abs().max())
    err32 = float((rc - c["r32"]).abs().max()) / scale
    our64 = float((rc.double() - c["r64"]).abs().max()) / scale
    ref64 = float((c["r32"].double() - c["r64"]).abs().max()) / scale
    mism = int(((gc[..., :3] - c["g32"][..., :3]).abs().amax(-1) > 1e-3).sum())
    note("config2 %s: score volume |ours-f32| %.3g, |ours-f64| %.3g, |f32-f64| %.3g of range; proposals %d/%d mismatched"
         % (mode, err32, our64, ref64, mism, gc.shape[0] * gc.shape[1]))
    assert err32 <= 1e-4, err32
    assert mism == 0, mism
