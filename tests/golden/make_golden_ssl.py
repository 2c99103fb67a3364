#!/usr/bin/env python
"""One SELF-SUPERVISED training step (forward + backward) of the unmodified reference ``MultiPersonPoseNetSSV``
(``/root/reference/lib``) on CPU: 3 views x 1 frame of 3 x 96 x 64 noise images per augmentation set, ResNet-50
backbone + ResNet-18 attention net in .train(), frozen root net, pose net in .train(), re-projection / Gaussian
rendering / attention / Hungarian-L1 losses (``cam5_posenet.yaml`` semantics at toy size).  Build container only:
``python tests/golden/make_golden_ssl.py`` -> ``tests/golden/ssl_step.npz`` (inputs are regenerated from seeds by
the tests through ``ssl_case`` below; only the results are stored)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

J, V, B, K = 4, 3, 1, 3
IMAGE, HEATMAP = (64, 96), (16, 24)


def configure(c):
    c.NETWORK.IMAGE_SIZE, c.NETWORK.HEATMAP_SIZE = np.array(IMAGE), np.array(HEATMAP)
    c.NETWORK.NUM_JOINTS = J
    c.DATASET.ROOTIDX = c.DATASET.ROOTIDX_PSEUDO = 2
    c.NETWORK.ROOTNET_ROOTHM, c.NETWORK.USE_GT, c.NETWORK.TRAIN_ONLY_2D, c.NETWORK.BETA = True, False, False, 100.0
    c.NETWORK.TRAIN_ONLY_ROOTNET = c.NETWORK.ROOTNET_TRAIN_SYNTH = c.NETWORK.SINGLE_AUG_TRAINING_POSENET = False
    c.NETWORK.FREEZE_ROOTNET, c.NETWORK.INIT_TRAIN_EPOCHS_ROOTNET, c.NETWORK.PRETRAINED = True, 0, ""
    c.MULTI_PERSON.SPACE_SIZE, c.MULTI_PERSON.SPACE_CENTER = [8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0]
    c.MULTI_PERSON.INITIAL_CUBE_SIZE, c.MULTI_PERSON.MAX_PEOPLE_NUM, c.MULTI_PERSON.THRESHOLD = [16, 16, 8], K, -1.0
    c.PICT_STRUCT.GRID_SIZE, c.PICT_STRUCT.CUBE_SIZE = [2000.0, 2000.0, 2000.0], [16, 16, 16]
    c.BACKBONE_MODEL, c.MODEL = "pose_resnet", "multi_person_posenet_ssv"
    c.WITH_ATTN, c.ATTN_WEIGHT, c.ATTN_NUM_LAYERS = True, 0.1, 18
    c.USE_L1, c.L1_WEIGHT, c.L1_ATTN = True, 0.1, True
    c.TRAIN.L1_EPOCH, c.TRAIN.BATCH_SIZE = 0, B
    c.EVAL_ROOTNET_ONLY = False
    return c


def ssl_case():
    """Seeded inputs of the step: three view sets (set 1: rotation / scale jitter; set 2: other jitter + h-flip; set 3
    plain), pseudo heat-maps, pseudo 2-D poses."""
    from selfpose3d_b200 import synthetic
    return synthetic.ssl_training_case(IMAGE, HEATMAP, J, B, V, K)


SYNTH_TORCH_SEED = 1234


def synth_root_step(cfg):
    """The synthetic-root RootNet branch of the reference ``CuboidProposalNetSoft`` (.train(), ROOTNET_TRAIN_SYNTH) on
    set 1 of ``ssl_case`` with seeded torch draws: real volume, synthetic volume, target volume, proposals and the
    gradients of ``100 * mse(synthetic, target)``."""
    from models.cuboid_proposal_net_soft import CuboidProposalNetSoft
    from selfpose3d_b200 import synthetic
    cfg.NETWORK.ROOTNET_TRAIN_SYNTH = True
    try:
        net = CuboidProposalNetSoft(cfg)
    finally:
        cfg.NETWORK.ROOTNET_TRAIN_SYNTH = False
    net.load_state_dict(synthetic.trained_like_state_dict(net, seed=96), strict=True)
    net.train()
    (_, meta, targets), _, _ = ssl_case()
    torch.manual_seed(SYNTH_TORCH_SEED)
    main, syn, target, gc = net(targets, meta, flip_xcoords=meta[0]["hflip"])
    (100.0 * torch.nn.functional.mse_loss(syn, target)).backward()
    named = list(net.named_parameters())
    return dict(syn_seed=96, syn_main=main.detach().numpy(), syn_cubes=syn.detach().numpy(), syn_target=target.numpy(),
                syn_grid_centers=gc.detach().numpy(), syn_param_names=np.array([n for n, _ in named]),
                syn_param_grad_norm=np.array([float(p.grad.double().norm()) for _, p in named]))


def main():
    sys.path.insert(0, HERE)
    import ref_import
    ref_import.install()
    from core.config import config as ref_cfg
    import models  # noqa: F401
    from models import multi_person_posenet_ssv
    from selfpose3d_b200 import synthetic
    torch.set_num_threads(8)
    cfg = configure(ref_cfg)
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=True)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=95), strict=True)
    model.train()
    model.root_net.eval()
    (v1, m1, t1), (v2, m2, t2), (v3, m3, t3) = ssl_case()
    pred, hm3, gc, losses = model(views1=v1, meta1=m1, targets_2d1=t1, views2=v2, meta2=m2, targets_2d2=t2,
                                  views3=v3, meta3=m3, targets_2d3=t3, inference=False, epoch=1)
    total = sum(losses.values())
    total.backward()
    named = list(model.named_parameters())
    out = dict(seed=95, pred=pred.numpy(), grid_centers=gc.detach().numpy(), heatmaps3=torch.stack(hm3).detach().numpy(),
               loss_names=np.array(sorted(losses)), loss_values=np.array([float(losses[k]) for k in sorted(losses)]),
               param_names=np.array([n for n, _ in named]),
               param_grad_norm=np.array([-1.0 if p.grad is None else float(p.grad.double().norm()) for _, p in named]),
               param_grad_sum=np.array([0.0 if p.grad is None else float(p.grad.double().sum()) for _, p in named]))
    out.update(synth_root_step(cfg))
    path = os.path.join(HERE, "ssl_step.npz")
    np.savez_compressed(path, **out)
    print("ssl_step.npz %.1f KB" % (os.path.getsize(path) / 1024), dict(zip(out["loss_names"], out["loss_values"])))
    print("grid_centers flags", gc[:, :, 3:].tolist(), "params with grad", int((out["param_grad_norm"] >= 0).sum()), len(named))


if __name__ == "__main__":
    main()
