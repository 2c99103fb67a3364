#!/usr/bin/env python
"""BASELINE configs[4]: self-supervised training step (backbone + attention net + frozen root net + pose net, SSL
losses on pseudo heat-maps) data-parallel over the GPUs of one box.

    python profiles/config5_train.py                                   # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/config5_train.py --batch-per-gpu 2

Per rank: ``--batch-per-gpu`` frames of 5 views 3x384x288 in three augmented sets, ``MultiPersonPoseNetSSV`` in
.train() (root net .eval(): FREEZE_ROOTNET), forward + backward on the float32 training path, gradient average with
``selfpose3d_b200.dist.allreduce_gradients``, no optimizer.  Prints one JSON line (frames/s over all ranks, max over
ranks, CUDA events).  The model-level path is
covered on CPU (tests/test_training_cpu.py, tests/test_dist_gloo.py).  Round 2: run on 1 / 2 / 8 B200s
(profiles/r02_config5_train_*.json); the V2VNet weight gradients run on tcgen05 (sp3d_conv_wgrad_tc)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from selfpose3d_b200 import dist as sd  # noqa: E402
from selfpose3d_b200 import ops, synthetic, _lib  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet_ssv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--proposals", type=int, default=4)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    ops.set_volume_dtype(torch.float32)
    cfg = bench.make_cfg(a.batch_per_gpu)
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = a.proposals
    cfg.WITH_ATTN, cfg.USE_L1, cfg.TRAIN.L1_EPOCH = True, True, 0
    cfg.NETWORK.FREEZE_ROOTNET = True
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
    model = model.to(dev).train()
    model.root_net.eval()
    sets = synthetic.ssl_training_case(bench.IMAGE_SIZE, bench.HEATMAP_SIZE, cfg.NETWORK.NUM_JOINTS, a.batch_per_gpu,
                                       bench.VIEWS, a.proposals, seed=77 + rank, image_seed=40 + 3 * rank)
    (v1, m1, t1), (v2, m2, t2), (v3, m3, t3) = [([v.to(dev) for v in vs], m, [t.to(dev) for t in ts]) for vs, m, ts in sets]

    def step():
        model.zero_grad(set_to_none=True)
        _, _, _, losses = model(views1=v1, meta1=m1, targets_2d1=t1, views2=v2, meta2=m2, targets_2d2=t2,
                                views3=v3, meta3=m3, targets_2d3=t3, inference=False, epoch=1)
        sum(losses.values()).backward()
        if world > 1:
            sd.allreduce_gradients(model.parameters())
        return losses

    for _ in range(a.warmup):
        step()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count
    e0.record()
    for _ in range(a.steps):
        losses = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "training frames/sec", "value": a.batch_per_gpu * world * a.steps / (ms * 1e-3),
                          "unit": "frames/s", "n_gpus": world, "steps": a.steps, "ms_per_step": ms / a.steps,
                          "dtype": "f32", "gpu_launches": _lib.launch_count - l0,
                          "losses": {k: float(v) for k, v in losses.items()},
                          "config": {"workload": "BASELINE configs[4] semantics: SSL step, %d frames per GPU x 5 views "
                                                 "3x384x288, %d proposals, float32 training path" % (a.batch_per_gpu, a.proposals)}}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
