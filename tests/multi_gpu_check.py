#!/usr/bin/env python
"""N-GPU equality check (run on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py
Every rank builds the same seeded model and frame batch; the view-sharded / cube-sharded inference
(selfpose3d_b200/dist.py: one NCCL all-reduce on the root grid, heat-map broadcast, joints all-gather) must
reproduce the single-GPU result computed locally on each rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from selfpose3d_b200 import dist as sd, synthetic  # noqa: E402
from selfpose3d_b200.config import default_config  # noqa: E402
from selfpose3d_b200.models import multi_person_posenet_ssv  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = default_config()
    cfg.NETWORK.NUM_JOINTS = 4
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = [96, 128], [24, 32]
    cfg.MULTI_PERSON.INITIAL_CUBE_SIZE = [24, 24, 8]
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = 3
    cfg.MULTI_PERSON.THRESHOLD = -1.0
    cfg.PICT_STRUCT.CUBE_SIZE = [16, 16, 16]
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=5), strict=True)
    model = model.cuda().eval()
    B, V = 2 if world <= 2 else world, 5
    meta = synthetic.make_meta(synthetic.ring_cameras(V, seed=0), B, (96, 128))
    images = [im.cuda() for im in synthetic.random_images(B, V, (96, 128), seed=1)]
    with torch.no_grad():
        want_pred, want_hm, want_gc = model(views1=images, meta1=meta, inference=True)
        v0, v1 = sd.view_range(rank, world, V)
        got_pred, got_hm, got_gc = sd.infer_view_sharded(model, {v: images[v] for v in range(v0, v1)}, meta)
        # the balanced split bench.py --gpus N measures: (view, sample) images sharded, reduce-scatter on the root grid
        ok2 = True
        if (B * V) % world == 0 and B % world == 0:
            (ib, ie), _ = sd.image_shard(rank, world, V, B)
            flat = torch.cat(images, dim=0)
            p2, h2, g2 = sd.infer_image_sharded(model, flat[ib:ie].contiguous(), meta, side_stream=torch.cuda.Stream())
            e2 = (max(float((a - b).abs().max()) for a, b in zip(h2, want_hm)), float((g2 - want_gc).abs().max()),
                  float((p2 - want_pred).abs().max()))
            ok2 = e2[0] <= 1e-5 and e2[1] <= 1e-3 and e2[2] <= 2e-2
            print("rank %d/%d images [%d,%d): |heatmaps| %.2g  |grid_centers| %.2g  |pred| %.2g mm  %s"
                  % (rank, world, ib, ie, e2[0], e2[1], e2[2], "OK" if ok2 else "MISMATCH"), flush=True)
    torch.cuda.synchronize()
    e_hm = max(float((a - b).abs().max()) for a, b in zip(got_hm, want_hm))
    e_gc = float((got_gc - want_gc).abs().max())
    e_pred = float((got_pred - want_pred).abs().max())
    ok = e_hm <= 1e-5 and e_gc <= 1e-3 and e_pred <= 2e-2 and ok2
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("rank %d/%d views [%d,%d): |heatmaps| %.2g  |grid_centers| %.2g  |pred| %.2g mm  %s"
          % (rank, world, v0, v1, e_hm, e_gc, e_pred, "OK" if ok else "MISMATCH"), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
