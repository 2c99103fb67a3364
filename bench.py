#!/usr/bin/env python
"""Benchmark of the voxelised multi-view pose path (BASELINE.json metric: frames/s).

  python bench.py --gpus N --steps K --warmup W            # this backend (one JSON line on rank 0)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation of the path

Workload (BASELINE.json configs[2], the largest single-GPU configuration): a batch of B = 8 synthetic frames, each 5
views of 3x384x288, PoseResNet-50 -> 80x80x20 root grid -> 10 proposals per frame (all forced valid) -> 64^3 person
cubes -> V2VNet -> soft-argmax.  A "step" is one forward of that batch.

Arithmetic: the float32-faithful tensor-core mode (``--volume-dtype f32x3``, the product's default): float32 values
throughout, convolutions on tcgen05 with both operands split into two bf16 terms and three term pairs accumulated
in float32 -- the mode whose results meet the parity bars of tests/test_gpu_fullsize.py.  The bf16-operand
throughput mode is timed beside it as a labelled side key (it is outside the parity tolerance).

`value` = frames/s with the images already resident in HBM; `e2e` = the same through the public module call with the
images in pinned host memory (H2D inside the timed region) and the predictions read back (D2H).

N > 1 (torchrun, one rank per GPU): BASELINE configs[3] -- the SAME batch of 8 frames strong-scaled over the N GPUs
(selfpose3d_b200/dist.py: the 40 (view, sample) images sharded over the ranks, partial root grids summed with one NCCL
reduce-scatter over NVLink, heat-maps exchanged with one all-gather that overlaps the root path, person cubes sharded
by (sample, proposal), joints all-gathered); the independent-replica number (one batch per GPU, no data-path
collective) is reported beside it as `replicas`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 8
VIEWS = 5
PROPOSALS = 10
IMAGE_SIZE = [288, 384]      # [w, h]
HEATMAP_SIZE = [72, 96]


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r01_ncu_*_summary.csv, profiles/r02b_ncu_*_summary.csv, profiles/r02c_ncu_stem_*), keyed by the launch's layer string; refreshed by hand with the captures
NCU_EVIDENCE = {
    # float32-faithful mode, 7^3 stem over 80 cubes: 2 bf16 term planes in (2 x 671 MB) + 2 planes out (2 x 671 MB)
    # algorithmic; measured 1.3497 GB read + 1.2976 GB written
    "conv algo2 k7 15->16 @80x64x64x64": {
        "dram_bytes_per_launch": 1.349722e9 + 1.297612e9,
        "source": "profiles/r02c_ncu_stem_80cubes_summary.csv (conv_tc_kernel<7,7,128,64,4,5,3,1,2,32,2,F=4,WD=2>, ncu --set "
                  "full: sm__pipe_tensor_cycles_active 60.4 %)"},
    "conv algo1 k7 15->16 @80x64x64x64": {
        "dram_bytes_per_launch": 672.02e6 + 634.73e6,   # bf16 throughput mode: 671 MB bf16 cubes in + 671 MB out
        "source": "profiles/r01_ncu_conv_pose_v11_summary.csv (conv_tc_kernel<7,7,64,32,...,F=2>, ncu --set full)"},
    # un-projection of the 80 person cubes, float32 arithmetic, term-pair output: 7.7 MB read (maps live in L2) + 1.2828 GB
    # written against 1.2749 GB algorithmic + the zero padding channel
    "unproject": {
        "dram_bytes_per_launch": 0.007710e9 + 1.282782e9,
        "source": "profiles/r02b_ncu_k1_k3_k4_summary.csv (unproject_kernel<1, bf16, PAIR>, 80 cubes, ncu --set full)"},
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def make_cfg(batch):
    from selfpose3d_b200.config import default_config
    cfg = default_config()
    cfg.NETWORK.IMAGE_SIZE, cfg.NETWORK.HEATMAP_SIZE = list(IMAGE_SIZE), list(HEATMAP_SIZE)
    cfg.MULTI_PERSON.MAX_PEOPLE_NUM = PROPOSALS
    cfg.MULTI_PERSON.THRESHOLD = -1e9       # untrained-like nets score low: force every proposal slot valid
    cfg.TEST.BATCH_SIZE = batch
    return cfg


def oracle_cfg(cfg):
    return dict(image_size=cfg.NETWORK.IMAGE_SIZE, heatmap_size=cfg.NETWORK.HEATMAP_SIZE,
                space_size=cfg.MULTI_PERSON.SPACE_SIZE, space_center=cfg.MULTI_PERSON.SPACE_CENTER,
                initial_cube_size=cfg.MULTI_PERSON.INITIAL_CUBE_SIZE, grid_size=cfg.PICT_STRUCT.GRID_SIZE,
                cube_size=cfg.PICT_STRUCT.CUBE_SIZE, max_people=cfg.MULTI_PERSON.MAX_PEOPLE_NUM,
                threshold=cfg.MULTI_PERSON.THRESHOLD, beta=cfg.NETWORK.BETA, root_idx=cfg.DATASET.ROOTIDX)



def cpu_inputs(frames, seed=0):
    """State dict, oracle configuration and one synthetic batch of ``frames`` frames for the CPU legs."""
    from selfpose3d_b200 import synthetic
    from selfpose3d_b200.models import multi_person_posenet_ssv
    cfg = make_cfg(frames)
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    sd = synthetic.trained_like_state_dict(model, seed=0)
    cams = synthetic.ring_cameras(VIEWS, seed=0)
    meta = synthetic.make_meta(cams, frames, IMAGE_SIZE)
    images = synthetic.random_images(frames, VIEWS, IMAGE_SIZE, seed=seed)
    return cfg, sd, meta, images


def port_args(cfg, sd, meta):
    cam_arrays = {k: np.stack([m["camera"][k].numpy() for m in meta]) for k in meta[0]["camera"]}
    return (sd, oracle_cfg(cfg), cam_arrays, [m["center"].numpy() for m in meta],
            [m["scale"].numpy() for m in meta], [m["rotation"].numpy() for m in meta])


def cpu_reference_frames_per_s(steps, warmup, frames_per_step=1, kind=None, images=None):
    """Times the reference's CPU implementation of the path on `frames_per_step` frames per step: the UNMODIFIED
    reference staged under oracle/_ref (kind "reference": MultiPersonPoseNetSSV.forward(inference=True) through its own
    module API, lib/models/multi_person_posenet_ssv.py:105-153) or, when that is absent, the oracle port
    (oracle/pipeline.py: the same torch CPU calls the reference makes).  Returns (frames/s, s/step, kind, last pred)."""
    from oracle import pipeline, ref_runner
    torch.set_num_threads(os.cpu_count() or 1)
    if kind is None:
        kind = "reference" if ref_runner.available() else "port"
    cfg, sd, meta, own_images = cpu_inputs(frames_per_step)
    images = own_images if images is None else images
    if kind == "reference":
        ocfg = oracle_cfg(cfg)
        model, _ = ref_runner.build_model(state_dict=sd, num_joints=cfg.NETWORK.NUM_JOINTS, **ocfg)

        def run():
            return ref_runner.inference(model, images, meta)[0]
    else:
        args = port_args(cfg, sd, meta)

        def run():
            return pipeline.inference(*args, images=images)[0]
    pred = None
    with torch.no_grad():
        for _ in range(warmup):
            pred = run()
        t0 = time.perf_counter()
        for _ in range(steps):
            pred = run()
        dt = time.perf_counter() - t0
    return frames_per_step * steps / dt, dt / steps, kind, pred


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}



def run_reference(args, rank, world):
    if rank != 0:
        return
    fps, spf, kind, _ = cpu_reference_frames_per_s(args.steps, args.warmup, 1)
    cores = os.cpu_count() or 1
    what = ("the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.sh): MultiPersonPoseNetSSV.forward("
            "inference=True) on torch CPU" if kind == "reference" else "oracle/pipeline.py (CPU port of the reference "
            "path) on torch CPU")
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": spf * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, 1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "1 frame per step (5 views 3x384x288, 10 proposals), %s with %d threads"
                                   % (what, torch.get_num_threads())},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(batch, world):
    cfg = {"workload": "BASELINE configs[%d]: 5-view 3x384x288 synthetic frames, one batch of %d frames per step%s, "
                       "PoseResNet-50 + RootNet (80x80x20) + PoseNet (10 proposals x 64^3), all proposals valid"
                       % (2 if world == 1 else 3, batch, "" if world == 1 else " strong-scaled over %d GPUs" % world),
           "batch": batch, "views": VIEWS, "proposals": PROPOSALS, "image": "3x384x288",
           "root_grid": "80x80x20", "person_cube": "64x64x64",
           "cache": "per-step working set (several GB of activations) exceeds the 126 MB L2; no explicit flush"}
    if world == 1:
        cfg["parallelism"] = "single GPU"
    else:
        cfg["parallelism"] = ("one batch over %d ranks: (view, sample) images sharded %d per rank; root grid = NCCL "
                              "reduce-scatter(sum) of partial numerators + view counts [8,2,128000] f32 (8.2 MB), proposals "
                              "all-gather; heat-maps one all-gather (16.6 MB) overlapping the root path; person cubes "
                              "sharded by (sample, proposal), joints all-gather" % (world, BATCH * VIEWS // world))
    return cfg

def build_training_step(cfg, dev, image_size, views, seed=5):
    """Model, inputs and the step closure of the training-step measurement: supervised ``MultiPersonPoseNet`` step
    (reference lib/models/multi_person_posenet.py:36-102 in .train()) on ONE frame -- frozen backbone, root net and
    pose net (every proposal slot matched to a ground-truth person) forward + backward, gradients only."""
    from selfpose3d_b200 import synthetic
    from selfpose3d_b200.models import multi_person_posenet
    cfg.MODEL = "multi_person_posenet"
    model = multi_person_posenet.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
    model = model.to(dev).train()
    model.backbone.eval()
    cams = synthetic.ring_cameras(views, seed=0)
    meta = synthetic.make_meta(cams, 1, image_size)
    images = [im.to(dev) for im in synthetic.random_images(1, views, image_size, seed=seed)]
    # ground truth next to the proposals the (randomly weighted) root net makes in training mode (batch statistics), so
    # that every slot is matched to a person
    with torch.no_grad():
        _, _, gc, _, _, _ = model(views=images, meta=meta)
    K, J = cfg.MULTI_PERSON.MAX_PEOPLE_NUM, cfg.NETWORK.NUM_JOINTS
    roots = gc[:, :, :3].detach().cpu().double() + 50.0
    meta[0].update(roots_3d=roots, num_person=torch.tensor([K]),
                   joints_3d=roots[:, :, None, :].expand(-1, -1, J, -1).contiguous(),
                   joints_3d_vis=torch.ones(1, K, J, 3, dtype=torch.float64))
    targets_3d = torch.rand(1, *cfg.MULTI_PERSON.INITIAL_CUBE_SIZE, generator=torch.Generator().manual_seed(seed)).to(dev)

    def step():
        model.zero_grad(set_to_none=True)
        _, _, grid, _, loss_3d, loss_cord = model(views=images, meta=meta, targets_3d=targets_3d)
        (loss_3d + loss_cord).backward()
        return grid

    return model, step


def time_training_step(dev, steps=2, mode="bf16x3"):
    """Runs in a child process of the bench (``--train-step-only``), so that nothing it does can touch the main
    measurement."""
    from selfpose3d_b200 import ops, _lib
    ops.set_volume_dtype(torch.float32)
    ops.set_float32_conv(mode)
    _, step = build_training_step(make_cfg(1), dev, IMAGE_SIZE, VIEWS)
    grid = step()
    matched = int((grid[:, :, 3] >= 0).sum())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"value": steps / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms / steps, "steps": steps,
            "gpu_launches": _lib.launch_count - l0, "matched_proposals": matched,
            "what": "supervised step, 1 frame x 5 views, frozen backbone, root net + pose net forward and backward; "
                    "float32 training path, convolutions: %s" % ("float32 FMA kernels" if mode == "simt" else
                    "forward, covered input gradients and all weight gradients on tcgen05 (split operands); all "
                    "proposal slots in one pose-net pass (grouped BatchNorm statistics keep the per-slot batches)")}



DTYPE_DETAIL = {
    "bf16": "convolutions: bf16 operands, float32 accumulation (tcgen05); un-projection geometry, NMS, soft-argmax: "
            "float32.  Throughput mode, OUTSIDE the parity tolerance",
    "f32": "float32 everywhere (FMA convolutions)",
    "f32x3": "float32 values end to end; convolutions on tcgen05 with both operands split into 2 bf16 terms (3 term "
             "pairs), float32 accumulation; activations travel between layers as the two bf16 term planes",
    "f32x6": "float32 activations; convolutions on tcgen05 with operands split into 3 bf16 terms (6 term pairs), "
             "float32 accumulation"}


def set_mode(ops, name):
    ops.set_volume_dtype(torch.bfloat16 if name == "bf16" else torch.float32)
    ops.set_float32_conv({"f32": "simt", "f32x3": "bf16x3", "f32x6": "bf16x6", "bf16": "bf16x3"}[name])


def check_against_oracle(pred0, hm0, gc0, images0, mode):
    """Outside every timed region: frame 0 of the batch against oracle/pipeline.py (float32 and float64) on the same
    images -- the bars of tests/test_gpu_fullsize.py, reported in the JSON line (never fatal for the measurement)."""
    from oracle import pipeline
    cfg, sd, meta, _ = cpu_inputs(1)
    args = port_args(cfg, sd, meta)
    t0 = time.perf_counter()
    with torch.no_grad():
        p32, h32, g32, _ = pipeline.inference(*args, images=images0)
        port_s = time.perf_counter() - t0
        p64, h64, _, _ = pipeline.inference(*args, images=images0, dtype=torch.float64, grid_centers=g32)
    scale = max(float(h.abs().max()) for h in h32)
    hm_err = max(float((a - b).abs().max()) for a, b in zip(hm0, h32)) / scale
    same = (gc0[..., :3] - g32[..., :3]).abs().amax(-1) <= 1e-3
    j_our = float((pred0[..., :3].double() - p64[..., :3])[same].abs().max()) if bool(same.any()) else None
    j_ref = float((p32[..., :3].double() - p64[..., :3])[same].abs().max()) if bool(same.any()) else None
    ok = bool(hm_err <= 1e-4 and bool(same.all()) and j_our is not None and j_our <= max(1.5 * j_ref, 1e-3))
    return {"frame": 0, "mode": mode, "heatmaps_rel_err_vs_oracle_f32": hm_err,
            "proposals_mismatched": int((~same).sum()), "proposals": int(same.numel()),
            "joints_mm_ours_vs_f64": j_our, "joints_mm_oracle_f32_vs_f64": j_ref,
            "joints_mm_ours_vs_oracle_f32": float((pred0[..., :3] - p32[..., :3])[same].abs().max()) if bool(same.any()) else None,
            "bars": "heat-maps <= 1e-4 of range; proposals identical; joints |ours-f64| <= max(1.5 |oracle_f32-f64|, 1e-3 mm)",
            "pass": ok, "oracle_port_s_per_frame": port_s}, p32


def run_ours(args, rank, world, local_rank):
    from selfpose3d_b200 import synthetic, _lib, ops
    from selfpose3d_b200 import dist as sd
    from selfpose3d_b200.models import multi_person_posenet_ssv
    import selfpose3d_b200.profiler as prof

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.load()
    set_mode(ops, args.volume_dtype)
    cfg = make_cfg(BATCH)
    model = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    model.load_state_dict(synthetic.trained_like_state_dict(model, seed=0), strict=True)
    model = model.to(dev).eval()
    cams = synthetic.ring_cameras(VIEWS, seed=0)
    meta = synthetic.make_meta(cams, BATCH, IMAGE_SIZE)
    # N = 1: the whole batch.  N > 1: ONE batch (the same on every rank, seed 0) of which this rank owns a slice of the
    # flattened (view, sample) image list; `replica` = a batch of its own per rank (the side measurement)
    images = synthetic.random_images(BATCH, VIEWS, IMAGE_SIZE, seed=0)
    if world > 1:
        (ib, ie), _ = sd.image_shard(rank, world, VIEWS, BATCH)
        flat = torch.cat(images, dim=0)                          # [V*B, 3, H, W], index v*B + i
        host_images = [flat[ib:ie].clone().pin_memory()]
        side = torch.cuda.Stream(device=dev)
    else:
        host_images = [im.pin_memory() for im in images]
    dev_images = [im.to(dev) for im in host_images]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def forward(imgs):
        if world > 1:
            return sd.infer_image_sharded(model, imgs[0], meta, side_stream=side)
        return model(views1=imgs, meta1=meta, inference=True)

    def step_resident():
        return forward(dev_images)[0]

    # end-to-end step through the public call: every step's images come from pinned host memory and the predictions
    # go back to the host.  The copy of step i+1's images runs on a side stream while step i computes (two device
    # buffers); it is still one full H2D per step inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [[torch.empty_like(im, device=dev) for im in host_images] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the step that read this slot has finished
            for d, h in zip(slots[slot], host_images):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["i"] & 1
        if not state["primed"]:
            consumed[0].record()
            consumed[1].record()
            upload(cur)
            state["primed"] = True
        upload(cur ^ 1)                                     # next step's images, overlapping this step's kernels
        torch.cuda.current_stream().wait_event(ready[cur])
        pred = forward(slots[cur])[0]
        consumed[cur].record()
        state["i"] += 1
        return pred.cpu() if rank == 0 else pred

    def timed(fn, steps, profile=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count
        t0 = time.perf_counter()
        e0.record()
        if profile:
            prof.enable()
        for _ in range(steps):
            fn()
        if profile:
            prof.disable()
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launch_count - l0, t0, t1

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, t0, t1 = timed(step_resident, args.steps)
    step_e2e()
    ms_e2e, _, _, t1 = timed(step_e2e, args.steps)
    clocks = sampler.stop(t0, t1) if sampler else None
    # per-kernel-family device time (CUDA events around every C-ABI launch, on the launching stream) over a few
    # extra steps: the roofline numerators/denominators; kept out of the two timed regions above
    prof_steps = min(args.steps, 5)
    ms_prof, _, _, _ = timed(step_resident, prof_steps, profile=True)
    kernels = prof.summary()
    layers = prof.detail_summary()

    # N > 1: the independent-replica measurement (one whole batch per rank, no data-path collective) beside the split
    replicas = None
    if world > 1:
        rep_images = [im.to(dev) for im in synthetic.random_images(BATCH, VIEWS, IMAGE_SIZE, seed=rank)]

        def step_replica():
            return model(views1=rep_images, meta1=meta, inference=True)[0]
        for _ in range(3):
            step_replica()
        n_rep = min(args.steps, 10)
        ms_rep, _, _, _ = timed(step_replica, n_rep)
        replicas = {"value": BATCH * world * n_rep / (ms_rep * 1e-3), "unit": "frames/s", "ms_per_step": ms_rep / n_rep,
                    "steps": n_rep, "scaling": "weak",
                    "what": "one batch of %d frames per rank, frames independent, no data-path collective" % BATCH}

    # outputs of the timed configuration for the check leg (frame 0)
    check = None
    with torch.no_grad():
        pred, hms, gc = forward(dev_images)
    pred0, gc0 = pred[:1].cpu(), gc[:1].cpu()
    hm0 = [h[:1].float().cpu() for h in hms]

    if rank != 0:
        return
    peaks = load_peaks()
    frames = BATCH * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d = sum(int(im.numel()) * 4 for im in host_images) * world     # every rank uploads its shard: the whole batch
    d2h = BATCH * PROPOSALS * cfg.NETWORK.NUM_JOINTS * 5 * 4

    # roofline of the DOMINANT KERNEL: the single launch type with the largest share of the step (the 7^3 V2V stem
    # over the person cubes), Sum algorithmic FLOPs / Sum CUDA-event time of its launches; the whole convolution
    # family and the un-projection (HBM) are reported beside it
    conv = kernels.get("conv", {"ms": 0.0, "work": 0.0, "launches": 0})
    unp = kernels.get("unproject", {"ms": 0.0, "work": 0.0, "launches": 0})
    top_name, top = max(layers.items(), key=lambda kv: kv[1]["ms"]) if layers else ("", {"ms": 0.0, "work": 0.0, "launches": 0})
    top_tflops = top["work"] / (top["ms"] * 1e-3) / 1e12 if top["ms"] > 0 else 0.0
    conv_tflops = conv["work"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    unp_gbs = unp["work"] / (unp["ms"] * 1e-3) / 1e9 if unp["ms"] > 0 else 0.0
    split = args.volume_dtype in ("f32x3", "f32x6")
    mma_per_flop = {"f32x3": 3.0, "f32x6": 6.0}.get(args.volume_dtype, 1.0)
    ev = NCU_EVIDENCE.get(top_name, {})
    roofline = {"kernel": "sp3d_conv_fwd / conv_tc_kernel: " + top_name, "bound": "tensor", "achieved": top_tflops,
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": top_tflops / peaks["bf16_tflops_sustained"],
                "traffic": ev.get("dram_bytes_per_launch"), "traffic_source": ev.get("source"),
                "peak_source": peaks["source"] + " (sustained bf16: kernels timed inside a long step)",
                "launches_per_step": top["launches"] / prof_steps,
                "avg_launch_ms": top["ms"] / max(top["launches"], 1),
                "share_of_step": top["ms"] / ms_prof if ms_prof > 0 else None,
                "flops_per_launch": top["work"] / max(top["launches"], 1),
                "note": ("achieved = ALGORITHMIC FLOPs (2 x voxels x Cin x Cout x k^3 of the float32 convolution) / time; the "
                         "float32-faithful mode issues %g bf16 tensor-core MACs per algorithmic MAC, i.e. the tensor pipe "
                         "runs at achieved_bf16_mma_tflops" % mma_per_flop) if split else None,
                "achieved_bf16_mma_tflops": top_tflops * mma_per_flop}
    roofline_conv_family = {"kernel": "sp3d_conv_fwd (all convolution launches of the step)", "bound": "tensor",
                            "achieved": conv_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                            "frac": conv_tflops / peaks["bf16_tflops_sustained"],
                            "achieved_bf16_mma_tflops": conv_tflops * mma_per_flop,
                            "launches_per_step": conv["launches"] / prof_steps,
                            "share_of_step": conv["ms"] / ms_prof if ms_prof > 0 else None}
    ev = NCU_EVIDENCE.get("unproject", {})
    roofline_unproject = {"kernel": "sp3d_unproject_fwd (person cubes + root grid)", "bound": "hbm",
                          "achieved": unp_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": unp_gbs / peaks["hbm_gbs"],
                          "traffic": ev.get("dram_bytes_per_launch"),
                          "traffic_source": ev.get("source"), "peak_source": peaks["source"],
                          "launches_per_step": unp["launches"] / prof_steps,
                          "share_of_step": unp["ms"] / ms_prof if ms_prof > 0 else None,
                          "note": "achieved = SURVEY 8(d) algorithmic bytes (float32 cubes written once + float32 maps "
                                  "read once) / time; float32 view accumulation, the cube leaves the kernel as two bf16 "
                                  "term planes (the same 4 bytes per value)"}

    # frame 0 of the timed configuration against the float32 / float64 oracle (outside the timed regions)
    port_pred = None
    if world == 1 and not args.no_check:
        try:
            check, port_pred = check_against_oracle(pred0, hm0, gc0, [im[:1] for im in images], args.volume_dtype)
        except Exception as exc:   # noqa: BLE001 -- the check must not take the bench line down
            check = {"error": repr(exc)[:300]}

    # the bf16-operand throughput mode on the same resident batch, beside the headline: N = 1 only, never fatal
    side_modes = None
    if world == 1 and not args.no_side_modes:
        side_modes = {}
        for name in ("bf16",):
            if name == args.volume_dtype:
                continue
            try:
                set_mode(ops, name)
                for _ in range(3):
                    step_resident()
                n = min(args.steps, 10)
                ms_f, launches_f, _, _ = timed(step_resident, n)
                side_modes[name] = {"value": BATCH * n / (ms_f * 1e-3), "unit": "frames/s", "ms_per_step": ms_f / n,
                                    "steps": n, "gpu_launches": launches_f, "dtype_detail": DTYPE_DETAIL[name]}
            except Exception as exc:   # noqa: BLE001
                side_modes[name] = {"error": repr(exc)[:300]}
            finally:
                set_mode(ops, args.volume_dtype)

    # one supervised training step (forward + backward of root net and pose net, heat-maps from the frozen backbone) on
    # ONE frame: N = 1 only, in a child process, never fatal
    train_step = None
    if world == 1 and not args.no_train_step:
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--train-step-only"], capture_output=True,
                                 text=True, timeout=240, env=env)
            lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            train_step = json.loads(lines[-1]) if lines else {"error": (res.stderr or "no output")[-300:]}
        except Exception as exc:   # noqa: BLE001
            train_step = {"error": repr(exc)[:300]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_frames = 2
        cpu_fps, cpu_spf, kind, ref_pred = cpu_reference_frames_per_s(cpu_frames, 1, 1, images=[im[:1] for im in images])
        cpu_baseline = {"value": cpu_fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": kind,
                        "sample": "%d x 1 frame (5 views 3x384x288, 10 proposals) through %s on torch CPU (all host "
                                  "threads), 1 warm-up, %.2f s per frame"
                                  % (cpu_frames, "the unmodified reference (oracle/_ref)" if kind == "reference"
                                     else "oracle/pipeline.py", cpu_spf)}
        if check is not None and "oracle_port_s_per_frame" in check:
            cpu_baseline["oracle_port_s_per_frame"] = check["oracle_port_s_per_frame"]
            if kind == "reference" and port_pred is not None:
                # the same frame through the port and through the reference itself: the port IS the reference's arithmetic
                cpu_baseline["port_vs_reference_joints_mm"] = float((ref_pred[..., :3] - port_pred[..., :3]).abs().max())
    line = {
        "metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "f32": "f32", "f32x3": "f32 (bf16x3 split)",
                                       "f32x6": "f32 (bf16x6 split)"}[args.volume_dtype],
        "dtype_detail": DTYPE_DETAIL[args.volume_dtype],
        "data": "synthetic", "config": workload_config(BATCH, world),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_conv_family": roofline_conv_family,
        "roofline_unproject": roofline_unproject,
        "kernel_ms_per_step": {k: v["ms"] / prof_steps for k, v in kernels.items()},
        "check": check,
        "replicas": replicas,
        "side_modes": side_modes,
        "train_step": train_step,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--volume-dtype", default="f32x3", choices=["f32", "bf16", "f32x3", "f32x6"],
                    help="f32x3 (default) = float32-faithful tensor-core mode: operands split into 2 bf16 terms, 3 term "
                         "pairs, float32 accumulation; f32x6 = 3 terms / 6 pairs; bf16 = bf16 operands (throughput mode, "
                         "outside the parity tolerance); f32 = float32 FMA convolutions")
    ap.add_argument("--no-side-modes", action="store_true", help="skip the bf16 throughput-mode side measurement")
    ap.add_argument("--no-train-step", action="store_true", help="skip the training-step side measurement")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle check of frame 0")
    ap.add_argument("--train-step-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""      # the reference arm is the CPU path: no device is touched
        run_reference(args, rank, world)
        return
    if args.train_step_only:
        torch.cuda.set_device(0)
        print(json.dumps(time_training_step(torch.device("cuda", 0))), flush=True)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION and above (WARN included); stdout carries
        # exactly one JSON line here, so anything below INFO is switched off
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ.pop("NCCL_DEBUG", None)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
