"""One-process-per-GPU sharding of the path, over ``torch.distributed``.

Two shardings, one exchange step each:

* **views -> ranks for the root grid** (BASELINE config 4).  Each rank holds the heat-maps of a contiguous range
  of views, un-projects them into a PARTIAL root grid (raw numerators + view count, ``sp3d_unproject_fwd`` with
  ``partial = 1``), the partial grids are summed with ONE all-reduce (``[B, C+1, X*Y*Z]`` float32: 8.2 MB at B = 8,
  C = 1), and ``sp3d_unproject_finalize`` applies the reference's ``sum / (count + 1e-6)``, NaN -> 0, clamp
  (``lib/models/project_layer.py:96-99``).  V2V-root and the proposal layer then run replicated.
* **(sample, proposal) cubes -> ranks for the pose net.**  Person cubes are independent units; the heat-maps
  (16.6 MB at B = 8) are broadcast view by view from their owners, each rank regresses its slice of the valid
  cubes, and the ``[n, J, 3]`` results are all-gathered.

``infer_image_sharded`` is the balanced form of the same split that ``bench.py --gpus N`` measures: the flattened
(view, sample) image list is sharded instead of whole views (5 views leave 3 of 8 GPUs idle), the root-grid sum is a
reduce-scatter over samples, the heat-maps travel in ONE all-gather that overlaps the root path.

The partition arithmetic is pure Python (tested under ``gloo`` on CPU); the collectives are NCCL on GPUs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def view_range(rank, world, num_views):
    """Contiguous block of views owned by ``rank`` (first ranks get the remainder; may be empty)."""
    base, rem = divmod(num_views, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def view_owner(view, world, num_views):
    for r in range(world):
        b, e = view_range(r, world, num_views)
        if b <= view < e:
            return r
    raise ValueError(view)


def shard_slice(n, rank, world):
    """Contiguous slice of ``n`` independent units owned by ``rank``."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_rows(local, counts, group=None):
    """All-gather row blocks of different lengths: ``local [counts[rank], ...]`` -> ``[sum(counts), ...]``."""
    world = dist.get_world_size(group)
    width = max(counts) if counts else 0
    tail = tuple(local.shape[1:])
    padded = torch.zeros((width,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * width,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width:r * width + counts[r]] for r in range(world)], dim=0)


def broadcast_views(local_heatmaps, num_views, shape, device, group=None):
    """Every rank ends up with all ``num_views`` heat-maps ``[B,J,h,w]``; view v comes from its owner rank.
    ``local_heatmaps``: dict {view index: tensor} of the views this rank computed."""
    world = dist.get_world_size(group)
    B, J, h, w = [int(s) for s in shape]
    pitch = ops.round_up(J, 4)
    out = []
    for v in range(num_views):
        owner = view_owner(v, world, num_views)
        # exchanged channel-last (the layout the un-projection kernel gathers from with 16-byte taps)
        buf = torch.zeros(B, h, w, pitch, device=device, dtype=torch.float32)
        if v in local_heatmaps:
            buf[..., :J] = local_heatmaps[v].permute(0, 2, 3, 1)
        dist.broadcast(buf, src=owner, group=group)
        out.append(buf.permute(0, 3, 1, 2)[:, :J])
    return out


def root_volume_view_sharded(root_net, local_heatmaps, meta, batch_size, group=None, cams=None):
    """View-sharded replacement of ``CuboidProposalNet.root_volume``.

    ``local_heatmaps``: dict {view index: ``[B,J,h,w]`` CUDA tensor} for the views in this rank's ``view_range``;
    ``meta``: the full list over views (camera parameters are tiny and replicated).  Returns the full
    ``root_cubes [B,X,Y,Z]`` on every rank.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    V = len(meta)
    v0, v1 = view_range(rank, world, V)
    pl = root_net.project_layer
    device = torch.device("cuda", torch.cuda.current_device())
    if cams is None:
        cams = ops.pack_cameras(meta, pl.img_size).to(device, non_blocking=True)
    centers, _ = pl.centers_tensor([list(root_net.grid_center)], batch_size, device)
    X, Y, Z = [int(s) for s in root_net.cube_size]
    n_vox = X * Y * Z
    if root_net.rootnet_roothm:
        views = {v: h[:, root_net.root_id:root_net.root_id + 1] for v, h in local_heatmaps.items()}
        C = 1
    else:
        views = dict(local_heatmaps)
        C = int(next(iter(views.values())).shape[1]) if views else int(root_net.v2v_net.input_channels)
    partial = torch.zeros(batch_size, C + 1, n_vox, device=device, dtype=torch.float32)
    if v1 > v0:
        ref = views[v0]
        hm_list = [views.get(v, ref) for v in range(V)]      # only [v0, v1) is dereferenced by the kernel
        strides = ref.stride()
        if any(views[v].stride() != strides for v in range(v0, v1)):
            hm_list = [h.contiguous() for h in hm_list]
            strides = hm_list[0].stride()
        ops.unproject(hm_list, strides, cams, centers, root_net.grid_size, (X, Y, Z), pl.img_size,
                      tuple(ref.shape[2:]), C, partial, ((C + 1) * n_vox, n_vox, 1), view_range=(v0, v1),
                      partial=True, heatmap_cfg_wh=pl.heatmap_size)
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)          # the voxel-grid exchange
    ops.unproject_finalize(partial, batch_size, C, n_vox, ((C + 1) * n_vox, n_vox, 1))
    cubes = partial[:, :C].reshape(batch_size, C, X, Y, Z)
    if ops.volume_dtype() == torch.bfloat16:
        cl = ops.to_channel_last(cubes, c_pitch=ops.round_up(C, 16), dtype=torch.bfloat16)
    elif ops.use_split():
        cl = ops.split_act(ops.to_channel_last(cubes), C)
    else:
        cl = ops.to_channel_last(cubes)
    root = root_net.v2v_net.forward_cl(cl, out_pitch=1)
    return root.view(batch_size, X, Y, Z)


def regress_sharded(pose_net, all_heatmaps, cams, grid_centers, pred, group=None):
    """(sample, proposal)-sharded replacement of ``_inference.regress_valid``: fills ``pred[..., 0:3]``."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    valid = torch.nonzero(grid_centers[:, :, 3] >= 0)
    n = int(valid.shape[0])
    if n == 0:
        return pred
    counts = [shard_slice(n, r, world)[1] - shard_slice(n, r, world)[0] for r in range(world)]
    b, e = shard_slice(n, rank, world)
    mine = valid[b:e]
    centers = grid_centers[mine[:, 0], mine[:, 1]].contiguous()
    J = pose_net.num_joints
    if e > b:
        local = pose_net.regress(all_heatmaps, cams, centers, mine[:, 0].to(torch.int32).contiguous())
    else:
        local = torch.zeros(0, J, 3, device=grid_centers.device)
    joints = all_gather_rows(local, counts, group=group)
    pred[valid[:, 0], valid[:, 1], :, 0:3] = joints
    return pred


def infer_view_sharded(model, local_views, meta, group=None):
    """Whole inference with the backbone sharded by view, the root grid exchanged by all-reduce and the person
    cubes sharded by (sample, proposal).  ``local_views``: dict {view index: ``[B,3,H,W]``} for this rank's views.
    Returns ``(pred, all_heatmaps, grid_centers)`` identical on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    V = len(meta)
    device = torch.device("cuda", torch.cuda.current_device())
    B = int(meta[0]["center"].shape[0])
    local_hm = {}
    if local_views:
        order = sorted(local_views)
        hm = model.backbone(torch.cat([local_views[v] for v in order], dim=0))
        for i, v in enumerate(order):
            local_hm[v] = hm[i * B:(i + 1) * B]
    img_size = model.root_net.project_layer.img_size
    cams = ops.pack_cameras(meta, img_size).to(device, non_blocking=True)
    root_cubes = root_volume_view_sharded(model.root_net, local_hm, meta, B, group=group, cams=cams)
    grid_centers = model.root_net.proposal_layer(root_cubes, meta)
    K, J = model.num_cand, model.num_joints
    h, w = int(model.root_net.project_layer.heatmap_size[1]), int(model.root_net.project_layer.heatmap_size[0])
    if local_hm:
        shape = tuple(next(iter(local_hm.values())).shape)
    else:
        shape = (B, J, h, w)
    all_heatmaps = broadcast_views(local_hm, V, shape, device, group=group)
    pred = torch.zeros(B, K, J, 5, device=device)
    pred[:, :, :, 3:] = grid_centers[:, :, 3:].reshape(B, -1, 1, 2)
    regress_sharded(model.pose_net, all_heatmaps, cams, grid_centers, pred, group=group)
    return pred, all_heatmaps, grid_centers


# ------------------------------------------------------------------------------------------------ image sharding
def image_shard(rank, world, num_views, batch_size):
    """Balanced sharding of the flattened (view, sample) image list (index ``v * B + i``): the contiguous slice
    owned by ``rank`` and, per view, the ``[begin, end)`` sample range of it that falls into that slice.
    40 images (5 views x 8 frames) give 20 / 10 / 5 images per rank on 2 / 4 / 8 GPUs (view sharding would leave 3 of
    8 GPUs without backbone work)."""
    b, e = shard_slice(num_views * batch_size, rank, world)
    per_view = {}
    for v in range(num_views):
        lo, hi = max(b, v * batch_size), min(e, (v + 1) * batch_size)
        if hi > lo:
            per_view[v] = (lo - v * batch_size, hi - v * batch_size)
    return (b, e), per_view


def infer_image_sharded(model, local_images, meta, group=None, side_stream=None):
    """BASELINE configs[3] on ``world`` GPUs: ONE batch strong-scaled.

    * backbone: the flattened (view, sample) image list is sharded (``image_shard``); ``local_images`` is this rank's
      ``[n_local, 3, H, W]`` slice;
    * root grid: every rank un-projects ITS images into partial numerators + view counts (``sp3d_unproject_fwd`` with
      ``partial = 1``, one launch per owned view over the owned samples); ONE ``reduce_scatter`` (sum) of the
      ``[B, C+1, X*Y*Z]`` float32 grid over NVLink leaves each rank with the summed grid of its samples (the
      all-reduce of the north star, of which every rank keeps only the slice it goes on to use);
      ``sp3d_unproject_finalize`` -> V2V-root -> NMS on ``B / world`` samples per rank; the ``[B, K, 5]`` proposals are
      all-gathered (1.6 KB);
    * heat-maps: ONE ``all_gather`` of the channel-last maps (16.6 MB at B = 8), issued on ``side_stream`` so that it
      overlaps the root path;
    * person cubes: sharded by (sample, proposal); the ``[n, J, 3]`` joints are all-gathered.

    Returns ``(pred, all_heatmaps, grid_centers)``, identical on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    V = len(meta)
    B = int(meta[0]["center"].shape[0])
    device = torch.device("cuda", torch.cuda.current_device())
    (ib, ie), per_view = image_shard(rank, world, V, B)
    K, J = model.num_cand, model.num_joints
    root_net = model.root_net
    pl = root_net.project_layer
    if (V * B) % world or B % world:
        raise ValueError("image sharding needs V*B and B divisible by the world size (got V=%d B=%d world=%d)" % (V, B, world))

    # ---- backbone on the local images: channel-last float32 heat-maps [n_local, h, w, pitch]
    hm_local = model.backbone(local_images)                        # [n_local, J, h, w] view of a channel-last buffer
    n_local, _, h, w = [int(v) for v in hm_local.shape]
    hm_cl = hm_local.permute(0, 2, 3, 1)                            # [n_local, h, w, J] (stride-1 channels)
    pitch = int(hm_cl.stride(2))
    base = torch.as_strided(hm_cl, (n_local, h, w, pitch), (h * w * pitch, w * pitch, pitch, 1))

    # ---- heat-map all-gather on the side stream (overlaps the root path)
    cur = torch.cuda.current_stream()
    gathered = torch.empty(V * B, h, w, pitch, device=device, dtype=torch.float32)
    produced = torch.cuda.Event()
    produced.record(cur)
    stream = side_stream if side_stream is not None else cur
    with torch.cuda.stream(stream):
        stream.wait_event(produced)
        dist.all_gather_into_tensor(gathered, base.contiguous(), group=group)
        gathered_done = torch.cuda.Event()
        gathered_done.record(stream)

    # ---- partial root grid of the local images
    cams = ops.pack_cameras(meta, pl.img_size).to(device, non_blocking=True)
    X, Y, Z = [int(v) for v in root_net.cube_size]
    n_vox = X * Y * Z
    if root_net.rootnet_roothm:
        C, c0 = 1, int(root_net.root_id)
    else:
        C, c0 = J, 0
    partial = torch.zeros(B, C + 1, n_vox, device=device, dtype=torch.float32)
    centers_all, _ = pl.centers_tensor([list(root_net.grid_center)], B, device)
    for v, (s0, s1) in per_view.items():
        # maps of view v, samples [s0, s1) = rows [first, first + cnt) of the local buffer; the launch sees them as a
        # batch of cnt samples (camera table and centres sliced the same way), summing view v only
        cnt, first = s1 - s0, v * B + s0 - ib
        hm_v = torch.as_strided(base, (cnt, C, h, w), (h * w * pitch, 1, w * pitch, pitch),
                                base.storage_offset() + first * h * w * pitch + c0)
        tmp = torch.empty(cnt, C + 1, n_vox, device=device, dtype=torch.float32)
        ops.unproject([hm_v] * V, hm_v.stride(), cams[s0:s1].contiguous(), centers_all[s0:s1].contiguous(),
                      root_net.grid_size, (X, Y, Z), pl.img_size, (h, w), C, tmp, ((C + 1) * n_vox, n_vox, 1),
                      view_range=(v, v + 1), partial=True, heatmap_cfg_wh=pl.heatmap_size)
        partial[s0:s1] += tmp
    # ---- the voxel-grid exchange: sum over ranks, each rank keeps its B / world samples
    per = B // world
    mine = torch.empty(per, C + 1, n_vox, device=device, dtype=torch.float32)
    if world > 1:
        dist.reduce_scatter_tensor(mine, partial, op=dist.ReduceOp.SUM, group=group)
    else:
        mine.copy_(partial)
    ops.unproject_finalize(mine, per, C, n_vox, ((C + 1) * n_vox, n_vox, 1))
    cubes = mine[:, :C].reshape(per, C, X, Y, Z)
    if ops.volume_dtype() == torch.bfloat16:
        cl = ops.to_channel_last(cubes, c_pitch=ops.round_up(C, 16), dtype=torch.bfloat16)
    elif ops.use_split():
        cl = ops.split_act(ops.to_channel_last(cubes), C)
    else:
        cl = ops.to_channel_last(cubes)
    root = root_net.v2v_net.forward_cl(cl, out_pitch=1).view(per, X, Y, Z)
    gc_mine = root_net.proposal_layer(root, None)                   # [per, K, 5]
    grid_centers = torch.empty(B, K, 5, device=device, dtype=torch.float32)
    if world > 1:
        dist.all_gather_into_tensor(grid_centers, gc_mine.contiguous(), group=group)
    else:
        grid_centers.copy_(gc_mine)

    # ---- person cubes, sharded by (sample, proposal), from ALL heat-maps
    cur.wait_event(gathered_done)
    all_heatmaps = [gathered[v * B:(v + 1) * B].permute(0, 3, 1, 2)[:, :J] for v in range(V)]
    pred = torch.zeros(B, K, J, 5, device=device)
    pred[:, :, :, 3:] = grid_centers[:, :, 3:].reshape(B, -1, 1, 2)
    if world > 1:
        regress_sharded(model.pose_net, all_heatmaps, cams, grid_centers, pred, group=group)
    else:
        from .models._inference import regress_valid
        regress_valid(model.pose_net, all_heatmaps, cams, grid_centers, pred)
    return pred, all_heatmaps, grid_centers


def allreduce_gradients(parameters, group=None, bucket_bytes=64 << 20):
    """Data-parallel gradient exchange of the training path: the gradients of all trainable
    ``parameters`` are averaged over the ranks with one sum all-reduce per bucket of ``bucket_bytes`` (flattened
    float32; 211 MB for the 52.7 M trainable parameters of the SSL model = 4 buckets), NCCL on GPUs.  Parameters
    without a gradient on this rank take part as zeros so that the buckets line up on every rank (the reference trains
    under ``nn.DataParallel``, which sums replica gradients the same way, ``tools/train_3d.py:75``)."""
    world = dist.get_world_size(group)
    params = [p for p in parameters if p.requires_grad]
    if world == 1 or not params:
        return
    bucket, size = [], 0
    buckets = []
    for p in params:
        n = p.numel() * 4
        if bucket and size + n > bucket_bytes:
            buckets.append(bucket)
            bucket, size = [], 0
        bucket.append(p)
        size += n
    if bucket:
        buckets.append(bucket)
    for bucket in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, group=group)
        flat /= world
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
            off += p.numel()
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
