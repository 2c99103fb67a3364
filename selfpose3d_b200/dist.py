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
