"""ProjectLayer -- multi-view un-projection of 2-D heat-maps into a voxel grid.

Same constructor and ``forward(heatmaps, meta, grid_size, grid_center, cube_size,
flip_xcoords=None) -> (cubes [B,C,X,Y,Z], grids [B,N,3])`` contract as the reference's
``lib/models/project_layer.py:14-106``.  The reference loops over samples and views in Python
and issues ~380 ATen ops per (sample, view); here one host pass packs the cameras and the
input affines into a ``[B,V,32]`` table and ONE launch of ``sp3d_unproject_fwd`` does the rest
(``csrc/unproject.cu``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _common_strides(heatmaps):
    """Views must share one stride tuple for the kernel; make contiguous copies otherwise."""
    s0 = heatmaps[0].stride()
    if all(h.stride() == s0 and h.dtype == torch.float32 for h in heatmaps):
        return list(heatmaps), s0
    hm = [h.float().contiguous() for h in heatmaps]
    return hm, hm[0].stride()


_CENTER_CACHE = {}


class ProjectLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.img_size = cfg.NETWORK.IMAGE_SIZE          # [w, h]
        self.heatmap_size = cfg.NETWORK.HEATMAP_SIZE    # [w, h]

    def compute_grid(self, boxSize, boxCenter, nBins, device=None):
        """Voxel-centre coordinates ``[N,3]`` (x slowest, z fastest) -- reference :22-40."""
        if isinstance(boxSize, (int, float)):
            boxSize = [boxSize] * 3
        if isinstance(nBins, int):
            nBins = [nBins] * 3
        axes = [torch.linspace(-boxSize[a] / 2, boxSize[a] / 2, int(nBins[a])) + float(boxCenter[a]) for a in range(3)]
        gx, gy, gz = torch.meshgrid(axes[0], axes[1], axes[2], indexing="ij")
        grid = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], dim=1)
        return grid.to(device) if device is not None else grid

    def centers_tensor(self, grid_center, batch_size, device):
        """``[[x,y,z]]`` (shared) or ``[B,>=3]`` -> float32 ``[B, k]`` on ``device`` and the skip-flag switch."""
        if isinstance(grid_center, torch.Tensor):
            gc = grid_center.to(device=device, dtype=torch.float32)
        else:
            # configuration constants: uploaded once per device (also keeps the call capturable in a CUDA graph)
            arr = np.asarray(grid_center, dtype=np.float32)
            key = (arr.shape, arr.tobytes(), str(device))
            gc = _CENTER_CACHE.get(key)
            if gc is None:
                if len(_CENTER_CACHE) > 64:
                    _CENTER_CACHE.clear()
                gc = _CENTER_CACHE[key] = torch.as_tensor(arr, device=device)
        if gc.dim() == 1:
            gc = gc[None]
        check = gc.shape[1] != 3                      # reference :54 -- rows with [3] < 0 are skipped
        if gc.shape[0] == 1 and batch_size > 1:       # reference :58-61 -- one centre for the whole batch
            gc = gc.expand(batch_size, gc.shape[1])
        return gc.contiguous(), check

    def project_cl(self, heatmaps, cams, centers, check_flag, grid_size, cube_size, channels=None,
                   cubes_per_sample=1, cube_sample=None, want_grids=False, dtype=torch.float32, c_pitch=None,
                   hms_f16=None):
        """Layout-native un-projection: returns channel-last cubes ``[n_cubes,X,Y,Z,pitch]``
        (padding channels zero) and optionally ``grids [n_cubes,N,3]``.

        ``heatmaps``: list[V] of ``[B,C,h,w]`` CUDA tensors (any common strides); ``cams``: packed
        ``[B,V,32]`` CUDA table from ``ops.pack_cameras``; ``centers``: ``[n_cubes,>=3]`` CUDA float32.
        """
        hms, st = _common_strides(heatmaps)
        C = int(hms[0].shape[1]) if channels is None else int(channels)
        X, Y, Z = [int(s) for s in cube_size]
        n_cubes = int(centers.shape[0])
        pitch = ops.round_up(C, 4) if c_pitch is None else int(c_pitch)
        dev = hms[0].device
        if dtype == "split":
            # float32-faithful tensor-core mode: the float32 form of the kernel (bit-faithful geometry, float32 view
            # accumulation) writes its result as the two bf16 term planes the convolutions read
            pitch = ops.split_pitch(C) if c_pitch is None else int(c_pitch)
            planes = torch.empty(2, n_cubes, X, Y, Z, pitch, device=dev, dtype=torch.bfloat16)
            ops.unproject(hms, st, cams, centers, grid_size, (X, Y, Z), self.img_size, tuple(hms[0].shape[2:]), C,
                          planes, (X * Y * Z * pitch, 1, pitch), out_c_pad=pitch, check_flag=check_flag,
                          cubes_per_sample=cubes_per_sample, cube_sample=cube_sample,
                          heatmap_cfg_wh=self.heatmap_size, pair_out=True)
            return ops.SplitAct(planes), None
        cubes = torch.empty(n_cubes, X, Y, Z, pitch, device=dev, dtype=dtype)
        if dtype == torch.bfloat16 and pitch == 16 and 1 <= C <= 16 and not want_grids and Z <= 256:
            # throughput form (bf16 volume mode): fp16 channel-last maps, half2 tap blending
            h, w = int(hms[0].shape[2]), int(hms[0].shape[3])
            if hms_f16 is None:
                hms_f16 = ops.heatmaps_to_f16(hms, st, C)
            ops.unproject([hms_f16[v] for v in range(len(hms))], (h * w * 16, 1, w * 16, 16), cams, centers, grid_size,
                          (X, Y, Z), self.img_size, (h, w), C, cubes, (X * Y * Z * 16, 1, 16), out_c_pad=16,
                          check_flag=check_flag, cubes_per_sample=cubes_per_sample, cube_sample=cube_sample,
                          heatmap_cfg_wh=self.heatmap_size, fast=True)
            return cubes, None
        grids = torch.empty(n_cubes, X * Y * Z, 3, device=dev, dtype=torch.float32) if want_grids else None
        # scaling uses cfg.NETWORK.HEATMAP_SIZE, sampling the tensor's own extent -- as the reference (:50,84-93)
        ops.unproject(hms, st, cams, centers, grid_size, (X, Y, Z), self.img_size, tuple(hms[0].shape[2:]), C,
                      cubes, (X * Y * Z * pitch, 1, pitch), out_c_pad=pitch, check_flag=check_flag,
                      cubes_per_sample=cubes_per_sample, cube_sample=cube_sample, grids=grids,
                      heatmap_cfg_wh=self.heatmap_size)
        return cubes, grids

    def get_voxel(self, heatmaps, meta, grid_size, grid_center, cube_size, flip_xcoords=None):
        device = heatmaps[0].device
        B, C = int(heatmaps[0].shape[0]), int(heatmaps[0].shape[1])
        X, Y, Z = [int(s) for s in cube_size]
        cams = ops.pack_cameras(meta, self.img_size, flip_xcoords).to(device, non_blocking=True)
        centers, check = self.centers_tensor(grid_center, B, device)
        hms, st = _common_strides(heatmaps)
        cubes = torch.empty(B, C, X, Y, Z, device=device, dtype=torch.float32)   # the reference's NCDHW layout
        grids = torch.empty(B, X * Y * Z, 3, device=device, dtype=torch.float32)
        ops.unproject(hms, st, cams, centers, grid_size, (X, Y, Z), self.img_size, tuple(hms[0].shape[2:]), C,
                      cubes, (C * X * Y * Z, X * Y * Z, 1), check_flag=check, grids=grids,
                      heatmap_cfg_wh=self.heatmap_size)
        return cubes, grids

    def forward(self, heatmaps, meta, grid_size, grid_center, cube_size, flip_xcoords=None):
        return self.get_voxel(heatmaps, meta, grid_size, grid_center, cube_size, flip_xcoords=flip_xcoords)
