"""Thin torch-tensor front end of the C ABI (``include/sp3d.h``).

PyTorch is used here for device memory, streams and nothing else: every
function takes CUDA tensors, fills the POD argument struct and launches the
hand-written kernel on torch's current stream.  There is no CPU path.
"""
from __future__ import annotations

import os
import weakref

import numpy as np
import torch

from . import _lib
from .utils.transforms import get_affine_transform

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


if hasattr(torch._C, "_cuda_getCurrentRawStream") and hasattr(torch._C, "_cuda_getDevice"):
    def _stream():
        """Raw handle of torch's current CUDA stream on the current device (the stream every launch goes to).
        ``torch.cuda.current_stream().cuda_stream`` costs ~17 us of Python per call -- as much as a small launch."""
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
else:       # (other torch builds: the public, slower route)
    def _stream():
        return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.Sp3dError(
                "selfpose3d_b200 kernels need CUDA tensors (got a %s tensor); this backend has no CPU path"
                % t.device.type)


def round_up(v, m):
    return (int(v) + m - 1) // m * m


# --------------------------------------------------------------------------------------------- cameras
def pack_cameras(meta, image_size, flip_xcoords=None):
    """``meta`` (list over views of collated dicts) -> ``[B, V, 32]`` float32 CPU tensor.

    One host pass replaces the per-(sample, view) ``get_affine_transform`` /
    ``unfold_camera_param`` calls of ``lib/models/project_layer.py:64-75``.  Camera
    parameters are cast to float32 exactly where the reference casts them
    (``lib/utils/cameras.py:14-23``, ``project_layer.py:69-72``).  If ``meta`` lives
    on a GPU (DataParallel scatter) this costs one device->host copy per tensor.
    """
    V = len(meta)
    B = int(meta[0]["center"].shape[0])
    table = np.zeros((B, V, _lib.CAM_FLOATS), dtype=np.float32)
    flips = None
    if flip_xcoords is not None:
        flips = np.asarray(torch.as_tensor(flip_xcoords).detach().cpu()).astype(bool)
    for c, m in enumerate(meta):
        center = np.asarray(m["center"].detach().cpu())
        scale = np.asarray(m["scale"].detach().cpu())
        rot = np.asarray(m["rotation"].detach().cpu())
        cam = {k: np.asarray(v.detach().cpu()) for k, v in m["camera"].items()
               if k in ("R", "T", "fx", "fy", "cx", "cy", "k", "p")}
        table[:, c, 0:9] = cam["R"].reshape(B, 9)
        table[:, c, 9:12] = cam["T"].reshape(B, 3)
        table[:, c, 12] = cam["fx"].reshape(B)
        table[:, c, 13] = cam["fy"].reshape(B)
        table[:, c, 14] = cam["cx"].reshape(B)
        table[:, c, 15] = cam["cy"].reshape(B)
        table[:, c, 16:19] = cam["k"].reshape(B, 3)
        table[:, c, 19:21] = cam["p"].reshape(B, 2)
        for i in range(B):
            # the affine depends on (center, scale, rotation, input size) only; views and samples of one rig share
            # it, so the float64 elimination runs once per distinct tuple (content-keyed: no stale entries)
            key = (center[i].tobytes(), scale[i].tobytes(), np.asarray(rot[i]).tobytes(), center.dtype.str,
                   scale.dtype.str, np.asarray(rot[i]).dtype.str, tuple(float(v) for v in image_size))
            trans = _AFFINE_CACHE.get(key)
            if trans is None:
                if len(_AFFINE_CACHE) > 4096:
                    _AFFINE_CACHE.clear()
                trans = get_affine_transform(center[i], scale[i], rot[i], image_size).reshape(6).astype(np.float32)
                _AFFINE_CACHE[key] = trans
            table[i, c, 21:27] = trans
            # width, height = center * 2, compared against float32 pixels (project_layer.py:68,78-80)
            table[i, c, 27] = center[i][0] * 2
            table[i, c, 28] = center[i][1] * 2
            table[i, c, 29] = 1.0 if (flips is not None and flips[i]) else 0.0
    return torch.from_numpy(table)


_LIN_CACHE = {}
_AFFINE_CACHE = {}   # (center, scale, rotation, input size) bytes -> float32 [6] affine

# dtype of the voxel cubes and V2V activations: float32 -> float32 SIMT convolutions (bit-faithful parity
# path), bfloat16 -> tcgen05 tensor-core convolutions with float32 accumulation.
_VOLUME_DTYPE = torch.float32


def set_volume_dtype(dtype):
    global _VOLUME_DTYPE
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("volume dtype must be torch.float32 or torch.bfloat16")
    _VOLUME_DTYPE = dtype


def volume_dtype():
    return _VOLUME_DTYPE


# How float32 activations are convolved.  "bf16x3" (the default: the product mode) / "bf16x6": the tcgen05 kernel on
# float32 data expanded into 2 / 3 bf16 terms (3 / 6 term pairs accumulated in float32 as extra K blocks) --
# float32-faithful results at tensor-core speed (SP3D_CONV_TC_BF16X3); "simt": the float32 FMA kernel (any shape; the
# cross-check path of the parity tests).  SP3D_F32_CONV in the environment overrides the default.
_F32_CONV = os.environ.get("SP3D_F32_CONV", "bf16x3")
if _F32_CONV not in ("simt", "bf16x3", "bf16x6"):
    raise ValueError('SP3D_F32_CONV must be "simt", "bf16x3" or "bf16x6"')


def set_float32_conv(mode):
    global _F32_CONV
    if mode not in ("simt", "bf16x3", "bf16x6"):
        raise ValueError('float32 convolution mode must be "simt", "bf16x3" or "bf16x6"')
    _F32_CONV = mode


def float32_conv():
    return _F32_CONV


# The 3-pair split mode evaluates x0 w0 + x0 w1 in one MMA of twice the columns where the kernel is instantiated for
# the layer (include/sp3d.h, split_terms 2); SP3D_WIDE_SPLIT=0 switches that off (A/B measurements).
_WIDE_SPLIT = os.environ.get("SP3D_WIDE_SPLIT", "1") != "0"
# SP3D_ZFOLD4=0 keeps the 7^3 stem on the 2-fold z-fold kernel (A/B measurements of the 4-fold one).
_ZFOLD4 = os.environ.get("SP3D_ZFOLD4", "1") != "0"
# SP3D_ZFOLD3=0 runs the 3^3 16 / 32 -> 32 layers un-folded (N = 32 kernels; A/B measurements under the power cap).
_ZFOLD3 = os.environ.get("SP3D_ZFOLD3", "1") != "0"


def use_split():
    """True in the float32-faithful tensor-core mode with 3 term pairs: evaluation-mode nets keep their activations as
    ``SplitAct`` term pairs from layer to layer."""
    return _VOLUME_DTYPE == torch.float32 and _F32_CONV == "bf16x3"


def linspace_axes(grid_size, cube_size, device):
    """The three ``torch.linspace(-s/2, s/2, n)`` vectors of ``compute_grid``
    (``lib/models/project_layer.py:28-30``), evaluated once on the host and cached on ``device``."""
    key = (tuple(float(s) for s in grid_size), tuple(int(n) for n in cube_size), str(device))
    if key not in _LIN_CACHE:
        _LIN_CACHE[key] = tuple(
            torch.linspace(-key[0][a] / 2, key[0][a] / 2, key[1][a]).to(device) for a in range(3))
    return _LIN_CACHE[key]


# --------------------------------------------------------------------------------------------- K1
def unproject(heatmaps, hm_strides, cams, centers, grid_size, cube_size, image_size, heatmap_hw, channels,
              out, out_strides, out_c_pad=0, check_flag=False, cubes_per_sample=1, cube_sample=None,
              grids=None, view_range=None, partial=False, heatmap_cfg_wh=None, fast=False, pair_out=False):
    """Launch the fused un-projection.  ``pair_out``: ``out`` is the two-plane bf16 tensor of a ``SplitAct``
    (``[2, n_cubes, ...]``, strides of one plane): float32 results written as term pairs.  ``fast``: the throughput form (fp16 channel-last maps from
    ``heatmaps_to_f16``, bf16 channel-last cubes with pitch 16; see ``csrc/unproject_fast.cu``).

    heatmaps: list[V] of CUDA float32 tensors sharing ``hm_strides = (b, c, h, w)`` element strides.
    cams ``[B,V,32]``, centers ``[n_cubes, >=3]`` float32 CUDA.  ``out`` receives the cubes through
    ``out_strides = (cube, channel, voxel)``.
    """
    a = unproject_args(heatmaps, hm_strides, cams, centers, grid_size, cube_size, image_size, heatmap_hw, channels,
                       out, out_strides, out_c_pad, check_flag, cubes_per_sample, cube_sample, grids, view_range,
                       partial, heatmap_cfg_wh, fast)
    if pair_out:
        a.out_dtype = _lib.BF16X2
    n_vox = a.X * a.Y * a.Z
    # SURVEY 8(d) algorithmic bytes: every float32 heat-map read once + every float32 cube written once (the bf16
    # volume mode physically moves half of the cube bytes; bench.py reports both)
    work = (a.view_end - a.view_begin) * a.B * a.C * a.h * a.w * 4 + a.n_cubes * a.C * n_vox * 4
    _lib.call("sp3d_unproject_fwd", a, _stream(), kind="unproject", work=work)


def unproject_args(heatmaps, hm_strides, cams, centers, grid_size, cube_size, image_size, heatmap_hw, channels,
                   out, out_strides, out_c_pad=0, check_flag=False, cubes_per_sample=1, cube_sample=None,
                   grids=None, view_range=None, partial=False, heatmap_cfg_wh=None, fast=False, a=None):
    """Fill ``sp3d_unproject_args`` (see ``unproject``); ``out`` may be None (backward: only the strides matter)."""
    _require_cuda(cams, centers, out, *heatmaps)
    a = _lib.UnprojectArgs() if a is None else a
    V = len(heatmaps)
    for v in range(V):
        a.heatmaps[v] = heatmaps[v].data_ptr()
    a.hm_stride_b, a.hm_stride_c, a.hm_stride_h, a.hm_stride_w = [int(s) for s in hm_strides]
    a.cams = cams.data_ptr()
    a.centers = centers.data_ptr()
    a.center_stride = int(centers.stride(0))
    a.check_flag = int(bool(check_flag))
    a.cubes_per_sample = int(cubes_per_sample)
    a.cube_sample = cube_sample.data_ptr() if cube_sample is not None else None
    lin = linspace_axes(grid_size, cube_size, cams.device)
    a.lin_x, a.lin_y, a.lin_z = lin[0].data_ptr(), lin[1].data_ptr(), lin[2].data_ptr()
    a.B = int(cams.shape[0])
    a.V = V
    a.C = int(channels)
    a.h, a.w = int(heatmap_hw[0]), int(heatmap_hw[1])
    a.n_cubes = int(centers.shape[0])
    a.X, a.Y, a.Z = [int(s) for s in cube_size]
    a.img_w, a.img_h = float(image_size[0]), float(image_size[1])
    if heatmap_cfg_wh is None:
        heatmap_cfg_wh = (a.w, a.h)
    a.hm_cfg_w, a.hm_cfg_h = float(heatmap_cfg_wh[0]), float(heatmap_cfg_wh[1])
    a.view_begin, a.view_end = (0, V) if view_range is None else (int(view_range[0]), int(view_range[1]))
    a.partial = int(bool(partial))
    a.cubes = out.data_ptr() if out is not None else None
    a.out_dtype = _DT[out.dtype] if out is not None else _lib.F32
    a.out_stride_cube, a.out_stride_c, a.out_stride_vox = [int(s) for s in out_strides]
    a.out_c_pad = int(out_c_pad)
    a.grids = grids.data_ptr() if grids is not None else None
    a.hm_dtype = _DT[heatmaps[0].dtype]
    a.math_mode = int(bool(fast))
    return a


def heatmaps_to_f16(heatmaps, hm_strides, channels):
    """list[V] of float32 ``[B,C,h,w]`` CUDA maps (common strides) -> one fp16 tensor ``[V,B,h,w,16]``
    (channels >= C zero): the heat-map layout of the throughput un-projection."""
    _require_cuda(*heatmaps)
    V = len(heatmaps)
    B, _, h, w = [int(s) for s in heatmaps[0].shape]
    out = torch.empty(V, B, h, w, 16, device=heatmaps[0].device, dtype=torch.float16)
    a = _lib.HeatmapsF16Args()
    for v in range(V):
        a.heatmaps[v] = heatmaps[v].data_ptr()
    a.stride_b, a.stride_c, a.stride_h, a.stride_w = [int(s) for s in hm_strides]
    a.V, a.B, a.C, a.h, a.w = V, B, int(channels), h, w
    a.out = out.data_ptr()
    _lib.call("sp3d_heatmaps_to_f16", a, _stream(), kind="layout", work=V * B * h * w * (int(channels) * 4 + 32))
    return out


def unproject_finalize(buf, n_cubes, channels, n_vox, strides):
    a = _lib.UnprojectFinalizeArgs()
    a.buf = buf.data_ptr()
    a.n_cubes, a.C, a.N = int(n_cubes), int(channels), int(n_vox)
    a.stride_cube, a.stride_c, a.stride_vox = [int(s) for s in strides]
    _lib.call("sp3d_unproject_finalize", a, _stream(), kind="unproject_finalize",
              work=2 * 4 * int(n_cubes) * (int(channels) + 1) * int(n_vox))


# --------------------------------------------------------------------------------------------- K3
def nms_topk(root_cubes, max_people, threshold, space_size, space_center, loc_f64=False, return_index=False):
    """``root_cubes [B,X,Y,Z]`` float32 contiguous -> ``grid_centers [B,K,5]``."""
    _require_cuda(root_cubes)
    if not root_cubes.is_contiguous() or root_cubes.dtype != torch.float32:
        raise _lib.Sp3dError("nms_topk expects a contiguous float32 [B,X,Y,Z] tensor")
    B, X, Y, Z = root_cubes.shape
    K = int(max_people)
    gc = torch.empty(B, K, 5, device=root_cubes.device, dtype=torch.float32)
    idx = torch.empty(B, K, device=root_cubes.device, dtype=torch.int32) if return_index else None
    a = _lib.NmsTopkArgs()
    a.root_cubes = root_cubes.data_ptr()
    a.B, a.X, a.Y, a.Z, a.K = B, X, Y, Z, K
    a.threshold = float(threshold)
    for d in range(3):
        a.space_size[d] = float(space_size[d])
        a.space_center[d] = float(space_center[d])
    a.loc_f64 = int(bool(loc_f64))
    a.grid_centers = gc.data_ptr()
    a.topk_index = idx.data_ptr() if idx is not None else None
    nbytes = int(_lib.load().sp3d_nms_topk3d_workspace(a))
    ws = torch.empty(max(nbytes // 4, 1), device=root_cubes.device, dtype=torch.float32)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    _lib.call("sp3d_nms_topk3d", a, _stream(), launches=2, kind="nms_topk", work=4 * B * X * Y * Z)
    return (gc, idx) if return_index else gc


# --------------------------------------------------------------------------------------------- K4
def softargmax_axes(x, strides, n_cubes, channels, cube_size, centers, lin, beta, check_flag=False):
    """Soft-argmax with explicit per-axis coordinate vectors ``lin = (x[X], y[Y], z[Z])`` (centre-free)."""
    return softargmax(x, strides, n_cubes, channels, cube_size, centers, None, beta, check_flag, lin=lin)


def softargmax(x, strides, n_cubes, channels, cube_size, centers, grid_size, beta, check_flag=False, lin=None):
    """``x`` addressed as ``[n_cubes, C, N]`` through ``strides = (cube, channel, voxel)`` -> ``[n_cubes, C, 3]``."""
    _require_cuda(x, centers)
    out = torch.empty(n_cubes, channels, 3, device=x.device, dtype=torch.float32)
    a = _lib.SoftargmaxArgs()
    a.x = x.data_ptr()
    a.x_dtype = _DT[x.dtype]
    a.stride_cube, a.stride_c, a.stride_vox = [int(s) for s in strides]
    a.n_cubes, a.C = int(n_cubes), int(channels)
    a.X, a.Y, a.Z = [int(s) for s in cube_size]
    a.centers = centers.data_ptr()
    a.center_stride = int(centers.stride(0))
    a.check_flag = int(bool(check_flag))
    if lin is None:
        lin = linspace_axes(grid_size, cube_size, x.device)
    a.lin_x, a.lin_y, a.lin_z = lin[0].data_ptr(), lin[1].data_ptr(), lin[2].data_ptr()
    a.beta = float(beta)
    a.out = out.data_ptr()
    nbytes = int(_lib.load().sp3d_softargmax3d_workspace(a))
    ws = torch.empty(max(nbytes // 8, 1), device=x.device, dtype=torch.float64)
    a.workspace = ws.data_ptr()
    a.workspace_bytes = nbytes
    _lib.call("sp3d_softargmax3d_fwd", a, _stream(), launches=2, kind="softargmax",
              work=a.n_cubes * a.C * a.X * a.Y * a.Z * x.element_size())
    return out


# --------------------------------------------------------------------------------------------- layout
def to_channel_last(x, c_pitch=None, dtype=None):
    """Channel-first ``[N, C, *spatial]`` (contiguous) -> channel-last ``[N, *spatial, c_pitch]``."""
    _require_cuda(x)
    x = x.contiguous()
    N, Cc = int(x.shape[0]), int(x.shape[1])
    spatial = tuple(int(s) for s in x.shape[2:])
    S = int(np.prod(spatial))
    c_pitch = round_up(Cc, 4) if c_pitch is None else int(c_pitch)
    dst = torch.empty((N,) + spatial + (c_pitch,), device=x.device, dtype=dtype or x.dtype)
    a = _lib.LayoutArgs()
    a.src, a.dst = x.data_ptr(), dst.data_ptr()
    a.N, a.C, a.S, a.c_pitch = N, Cc, S, c_pitch
    a.to_channel_last = 1
    a.src_dtype, a.dst_dtype = _DT[x.dtype], _DT[dst.dtype]
    _lib.call("sp3d_layout_convert", a, _stream(), kind="layout", work=N * Cc * S * (x.element_size() + dst.element_size()))
    return dst


def to_channel_first(x, channels, dtype=None):
    """Channel-last ``[N, *spatial, c_pitch]`` -> contiguous channel-first ``[N, channels, *spatial]``."""
    _require_cuda(x)
    if not x.is_contiguous():
        raise _lib.Sp3dError("to_channel_first expects a contiguous channel-last tensor")
    N = int(x.shape[0])
    spatial = tuple(int(s) for s in x.shape[1:-1])
    S = int(np.prod(spatial))
    dst = torch.empty((N, int(channels)) + spatial, device=x.device, dtype=dtype or x.dtype)
    a = _lib.LayoutArgs()
    a.src, a.dst = x.data_ptr(), dst.data_ptr()
    a.N, a.C, a.S, a.c_pitch = N, int(channels), S, int(x.shape[-1])
    a.to_channel_last = 0
    a.src_dtype, a.dst_dtype = _DT[x.dtype], _DT[dst.dtype]
    _lib.call("sp3d_layout_convert", a, _stream(), kind="layout",
              work=N * int(channels) * S * (x.element_size() + dst.element_size()))
    return dst


def space_to_depth(x, channels, strides, n, h, w, dst_pitch, pair=False):
    """``x`` addressed as ``[n, channels, h, w]`` through element ``strides = (n, c, y, x)`` (float32 or bf16) ->
    channel-last bf16 ``[n, 1, h/2, w/2, dst_pitch]`` with channel ``(py*2+px)*channels + c``.  ``pair`` (float32
    source): the result is the two-plane tensor ``[2, n, 1, h/2, w/2, dst_pitch]`` of a ``SplitAct``."""
    _require_cuda(x)
    shape = (n, 1, h // 2, w // 2, dst_pitch)
    out = torch.empty(((2,) + shape) if pair else shape, device=x.device, dtype=torch.bfloat16)
    a = _lib.S2DArgs()
    a.src, a.dst = x.data_ptr(), out.data_ptr()
    a.src_dtype = _DT[x.dtype]
    a.dst_dtype = _lib.BF16X2 if pair else _lib.BF16
    a.stride_n, a.stride_c, a.stride_y, a.stride_x = [int(s) for s in strides]
    a.N, a.C, a.H, a.W, a.dst_pitch = int(n), int(channels), int(h), int(w), int(dst_pitch)
    _lib.call("sp3d_space_to_depth", a, _stream(), kind="layout",
              work=n * h * w * channels * x.element_size() + out.numel() * 2)
    return out


def stack_x_shifts(x, taps, pad):
    """Channel-last bf16 ``[N,X,Y,Z,pitch]`` (channel 0 used) -> ``[N,X,Y,Z,16]`` with channel ``j`` =
    ``x[:, x+j-pad, y, z, 0]`` (zero outside), ``j < taps``."""
    _require_cuda(x)
    N, X, Y, Z, pitch = [int(v) for v in x.shape]
    out = torch.empty(N, X, Y, Z, 16, device=x.device, dtype=torch.bfloat16)
    a = _lib.StackArgs()
    a.src, a.dst = x.data_ptr(), out.data_ptr()
    a.N, a.X, a.Y, a.Z, a.src_pitch, a.taps, a.pad = N, X, Y, Z, pitch, int(taps), int(pad)
    _lib.call("sp3d_stack_x_shifts", a, _stream(), kind="layout", work=N * X * Y * Z * (2 * taps + 32))
    return out


def split_bf16(x, channels, c_block, blocks):
    """float32 channel-last ``[..., pitch]`` -> bf16 ``[blocks, ..., c_block]``: plane ``s`` holds term ``s`` of
    the expansion ``x = x0 + x1 (+ x2)`` into bf16 values (``sp3d_split_bf16``)."""
    _require_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise _lib.Sp3dError("split_bf16 expects a contiguous float32 channel-last tensor")
    out = torch.empty((blocks,) + tuple(x.shape[:-1]) + (c_block,), device=x.device, dtype=torch.bfloat16)
    a = _lib.SplitArgs()
    a.src, a.dst = x.data_ptr(), out.data_ptr()
    a.P = x.numel() // int(x.shape[-1])
    a.C, a.src_pitch, a.c_block, a.S = int(channels), int(x.shape[-1]), int(c_block), int(blocks)
    _lib.call("sp3d_split_bf16", a, _stream(), kind="layout", work=a.P * (int(channels) * 4 + blocks * c_block * 2))
    return out


def bf16_terms(w, n):
    """float32 tensor -> ``n`` float32 tensors of bf16-representable values with ``w ~= sum`` (each term is the
    bf16 rounding of what the previous ones left; the host-side twin of ``sp3d_split_bf16``)."""
    terms, rest = [], w.float()
    for _ in range(n):
        t = rest.to(torch.bfloat16).float()
        terms.append(t)
        rest = rest - t
    return terms


class SplitAct:
    """float32 channel-last activations held as TWO bf16 term planes (``SP3D_BF16X2`` of include/sp3d.h):
    ``planes [2, N, D, H, W, pitch]`` with ``x ~= planes[0] + planes[1]``.  This is what the float32-faithful
    tensor-core convolutions read; their epilogue, the max-pool, the un-projection and the space-to-depth kernel write
    it directly, so no separate split pass runs between the layers of a net."""
    __slots__ = ("planes",)

    def __init__(self, planes):
        if planes.dtype != torch.bfloat16 or planes.dim() != 6 or planes.shape[0] != 2 or not planes.is_contiguous():
            raise _lib.Sp3dError("SplitAct expects a contiguous bf16 [2, N, D, H, W, pitch] tensor")
        self.planes = planes

    @property
    def shape(self):
        return self.planes.shape[1:]

    @property
    def device(self):
        return self.planes.device


def split_pitch(channels):
    """Channel pitch of a ``SplitAct`` with ``channels`` channels = the K extent the consuming convolution is packed for."""
    return round_up(channels, 16) if channels < 64 else round_up(channels, 64)


def split_act(x, channels, pitch=None):
    """float32 channel-last ``[N,D,H,W,p]`` -> ``SplitAct`` (one ``sp3d_split_bf16`` pass)."""
    return SplitAct(split_bf16(x, channels, split_pitch(channels) if pitch is None else int(pitch), 2))


def merge_act(x, channels, pitch=None):
    """``SplitAct`` -> float32 channel-last ``[N,D,H,W,pitch]`` (``sp3d_merge_bf16``; padding channels zero)."""
    planes = x.planes
    pitch = round_up(channels, 4) if pitch is None else int(pitch)
    out = torch.empty(tuple(planes.shape[1:-1]) + (pitch,), device=planes.device, dtype=torch.float32)
    a = _lib.SplitArgs()
    a.src, a.dst = out.data_ptr(), planes.data_ptr()
    a.P = planes[0].numel() // int(planes.shape[-1])
    a.C, a.src_pitch, a.c_block, a.S = int(channels), pitch, int(planes.shape[-1]), 2
    _lib.call("sp3d_merge_bf16", a, _stream(), kind="layout", work=a.P * int(channels) * 8)
    return out


# (activation term, weight term) per K block of the split-operand convolution (include/sp3d.h, split_terms): the
# small correction products are accumulated FIRST -- the tensor core's float32 accumulator loses ~2^-24 of its
# current magnitude per MMA, so only the final x0 w0 chain should run on a full-size accumulator
SPLIT_PAIRS = {3: ((1, 0), (0, 1), (0, 0)), 6: ((2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0))}


def _tc_finish(full, terms):
    """float32 ``[n_tiles, n_chunks, taps, N, chunk]`` -> the bf16 tensor the kernel streams: as is (``terms`` 0), or
    ``[n_tiles, K blocks, n_chunks, taps, N, chunk]`` with K block ``b`` = weight term of ``SPLIT_PAIRS[terms][b]``;
    ``terms`` 2 (the 3 pairs in two K blocks, include/sp3d.h): ``[n_tiles, rows, chunk]`` -- per tile the x0 block
    ``[chunk][tap][w0 rows | w1 rows]`` followed by the x1 block ``[chunk][tap][w0 rows]``."""
    if not terms:
        return full.to(torch.bfloat16).contiguous()
    if terms == 2:
        w0, w1 = bf16_terms(full, 2)
        nt, chunk = int(full.shape[0]), int(full.shape[-1])
        wide = torch.cat([w0, w1], 3).reshape(nt, -1, chunk)
        return torch.cat([wide, w0.reshape(nt, -1, chunk)], 1).to(torch.bfloat16).contiguous()
    wts = bf16_terms(full, max(wt for _, wt in SPLIT_PAIRS[terms]) + 1)
    return torch.stack([wts[wt] for _, wt in SPLIT_PAIRS[terms]], 1).to(torch.bfloat16).contiguous()


# (kernel extent along x, along y / z, shared-memory row bytes, N = MMA columns, z-fold) of every conv_tc_kernel
# instantiation in csrc/conv_tc.cu (the SP3D_TC_CASE list; tests/test_host_cpu.py keeps the two in step)
# instantiations of the 2-K-block form of the 3-pair split mode (split_terms 2: accumulators of 2 N columns)
TC_WIDE_CASES = frozenset({(7, 7, 64, 32, 2), (1, 7, 64, 32, 2), (3, 3, 64, 32, 1), (3, 3, 128, 64, 1), (3, 3, 128, 64, 2),
                           (3, 3, 64, 64, 2), (7, 7, 128, 64, 4)})
TC_CASES = frozenset({
    (7, 7, 32, 16, 1), (7, 7, 64, 32, 2), (3, 3, 64, 64, 2), (1, 7, 64, 32, 2), (3, 3, 128, 64, 2), (3, 3, 32, 32, 1),
    (3, 3, 64, 32, 1), (3, 3, 64, 64, 1), (3, 3, 128, 64, 1), (3, 3, 128, 128, 1), (3, 3, 64, 16, 1), (3, 3, 128, 32, 1),
    (1, 1, 32, 32, 1), (1, 1, 64, 64, 1),
    (1, 1, 64, 16, 1), (1, 1, 128, 16, 1), (1, 1, 128, 32, 1), (1, 1, 128, 64, 1), (1, 1, 128, 128, 1),
    (1, 3, 128, 64, 1), (1, 3, 128, 128, 1), (1, 2, 128, 128, 1), (1, 4, 32, 64, 1)})


def tc_case(ksize, cin, n, zfold=1):
    """The kernel instantiation ``conv_tc`` (csrc/conv_tc.cu) dispatches a launch to: ``cin`` = channels per K block
    as passed to the launch, ``n`` = ``cout_pitch_w``."""
    return (int(ksize[0]), int(ksize[1]), 128 if cin >= 64 else int(cin) * 2 * max(int(zfold), 1), int(n), max(int(zfold), 1))


class S2DConv:
    """A stride-2 2-D convolution (3x3/p1 or 7x7/p3, + folded BatchNorm + ReLU) evaluated on the tcgen05 path as a
    stride-1 convolution over the 2x2 space-to-depth tensor: tap ``d`` with offset ``t = d - pad`` lands on
    sub-pixel ``t mod 2`` of tap ``floor(t / 2)``, so the kernel shrinks to 2x2 / 4x4 over ``4 * cin`` channels."""

    def __init__(self, weight, bn, padding, relu):
        w = weight.detach().float()
        self.cout, self.cin, k = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
        pad = int(padding)
        self.dmin = (0 - pad) // 2
        self.kp = (k - 1 - pad) // 2 - self.dmin + 1
        c4 = 4 * self.cin
        self.cin_tc = round_up(c4, 16) if c4 < 64 else round_up(c4, 64)
        chunk = min(self.cin_tc, 64)
        self.n = next((v for v in (64, 128) if v >= self.cout), 128)
        n_tiles = -(-self.cout // self.n)
        w2 = torch.zeros(self.kp, self.kp, n_tiles * self.n, self.cin_tc, device=w.device, dtype=torch.float32)
        for dy in range(k):
            ty = dy - pad
            for dx in range(k):
                tx = dx - pad
                q = (ty - 2 * (ty // 2)) * 2 + (tx - 2 * (tx // 2))
                w2[ty // 2 - self.dmin, tx // 2 - self.dmin, :self.cout, q * self.cin:(q + 1) * self.cin] = w[:, :, dy, dx]
        taps = self.kp * self.kp
        full = w2.reshape(taps, n_tiles, self.n, self.cin_tc // chunk, chunk).permute(1, 3, 0, 2, 4)
        self.weight = full.to(torch.bfloat16).contiguous()
        self._full = full
        self._weight3 = None
        inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
        self.scale = (bn.weight.detach().float() * inv).contiguous()
        self.shift = (bn.bias.detach().float() - bn.running_mean.detach().float() * self.scale).contiguous()
        self.relu = int(relu)

    @staticmethod
    def supported(conv, h, w):
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        return (s == 2 and (k, p) in ((3, 1), (7, 3)) and h % 2 == 0 and w % 2 == 0 and conv.bias is None
                and (conv.in_channels * 4 < 64 or conv.in_channels % 16 == 0) and conv.out_channels % 64 == 0)

    def __call__(self, x, strides, n, h, w):
        """``x``: source addressed as ``[n, cin, h, w]`` through ``strides``; returns channel-last bf16
        ``[n, 1, h/2, w/2, cout]``."""
        xs = space_to_depth(x, self.cin, strides, n, h, w, self.cin_tc)
        oh, ow = h // 2, w // 2
        out = torch.empty(n, 1, oh, ow, round_up(self.cout, 16), device=x.device, dtype=torch.bfloat16)
        conv_launch(xs.view(1, n, oh, ow, self.cin_tc), self.weight, self.scale, self.shift, None,
                    out.view(1, n, oh, ow, out.shape[-1]), self.cin_tc, self.cout, (n, oh, ow), [1, self.kp, self.kp],
                    [1, 1, 1], [0, self.dmin, self.dmin], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu, _lib.CONV_TC_BF16,
                    cin_real=self.cin * 4, cout_pitch_w=self.n)
        return out

    def call_split(self, x, strides=None, n=None, h=None, w=None):
        """Float32-faithful form (``SP3D_CONV_TC_BF16X3``, 3 term pairs): ``x`` is a ``SplitAct`` ``[2,n,1,h,w,cin]``, or
        a float32 tensor addressed as ``[n, cin, h, w]`` through ``strides`` (the NCHW image).  Returns a ``SplitAct``."""
        if isinstance(x, SplitAct):
            _, n, _, h, w, c = [int(v) for v in x.planes.shape]
            xs = space_to_depth(x.planes, self.cin, (h * w * c, 1, w * c, c), 2 * n, h, w, self.cin_tc)   # plane by plane
        else:
            xs = space_to_depth(x, self.cin, strides, n, h, w, self.cin_tc, pair=True)
        if self._weight3 is None:
            self._weight3 = _tc_finish(self._full, 3)
        oh, ow = h // 2, w // 2
        pitch = split_pitch(self.cout)
        out = torch.empty(2, n, 1, oh, ow, pitch, device=xs.device, dtype=torch.bfloat16)
        conv_launch(xs.view(2, 1, n, oh, ow, self.cin_tc)[0], self._weight3, self.scale, self.shift, None,
                    out.view(2, 1, n, oh, ow, pitch)[0], self.cin_tc, self.cout, (n, oh, ow), [1, self.kp, self.kp],
                    [1, 1, 1], [0, self.dmin, self.dmin], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu, _lib.CONV_TC_BF16X3,
                    cin_real=self.cin * 4, cout_pitch_w=self.n, split_terms=3, pair_out=True)
        return SplitAct(out)


# --------------------------------------------------------------------------------------------- conv family
def _set3(field, vals):
    for i in range(3):
        field[i] = int(vals[i])


def conv_launch(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0, tap_step,
                ostride, ooffset, relu, algo=_lib.CONV_SIMT_F32, cin_real=None, cout_pitch_w=None, fused_phases=False,
                zfold=0, head=None, split_terms=0, pair_out=False):
    """One implicit-GEMM convolution launch.  ``x`` / ``out`` are channel-last 5-D ``[N,D,H,W,pitch]``.
    ``pair_out``: ``out`` (and ``residual``) are plane 0 of two-plane bf16 tensors (``SplitAct.planes[0]`` views).
    ``head``: a ``SoftargmaxHead`` -- the output is consumed on chip by the fused soft-argmax (``out`` is then a
    shape-only placeholder: a ``torch.Size``-like tuple ``(N, D, H, W, pitch)``)."""
    if head is not None:
        return _conv_launch_head(x, weight, scale, shift, out, cin, cout, out_grid, ksize, relu, algo, cin_real,
                                 cout_pitch_w, head)
    # the argument struct of a (layer, shapes, launch geometry) is built once and re-used: only the three activation
    # pointers change between calls (filling ~40 ctypes fields costs more host time than a small kernel runs)
    key = (weight.data_ptr(), tuple(x.shape), x.dtype, tuple(out.shape), out.dtype, residual is not None,
           tuple(out_grid), tuple(tap_off0), tuple(ostride), tuple(ooffset), int(algo), bool(fused_phases), int(zfold),
           int(split_terms), bool(pair_out))
    hit = _CONV_ARGS_CACHE.get(key)
    if hit is None:
        a = _lib.ConvArgs()
        a.weight = weight.data_ptr()
        a.scale = scale.data_ptr() if scale is not None else None
        a.shift = shift.data_ptr() if shift is not None else None
        a.N, a.D, a.H, a.W = [int(s) for s in x.shape[:4]]
        a.cin = int(cin)
        a.cin_pitch = int(x.shape[4])
        a.OD, a.OH, a.OW = [int(s) for s in out_grid]
        a.TD, a.TH, a.TW = [int(s) for s in out.shape[1:4]]
        a.cout = int(cout)
        a.cout_pitch = int(out.shape[4])
        a.cout_pitch_w = int(weight.shape[-1]) if cout_pitch_w is None else int(cout_pitch_w)
        _set3(a.ksize, ksize)
        _set3(a.stride, stride)
        _set3(a.tap_off0, tap_off0)
        _set3(a.tap_step, tap_step)
        _set3(a.ostride, ostride)
        _set3(a.ooffset, ooffset)
        a.relu = int(relu)
        a.algo = int(algo)
        a.in_dtype, a.out_dtype = _DT[x.dtype], (_lib.BF16X2 if pair_out else _DT[out.dtype])
        a.fused_phases = int(bool(fused_phases))
        a.zfold = int(zfold)
        a.split_terms = int(split_terms)
        flops = 2.0 * a.N * a.OD * a.OH * a.OW * a.cout * (cin_real or a.cin) * (a.ksize[0] * a.ksize[1] * a.ksize[2])
        if fused_phases:
            flops *= 8
        detail = "conv%s algo%d k%d %d->%d @%dx%dx%dx%d" % ("T8" if fused_phases else "", a.algo, a.ksize[2],
                                                           cin_real or a.cin, a.cout, a.N, a.OD, a.OH, a.OW)
        if len(_CONV_ARGS_CACHE) > 256:
            # entries whose packed weight is gone (the training path re-packs every layer after each optimizer step)
            for k in [k for k, v in _CONV_ARGS_CACHE.items() if v[3]() is None]:
                del _CONV_ARGS_CACHE[k]
            if len(_CONV_ARGS_CACHE) > 2048:
                _CONV_ARGS_CACHE.clear()
        # the entry holds a WEAK reference to the packed weight: it dies with the PackedConv that owns the tensor (no
        # stale packs pinned in device memory), and a recycled address is detected by the dead / different referent
        hit = _CONV_ARGS_CACHE[key] = (a, flops, detail, weakref.ref(weight), (int(relu), int(cout), int(cin)),
                                       (scale.data_ptr() if scale is not None else 0,
                                        shift.data_ptr() if shift is not None else 0))
    a, flops, detail, wref, sig, ss = hit
    if wref() is not weight or ss != (scale.data_ptr() if scale is not None else 0,
                                      shift.data_ptr() if shift is not None else 0):
        del _CONV_ARGS_CACHE[key]        # the address was recycled by another tensor: rebuild
        return conv_launch(x, weight, scale, shift, residual, out, cin, cout, out_grid, ksize, stride, tap_off0, tap_step,
                           ostride, ooffset, relu, algo, cin_real, cout_pitch_w, fused_phases, zfold, head, split_terms,
                           pair_out)
    if sig != (int(relu), int(cout), int(cin)):
        raise _lib.Sp3dError("conv_launch cache collision")
    a.in_ = x.data_ptr()
    a.residual = residual.data_ptr() if residual is not None else None
    a.out = out.data_ptr()
    _lib.call("sp3d_conv_fwd", a, _stream(), kind="conv", work=flops, detail=detail)


_CONV_ARGS_CACHE = {}


class SoftargmaxHead:
    """Arguments of the soft-argmax fused behind a 1x1x1 tensor-core convolution (``sp3d_conv_args.head_softargmax``):
    ``centers [n, >=3]`` float32 CUDA, cube geometry, beta.  ``out`` receives ``[n, channels, 3]``."""

    def __init__(self, n_cubes, channels, cube_size, centers, grid_size, beta):
        _require_cuda(centers)
        dev = centers.device
        self.out = torch.empty(n_cubes, channels, 3, device=dev, dtype=torch.float32)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.ws = torch.empty(max(n_cubes * 2 * sms * channels * 5, 1), device=dev, dtype=torch.float64)
        self.centers = centers
        self.lin = linspace_axes(grid_size, cube_size, dev)
        a = _lib.SoftargmaxArgs()
        a.x = None
        a.n_cubes, a.C = int(n_cubes), int(channels)
        a.X, a.Y, a.Z = [int(v) for v in cube_size]
        a.centers = centers.data_ptr()
        a.center_stride = int(centers.stride(0))
        a.check_flag = 0
        a.lin_x, a.lin_y, a.lin_z = self.lin[0].data_ptr(), self.lin[1].data_ptr(), self.lin[2].data_ptr()
        a.beta = float(beta)
        a.out = self.out.data_ptr()
        a.workspace = self.ws.data_ptr()
        a.workspace_bytes = self.ws.numel() * 8
        self.args = a


def _conv_launch_head(x, weight, scale, shift, out_shape, cin, cout, out_grid, ksize, relu, algo, cin_real,
                      cout_pitch_w, head):
    import ctypes
    a = _lib.ConvArgs()
    a.in_ = x.data_ptr()
    a.weight = weight.data_ptr()
    a.scale = scale.data_ptr() if scale is not None else None
    a.shift = shift.data_ptr() if shift is not None else None
    a.residual = None
    a.out = None
    a.N, a.D, a.H, a.W = [int(s) for s in x.shape[:4]]
    a.cin = int(cin)
    a.cin_pitch = int(x.shape[4])
    a.OD, a.OH, a.OW = [int(s) for s in out_grid]
    a.TD, a.TH, a.TW = [int(s) for s in out_shape[1:4]]
    a.cout = int(cout)
    a.cout_pitch = int(out_shape[4])
    a.cout_pitch_w = int(cout_pitch_w)
    _set3(a.ksize, ksize)
    _set3(a.stride, [1, 1, 1])
    _set3(a.tap_off0, [0, 0, 0])
    _set3(a.tap_step, [1, 1, 1])
    _set3(a.ostride, [1, 1, 1])
    _set3(a.ooffset, [0, 0, 0])
    a.relu = int(relu)
    a.algo = int(algo)
    a.in_dtype, a.out_dtype = _DT[x.dtype], _lib.F32
    a.head_softargmax = ctypes.pointer(head.args)
    flops = 2.0 * a.N * a.OD * a.OH * a.OW * a.cout * (cin_real or a.cin)
    detail = "conv+softargmax algo%d k1 %d->%d @%dx%dx%dx%d" % (a.algo, cin_real or a.cin, a.cout, a.N, a.OD, a.OH, a.OW)
    _lib.call("sp3d_conv_fwd", a, _stream(), launches=2, kind="conv", work=flops, detail=detail)
    return head.out


def maxpool(x, channels, k, s, p):
    """Channel-last ``[N,D,H,W,pitch]`` max pooling; returns the pooled channel-last tensor (a ``SplitAct`` for a
    ``SplitAct``: the maximum of the float32 values the term pairs stand for, re-split)."""
    pair = isinstance(x, SplitAct)
    if pair:
        x = x.planes
        N, D, H, W, pitch = [int(v) for v in x.shape[1:]]
    else:
        N, D, H, W, pitch = [int(v) for v in x.shape]
    _require_cuda(x)
    dims = (D, H, W)
    o = [(dims[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3)]
    out = torch.empty(((2,) if pair else ()) + (N, o[0], o[1], o[2], pitch), device=x.device, dtype=x.dtype)
    a = _lib.MaxpoolArgs()
    a.in_, a.out = x.data_ptr(), out.data_ptr()
    a.N, a.D, a.H, a.W, a.C, a.c_pitch = N, D, H, W, int(channels), pitch
    a.OD, a.OH, a.OW = o
    _set3(a.k, k)
    _set3(a.s, s)
    _set3(a.p, p)
    a.dtype = _lib.BF16X2 if pair else _DT[x.dtype]
    _lib.call("sp3d_maxpool_fwd", a, _stream(), kind="maxpool",
              work=(N * D * H * W + N * o[0] * o[1] * o[2]) * int(channels) * x.element_size() * (2 if pair else 1))
    return SplitAct(out) if pair else out


class PackedConv:
    """A convolution (regular or transposed) with evaluation-mode BatchNorm folded in, packed for
    ``sp3d_conv_fwd``: weights ``[taps, cin_p, cout_p]``, per-channel ``scale`` / ``shift``.

    Built from the reference-shaped parameters (``nn.Conv{2,3}d`` ``[Cout,Cin,k..]`` or
    ``nn.ConvTranspose{2,3}d`` ``[Cin,Cout,k..]``, optional ``nn.BatchNorm``), so the module tree
    and its state dict stay exactly the reference's.
    """

    def __init__(self, weight, bias=None, bn=None, stride=1, padding=0, transposed=False, relu=0):
        w = weight.detach()
        self.nd = w.dim() - 2
        self.transposed = bool(transposed)
        self.relu = int(relu)
        k = [1] * (3 - self.nd) + [int(s) for s in w.shape[2:]]
        self.k = k
        self.stride = [1] * (3 - self.nd) + [int(stride)] * self.nd
        self.padding = [0] * (3 - self.nd) + [int(padding)] * self.nd
        w5 = w.reshape(w.shape[0], w.shape[1], *k).float()
        if transposed:
            self.cin, self.cout = int(w.shape[0]), int(w.shape[1])
            w5 = w5.permute(1, 0, 2, 3, 4)  # -> [Cout, Cin, kd, kh, kw]
        else:
            self.cout, self.cin = int(w.shape[0]), int(w.shape[1])
        self.cin_p = round_up(self.cin, 4)
        self.cout_pw = round_up(self.cout, 4)
        dev = w.device

        scale = None
        shift = bias.detach().float().clone() if bias is not None else None
        if bn is not None:
            inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
            scale = bn.weight.detach().float() * inv
            base = shift if shift is not None else torch.zeros(self.cout, device=dev)
            shift = (base - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
        self.scale = scale.contiguous() if scale is not None else None
        self.shift = shift.contiguous() if shift is not None else None

        def pack(sub):  # sub: [Cout, Cin, a, b, c] -> [a*b*c, cin_p, cout_pw]
            t = sub.permute(2, 3, 4, 1, 0).reshape(-1, self.cin, self.cout)
            out = torch.zeros(t.shape[0], self.cin_p, self.cout_pw, device=dev, dtype=torch.float32)
            out[:, :self.cin, :self.cout] = t
            return out.contiguous()

        self._subs = []      # float32 [Cout, Cin, a, b, c] sub-kernels (one, or one per transposed-conv phase)
        self._tc = None      # lazily built bf16 packing for the tcgen05 path
        if not transposed:
            self.weights = [pack(w5)]
            self.phases = [None]
            self._subs = [w5]
        else:
            # one stride-1 sub-convolution per output phase (see include/sp3d.h, sp3d_conv_args)
            self.weights, self.phases = [], []
            s, p = self.stride, self.padding
            for pd in range(s[0]):
                for ph in range(s[1]):
                    for pw in range(s[2]):
                        phase = (pd, ph, pw)
                        t0 = [(phase[i] + p[i]) % s[i] for i in range(3)]
                        sub = w5[:, :, t0[0]::s[0], t0[1]::s[1], t0[2]::s[2]]
                        off0 = [(phase[i] + p[i] - t0[i]) // s[i] for i in range(3)]
                        # (a phase without taps, e.g. kernel 1 / stride 2, only ever produces the shift)
                        self.weights.append(pack(sub) if sub.numel() else None)
                        self.phases.append((phase, off0, [int(v) for v in sub.shape[2:]]))
                        self._subs.append(sub)

    def out_shape(self, dims):
        if not self.transposed:
            return [(dims[i] + 2 * self.padding[i] - self.k[i]) // self.stride[i] + 1 for i in range(3)]
        return [(dims[i] - 1) * self.stride[i] - 2 * self.padding[i] + self.k[i] for i in range(3)]

    # ------------------------------------------------------------------ tcgen05 (bf16) path
    def tc_supported(self):
        """What ``csrc/conv_tc.cu`` takes: 3-D "same" convolutions with k in {1,3,7} and k2/s2 transposed
        convolutions (V2VNet); 2-D 1x1 (stride 1 or 2), 3x3 stride-1 "same" convolutions and k4/s2/p1
        transposed convolutions on >= 64 input channels (PoseResNet)."""
        k, s, p = self.k, self.stride, self.padding
        if self.nd == 3:
            if self.transposed:
                return k == [2, 2, 2] and s == [2, 2, 2] and p == [0, 0, 0]
            return k[0] in (1, 3, 7) and k == [k[0]] * 3 and s == [1, 1, 1] and p == [k[0] // 2] * 3
        if self.nd == 2 and self.cin % 64 == 0:
            if self.transposed:
                return k == [1, 4, 4] and s == [1, 2, 2] and p == [0, 1, 1]
            if k == [1, 1, 1]:
                return s in ([1, 1, 1], [1, 2, 2]) and p == [0, 0, 0]
            return k == [1, 3, 3] and s == [1, 1, 1] and p == [0, 1, 1]
        return False

    # Weight packings of the tcgen05 path.  Each returns the bf16 tensor the kernel streams,
    # [n_tiles, (K blocks,) n_chunks, taps, N, chunk] (rows = output channel, K-major); ``terms`` = 0 for bf16
    # operands, 3 / 6 for float32 operands split into bf16 terms (one K block per term pair, ``_tc_finish``).
    def _tc_cached(self, kind, terms, build):
        cache = self.__dict__.setdefault("_tc_cache", {})
        if (kind, terms) not in cache:
            cache[(kind, terms)] = build()
        return cache[(kind, terms)]

    def _tc_dims(self):
        """(N = output-channel tile, cin padded for the kernel, K chunk = channels of one swizzled smem row)."""
        n = next((v for v in (16, 32, 64, 128) if v >= self.cout), 128)
        cin_tc = round_up(self.cin, 16) if self.cin < 64 else round_up(self.cin, 64)
        return n, cin_tc, min(cin_tc, 64)

    def _tc_pack(self, terms=0):
        """One tensor per sub-kernel (transposed-convolution phase): N = cout padded to 16/32/64/128, or tiles of
        128.  Taps are ordered by ascending input offset (transposed sub-kernels are flipped)."""
        def build():
            n, cin_tc, chunk = self._tc_dims()
            n_tiles = -(-self.cout // n)
            packs = []
            for sub in self._subs:
                if self.transposed:
                    sub = sub.flip(2, 3, 4)
                taps = int(sub.shape[2] * sub.shape[3] * sub.shape[4])
                full = torch.zeros(taps, n_tiles * n, cin_tc, device=sub.device, dtype=torch.float32)
                full[:, :self.cout, :self.cin] = sub.permute(2, 3, 4, 0, 1).reshape(taps, self.cout, self.cin)
                full = full.reshape(taps, n_tiles, n, cin_tc // chunk, chunk).permute(1, 3, 0, 2, 4)
                packs.append(_tc_finish(full, terms))
            return packs, n, cin_tc
        return self._tc_cached("plain", terms, build)

    ZFOLD = 2

    def _tc_zfold_ok(self, w_extent, out_pitch):
        """Layers whose channel tile is narrow (N = 16 / 32 leaves the tensor core waiting on its A operand) run
        z-folded: 2 output positions per GEMM row.  Covered: the 7^3 stem (cin <= 16 -> 16, cout 16) and the 3^3
        convolutions with cin in {16, 32} and cout 32."""
        if not (self.nd == 3 and not self.transposed and self.stride == [1, 1, 1] and w_extent % self.ZFOLD == 0):
            return False
        if self.k == [7, 7, 7] and self.padding == [3, 3, 3]:
            return self.cin <= 16 and self.cout == 16 and out_pitch == 16
        if self.k == [3, 3, 3] and self.padding == [1, 1, 1]:
            # only where the folded extent still fills the 8-row bricks (W = 20 would run 10 of 16 rows)
            return (_ZFOLD3 and self.cin in (16, 32) and self.cout == 32 and out_pitch == 32 and w_extent % 16 == 0)
        return False

    def _zfold_factor(self, w_extent, wide_ok):
        """Positions per GEMM row of a z-folded launch: 4 for the 7^3 stem in the 2-K-block split form where the row count
        still fills the 8-row bricks (its N = 4 x 16 kernel exists in that form only, csrc/conv_tc.cu), else 2."""
        if (wide_ok and self.k == [7, 7, 7] and w_extent % 32 == 0 and _ZFOLD4
                and tc_case(self.k, 16, 64, 4) in TC_WIDE_CASES):
            return 4
        return self.ZFOLD

    def _tc_pack_zfold(self, terms=0, F=None):
        """``[1, (K blocks,) 1, kd*kh*(k+F-1), F*cout_p, cin_p]``: per (kd, kh) the k+F-1 windows e, rows (ro, co),
        tap kw = e - ro (zero rows where that falls outside the kernel)."""
        F = self.ZFOLD if F is None else int(F)

        def build():
            k = self.k[2]
            cin_p, cout_p = round_up(self.cin, 16), round_up(self.cout, 16)
            w5 = self._subs[0]                                     # [Cout, Cin, kd, kh, kw]
            full = torch.zeros(self.k[0], self.k[1], k + F - 1, F, cout_p, cin_p, device=w5.device, dtype=torch.float32)
            for e in range(k + F - 1):
                for ro in range(F):
                    kw = e - ro
                    if 0 <= kw < k:
                        full[:, :, e, ro, :self.cout, :self.cin] = w5[:, :, :, :, kw].permute(2, 3, 0, 1)
            return _tc_finish(full.reshape(1, 1, -1, F * cout_p, cin_p), terms)
        return self._tc_cached("zfold%d" % F, terms, build)

    def _tc_stack_ok(self, w_extent, out_pitch):
        """The root net's 1 -> 16 channel 7^3 stem: x taps stacked into channels (1 x 7 x 7 over 7 tap channels)."""
        return (self.nd == 3 and not self.transposed and self.k == [7, 7, 7] and self.stride == [1, 1, 1]
                and self.padding == [3, 3, 3] and self.cin == 1 and self.cout == 16 and out_pitch == 16
                and w_extent % self.ZFOLD == 0)

    def _tc_pack_stack(self, terms=0):
        """``[1, (K blocks,) 1, kh*(k+F-1), F*16, 16]``: K index j = x tap, per kh the k+F-1 z-windows, rows (ro, co)."""
        def build():
            F, k = self.ZFOLD, 7
            w5 = self._subs[0]                                     # [Cout, 1, kd, kh, kw]
            full = torch.zeros(k, k + F - 1, F, 16, 16, device=w5.device, dtype=torch.float32)
            for e in range(k + F - 1):
                for ro in range(F):
                    kw = e - ro
                    if 0 <= kw < k:
                        full[:, e, ro, :self.cout, :k] = w5[:, 0, :, :, kw].permute(2, 0, 1)   # [kh, co, kd]
            return _tc_finish(full.reshape(1, 1, -1, F * 16, 16), terms)
        return self._tc_cached("stack", terms, build)

    def _tc_fused_ok(self, out_pitch, out_dtype):
        """k2/s2 transposed 3-D convolution as ONE launch (all 8 output phases are extra GEMM columns)."""
        esz = 2 if out_dtype == torch.bfloat16 else 4      # staged bytes per value (float32, or a bf16 term pair)
        return (self.transposed and self.nd == 3 and self.k == [2, 2, 2] and (8 * self.cout) % 128 == 0
                and out_pitch == self.cout and (2 * self.cout * esz) % 128 == 0)

    def _tc_pack_fused(self, terms=0):
        """``[n_tiles, (K blocks,) n_chunks, 1, 128, chunk]``: GEMM rows ordered (px, py, pz, co)."""
        def build():
            w = torch.stack([sub.reshape(self.cout, self.cin) for sub in self._subs], 0)   # [8 phases (pd,ph,pw), Cout, Cin]
            _, cin_tc, chunk = self._tc_dims()
            full = torch.zeros(8 * self.cout, cin_tc, device=w.device, dtype=torch.float32)
            full[:, :self.cin] = w.reshape(8 * self.cout, self.cin)
            n_tiles = 8 * self.cout // 128
            full = full.reshape(n_tiles, 128, cin_tc // chunk, chunk).permute(0, 2, 1, 3).unsqueeze(2)
            return _tc_finish(full, terms)
        return self._tc_cached("fused", terms, build)

    def tc_plan(self, w_extent, out_pitch, out_dtype, with_residual):
        """The kernel instantiations ``_call_tc`` would launch for an input whose innermost spatial extent is
        ``w_extent`` (same branch order as ``_call_tc``)."""
        n, cin_tc, _ = self._tc_dims()
        if with_residual is False and self._tc_stack_ok(w_extent, out_pitch):
            return [tc_case([1, 7, 7], 16, 16 * self.ZFOLD, self.ZFOLD)]
        if not self.transposed and cin_tc == round_up(self.cin, 16) and self._tc_zfold_ok(w_extent, out_pitch):
            F = self._zfold_factor(w_extent, _WIDE_SPLIT and _F32_CONV == "bf16x3" and volume_dtype() == torch.float32)
            return [tc_case(self.k, cin_tc, out_pitch * F, F)]
        if not self.transposed:
            return [tc_case(self.k, cin_tc, n)]
        if self._tc_fused_ok(out_pitch, out_dtype):
            return [tc_case([1, 1, 1], cin_tc, 128)]
        return [tc_case(ks, cin_tc, n) for _, _, ks in self.phases]

    def tc_available(self, w_extent, out_pitch, out_dtype, with_residual):
        """``tc_supported()`` and every launch of the plan has a compiled kernel instantiation."""
        return self.tc_supported() and all(c in TC_CASES or (_WIDE_SPLIT and c in TC_WIDE_CASES)     # (wide-only kernels)
                                           for c in self.tc_plan(w_extent, out_pitch, out_dtype, with_residual))

    def _call_tc(self, x, residual, out_pitch, out_dtype, head=None, terms=0, pair_out=False):
        """tcgen05 path.  ``terms`` = 0: ``x`` is bf16 channel-last.  ``terms`` = 3 / 6 (SP3D_CONV_TC_BF16X3): ``x`` is
        float32 channel-last (expanded into bf16 term planes by ``split_bf16`` first) or, ``terms`` = 3, a ``SplitAct``
        (already two term planes); the kernel multiplies the planes with the matching weight terms as extra K blocks of
        the same implicit GEMM.  Output and residual are float32, or ``SplitAct`` with ``pair_out`` (the epilogue writes
        the two term planes of its float32 result: no split pass in front of the next layer)."""
        if not self.tc_supported():
            raise _lib.Sp3dError("convolution shape not covered by the tensor-core path")
        packs, n, cin_tc = self._tc_pack(terms)
        pre_split = isinstance(x, SplitAct)
        N, D, H, W, pitch = [int(v) for v in x.shape]
        o = self.out_shape((D, H, W))
        if terms:
            if head is not None or out_dtype not in (None, torch.float32) or (not pre_split and x.dtype != torch.float32):
                raise _lib.Sp3dError("the split-operand mode takes float32 activations, writes float32 and has no fused head")
            if pre_split and (terms != 3 or pitch != cin_tc):
                raise _lib.Sp3dError("a SplitAct input needs the 3-pair mode and pitch %d (got %d)" % (cin_tc, pitch))
            if pitch < self.cin:
                raise _lib.Sp3dError("activation pitch %d smaller than the input channel count %d" % (pitch, self.cin))
            out_dtype = torch.float32
            planes = 2 if terms == 3 else 3
            x = x.planes if pre_split else split_bf16(x, self.cin, cin_tc, planes)   # [planes, N, D, H, W, cin_tc]
            pitch = cin_tc
            algo = _lib.CONV_TC_BF16X3
            if pair_out:
                if terms != 3:
                    raise _lib.Sp3dError("term-pair output exists in the 3-pair mode only")
                out_pitch = split_pitch(self.cout) if out_pitch is None else int(out_pitch)
            else:
                out_pitch = round_up(self.cout, 4) if out_pitch is None else int(out_pitch)
        else:
            if pair_out:
                raise _lib.Sp3dError("term-pair output needs the split-operand mode")
            if pitch < cin_tc or pitch % 8:
                raise _lib.Sp3dError("bf16 activation pitch %d incompatible with packed cin %d" % (pitch, cin_tc))
            out_dtype = torch.bfloat16 if out_dtype is None else out_dtype
            out_pitch = round_up(self.cout, 16) if out_pitch is None else int(out_pitch)
            planes = 0
            algo = _lib.CONV_TC_BF16
        if head is not None:      # fused soft-argmax head: the volume is never materialised
            if self.k != [1, 1, 1] or self.transposed or self.nd != 3 or residual is not None or self.cout > 15:
                raise _lib.Sp3dError("the fused soft-argmax head needs a 1x1x1 convolution with at most 15 output channels")
            return conv_launch(x, packs[0], self.scale, self.shift, None, (N, o[0], o[1], o[2], 16), cin_tc, self.cout, o,
                               self.k, self.stride, [0, 0, 0], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu,
                               _lib.CONV_TC_BF16, cin_real=self.cin, cout_pitch_w=n, head=head)
        oshape = (N, o[0], o[1], o[2], out_pitch)
        if pair_out:
            out_full = torch.empty((2,) + oshape, device=x.device, dtype=torch.bfloat16)
            out = out_full[0]
            if residual is not None:
                if not isinstance(residual, SplitAct) or tuple(residual.planes.shape) != tuple(out_full.shape):
                    raise _lib.Sp3dError("residual must be a SplitAct of the output's shape")
                residual = residual.planes[0]
        else:
            out = torch.empty(oshape, device=x.device, dtype=out_dtype)
            if residual is not None and (isinstance(residual, SplitAct) or residual.dtype != out_dtype
                                         or residual.shape != out.shape):
                raise _lib.Sp3dError("residual must match the output dtype and shape")
        # kernel view of the activations: [N, D, H, W, pitch] (split mode: plane 0; the kernel steps over the planes
        # through the outer index).  2-D: image batch -> the brick's x axis, [N,1,H,W,C] viewed as [1,N,H,W,C]
        # (a term-pair output / residual is passed as its plane 0: plane 1 follows at the outer index n_outer + n)
        outk, resk = out, residual
        if self.nd == 2:
            xk = x.view(planes, 1, N, H, W, pitch)[0] if planes else x.view(1, N, H, W, pitch)
            if pair_out:
                outk = out_full.view(2, 1, N, o[1], o[2], out_pitch)[0]
                resk = residual.view(1, N, o[1], o[2], out_pitch) if residual is not None else None
            else:
                outk = out.view(1, N, o[1], o[2], out_pitch)
                resk = residual.view(1, N, o[1], o[2], out_pitch) if residual is not None else None
            D, o = N, [N, o[1], o[2]]
        else:
            xk = x[0] if planes else x
        kw = dict(algo=algo, split_terms=terms, pair_out=pair_out)
        tma_out = (out_pitch * (2 if pair_out or out_dtype == torch.bfloat16 else 4)) % 16 == 0

        def wide(case):
            """3 term pairs: take the 2-K-block form (``split_terms`` 2) where the kernel has it for this launch."""
            return terms == 3 and _WIDE_SPLIT and tma_out and case in TC_WIDE_CASES

        if pitch == 16 and residual is None and self._tc_stack_ok(W, out_pitch):
            xs = stack_x_shifts(x.view(-1, D, H, W, pitch), 7, 3)       # every plane: [planes * N, ...]
            t = 2 if wide(tc_case([1, 7, 7], 16, 16 * self.ZFOLD, self.ZFOLD)) else terms
            conv_launch(xs[:N], self._tc_pack_stack(t), self.scale, self.shift, None, outk, 16, self.cout, o, [1, 7, 7],
                        self.stride, [0, -3, -3], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu,
                        cin_real=7, cout_pitch_w=16 * self.ZFOLD, zfold=self.ZFOLD, **dict(kw, split_terms=t))
        elif not self.transposed and pitch == round_up(self.cin, 16) and self._tc_zfold_ok(W, out_pitch):
            F = self._zfold_factor(W, terms == 3 and _WIDE_SPLIT and tma_out)
            t = 2 if wide(tc_case(self.k, pitch, out_pitch * F, F)) else terms
            conv_launch(xk, self._tc_pack_zfold(t, F), self.scale, self.shift, resk, outk, pitch, self.cout, o, self.k,
                        self.stride, [-p for p in self.padding], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu,
                        cin_real=self.cin, cout_pitch_w=out_pitch * F, zfold=F, **dict(kw, split_terms=t))
        elif not self.transposed:
            t = 2 if wide(tc_case(self.k, cin_tc, n)) else terms
            wgt = packs[0] if t == terms else self._tc_pack(2)[0][0]
            conv_launch(xk, wgt, self.scale, self.shift, resk, outk, cin_tc, self.cout, o, self.k, self.stride,
                        [-p for p in self.padding], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu,
                        cin_real=self.cin, cout_pitch_w=n, **dict(kw, split_terms=t))
        elif self._tc_fused_ok(out_pitch, "pair" if pair_out else out_dtype):
            conv_launch(xk, self._tc_pack_fused(terms), self.scale, self.shift, resk, outk, cin_tc, self.cout, (D, H, W),
                        [1, 1, 1], [1, 1, 1], [0, 0, 0], [1, 1, 1], self.stride, [0, 0, 0], self.relu, cin_real=self.cin,
                        cout_pitch_w=128, fused_phases=True, **kw)
        else:
            for wgt, (phase, off0, ks) in zip(packs, self.phases):
                origin = [off0[i] - (ks[i] - 1) for i in range(3)]     # taps ascend from the lowest input offset
                conv_launch(xk, wgt, self.scale, self.shift, resk, outk, cin_tc, self.cout, (D, H, W), ks, [1, 1, 1],
                            origin, [1, 1, 1], self.stride, phase, self.relu, cin_real=self.cin, cout_pitch_w=n, **kw)
        return SplitAct(out_full) if pair_out else out

    def _call_split(self, x, residual, out_pitch, out_dtype, head, out_dims):
        """``x``: ``SplitAct``.  Covered shapes run the 3-pair tensor-core form and return a ``SplitAct`` (float32 with
        ``out_dtype=torch.float32``); anything else goes through the float32 kernel and is re-split."""
        if head is not None or out_dims is not None or out_dtype not in (None, torch.float32):
            raise _lib.Sp3dError("a SplitAct input takes out_dtype None (SplitAct) or torch.float32 only")
        pair_out = out_dtype is None
        w_extent = int(x.shape[3])
        pitch_eff = (split_pitch(self.cout) if pair_out else round_up(self.cout, 4)) if out_pitch is None else int(out_pitch)
        if int(x.shape[4]) == self._tc_dims()[1] and self.tc_available(w_extent, pitch_eff, "pair" if pair_out else torch.float32,
                                                                      residual is not None):
            return self._call_tc(x, residual, out_pitch, None, terms=3, pair_out=pair_out)
        xf = merge_act(x, self.cin)
        rf = None
        if residual is not None:
            rf = merge_act(residual, self.cout) if isinstance(residual, SplitAct) else residual
        y = self.__call__(xf, residual=rf, algo=_lib.CONV_SIMT_F32)
        return split_act(y, self.cout, out_pitch) if pair_out else y

    def __call__(self, x, residual=None, out_pitch=None, algo=None, out_dtype=None, head=None, out_dims=None):
        """``x``: channel-last ``[N,D,H,W,pitch]``.  float32 activations take the float32 SIMT kernel; bf16
        activations the tcgen05 kernel where the shape is covered (``out_dtype`` float32 there gives a float32
        result), else the SIMT kernel with bf16 storage and float32 math.  ``out_dims``: explicit output extent of a
        transposed convolution (``output_padding``; SIMT path) -- how the input gradient of a strided convolution
        recovers the forward input's extent."""
        if isinstance(x, SplitAct):
            return self._call_split(x, residual, out_pitch, out_dtype, head, out_dims)
        if algo is None:
            if x.dtype == torch.bfloat16 and self.tc_supported():
                algo = _lib.CONV_TC_BF16
            elif (x.dtype == torch.float32 and _F32_CONV != "simt" and head is None and out_dims is None
                  and out_dtype in (None, torch.float32)
                  and self.tc_available(int(x.shape[3]), round_up(self.cout, 4) if out_pitch is None else int(out_pitch),
                                        torch.float32, residual is not None)):
                algo = _lib.CONV_TC_BF16X3
            else:
                algo = _lib.CONV_SIMT_F32
        if algo == _lib.CONV_TC_BF16:
            return self._call_tc(x, residual, out_pitch, out_dtype, head=head)
        if algo == _lib.CONV_TC_BF16X3:
            return self._call_tc(x, residual, out_pitch, out_dtype, terms=6 if _F32_CONV == "bf16x6" else 3)
        if head is not None:
            raise _lib.Sp3dError("the fused soft-argmax head exists on the tensor-core path only")
        N, D, H, W, pitch = [int(v) for v in x.shape]
        if pitch < self.cin_p:
            raise _lib.Sp3dError("activation pitch %d smaller than packed cin %d" % (pitch, self.cin_p))
        o = self.out_shape((D, H, W))
        if out_dims is not None:
            if not self.transposed or any(not (0 <= int(out_dims[i]) - o[i] < self.stride[i]) for i in range(3)):
                raise _lib.Sp3dError("out_dims must extend a transposed convolution's output by less than the stride")
            o = [int(v) for v in out_dims]
        if out_dtype is None:
            out_dtype = x.dtype
        if out_pitch is None:
            out_pitch = round_up(self.cout, 4 if out_dtype == torch.float32 else 16)
        empty_phase = self.transposed and any(w is None for w in self.weights)
        if empty_phase and (self.shift is not None or residual is not None):
            raise _lib.Sp3dError("transposed convolutions with tap-less phases are supported without shift / residual")
        alloc = torch.zeros if empty_phase else torch.empty
        out = alloc((N, o[0], o[1], o[2], int(out_pitch)), device=x.device, dtype=out_dtype)
        if not self.transposed:
            conv_launch(x, self.weights[0], self.scale, self.shift, residual, out, self.cin_p, self.cout, o, self.k,
                        self.stride, [-p for p in self.padding], [1, 1, 1], [1, 1, 1], [0, 0, 0], self.relu, algo,
                        cin_real=self.cin)
        else:
            for wgt, (phase, off0, ks) in zip(self.weights, self.phases):
                grid = [(o[i] - phase[i] + self.stride[i] - 1) // self.stride[i] for i in range(3)]
                if wgt is None or min(grid) <= 0:
                    continue
                conv_launch(x, wgt, self.scale, self.shift, residual, out, self.cin_p, self.cout, grid, ks,
                            [1, 1, 1], off0, [-1, -1, -1], self.stride, phase, self.relu, algo, cin_real=self.cin)
        return out
