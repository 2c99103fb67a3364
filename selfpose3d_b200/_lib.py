"""ctypes binding of ``libsp3d.so`` (the C ABI declared in ``include/sp3d.h``).

There is no fallback: if the shared library is missing or a call returns a
non-zero status, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsp3d.so")

ABI_VERSION = 4      # SP3D_ABI_VERSION of include/sp3d.h these bindings were written against
MAX_VIEWS = 8
CAM_FLOATS = 32
F32, BF16, F16, BF16X2 = 0, 1, 2, 3
CONV_SIMT_F32, CONV_TC_BF16, CONV_TC_BF16X3 = 0, 1, 2

_i3 = C.c_int * 3


class UnprojectArgs(C.Structure):
    _fields_ = [
        ("heatmaps", C.c_void_p * MAX_VIEWS),
        ("hm_stride_b", C.c_int64), ("hm_stride_c", C.c_int64), ("hm_stride_h", C.c_int64), ("hm_stride_w", C.c_int64),
        ("cams", C.c_void_p), ("centers", C.c_void_p),
        ("center_stride", C.c_int), ("check_flag", C.c_int), ("cubes_per_sample", C.c_int),
        ("cube_sample", C.c_void_p),
        ("lin_x", C.c_void_p), ("lin_y", C.c_void_p), ("lin_z", C.c_void_p),
        ("B", C.c_int), ("V", C.c_int), ("C", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("n_cubes", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int),
        ("img_w", C.c_float), ("img_h", C.c_float), ("hm_cfg_w", C.c_float), ("hm_cfg_h", C.c_float),
        ("view_begin", C.c_int), ("view_end", C.c_int), ("partial", C.c_int),
        ("cubes", C.c_void_p), ("out_dtype", C.c_int),
        ("out_stride_cube", C.c_int64), ("out_stride_c", C.c_int64), ("out_stride_vox", C.c_int64),
        ("out_c_pad", C.c_int),
        ("grids", C.c_void_p),
        ("hm_dtype", C.c_int), ("math_mode", C.c_int),
    ]


class HeatmapsF16Args(C.Structure):
    _fields_ = [
        ("heatmaps", C.c_void_p * MAX_VIEWS),
        ("stride_b", C.c_int64), ("stride_c", C.c_int64), ("stride_h", C.c_int64), ("stride_w", C.c_int64),
        ("V", C.c_int), ("B", C.c_int), ("C", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("out", C.c_void_p),
    ]


class UnprojectFinalizeArgs(C.Structure):
    _fields_ = [
        ("buf", C.c_void_p), ("n_cubes", C.c_int64), ("C", C.c_int64), ("N", C.c_int64),
        ("stride_cube", C.c_int64), ("stride_c", C.c_int64), ("stride_vox", C.c_int64),
    ]


class NmsTopkArgs(C.Structure):
    _fields_ = [
        ("root_cubes", C.c_void_p),
        ("B", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int), ("K", C.c_int),
        ("threshold", C.c_float),
        ("space_size", C.c_double * 3), ("space_center", C.c_double * 3),
        ("loc_f64", C.c_int),
        ("grid_centers", C.c_void_p), ("topk_index", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class SoftargmaxArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_dtype", C.c_int),
        ("stride_cube", C.c_int64), ("stride_c", C.c_int64), ("stride_vox", C.c_int64),
        ("n_cubes", C.c_int), ("C", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int),
        ("centers", C.c_void_p), ("center_stride", C.c_int), ("check_flag", C.c_int),
        ("lin_x", C.c_void_p), ("lin_y", C.c_void_p), ("lin_z", C.c_void_p),
        ("beta", C.c_float),
        ("out", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p), ("weight", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("residual", C.c_void_p), ("out", C.c_void_p),
        ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int), ("cin", C.c_int), ("cin_pitch", C.c_int),
        ("OD", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("TD", C.c_int), ("TH", C.c_int), ("TW", C.c_int),
        ("cout", C.c_int), ("cout_pitch", C.c_int), ("cout_pitch_w", C.c_int),
        ("ksize", _i3), ("stride", _i3), ("tap_off0", _i3), ("tap_step", _i3), ("ostride", _i3), ("ooffset", _i3),
        ("relu", C.c_int), ("algo", C.c_int), ("in_dtype", C.c_int), ("out_dtype", C.c_int),
        ("fused_phases", C.c_int), ("zfold", C.c_int),
        ("head_softargmax", C.POINTER(SoftargmaxArgs)),
        ("split_terms", C.c_int),
    ]


class MaxpoolArgs(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p), ("out", C.c_void_p),
        ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("c_pitch", C.c_int),
        ("OD", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("k", _i3), ("s", _i3), ("p", _i3),
        ("dtype", C.c_int),
    ]


class S2DArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("src_dtype", C.c_int),
        ("stride_n", C.c_int64), ("stride_c", C.c_int64), ("stride_y", C.c_int64), ("stride_x", C.c_int64),
        ("N", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("dst_pitch", C.c_int), ("dst_dtype", C.c_int),
    ]


class SplitArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("P", C.c_int64),
        ("C", C.c_int), ("src_pitch", C.c_int), ("c_block", C.c_int), ("S", C.c_int),
    ]


class StackArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p),
        ("N", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int),
        ("src_pitch", C.c_int), ("taps", C.c_int), ("pad", C.c_int),
    ]


class LayoutArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p),
        ("N", C.c_int64), ("C", C.c_int64), ("S", C.c_int64), ("c_pitch", C.c_int64),
        ("to_channel_last", C.c_int), ("src_dtype", C.c_int), ("dst_dtype", C.c_int),
    ]


class UnprojectBwdArgs(C.Structure):
    _fields_ = [("fwd", UnprojectArgs), ("grad_cubes", C.c_void_p), ("grad_heatmaps", C.c_void_p * MAX_VIEWS)]


class SoftargmaxBwdArgs(C.Structure):
    _fields_ = [("fwd", SoftargmaxArgs), ("grad_out", C.c_void_p), ("grad_x", C.c_void_p)]


class MaxpoolBwdArgs(C.Structure):
    _fields_ = [("fwd", MaxpoolArgs), ("grad_out", C.c_void_p), ("grad_in", C.c_void_p)]


class ConvWgradArgs(C.Structure):
    _fields_ = [("fwd", ConvArgs), ("grad_out", C.c_void_p), ("grad_weight", C.c_void_p), ("grad_bias", C.c_void_p)]


class BnStatsArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("P", C.c_int64), ("C", C.c_int), ("pitch", C.c_int),
                ("mean", C.c_void_p), ("var", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("n_items", C.c_int), ("n_groups", C.c_int), ("item_group", C.c_void_p), ("group_items", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("momentum", C.c_float)]


class BnApplyArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("residual", C.c_void_p), ("y", C.c_void_p),
                ("P", C.c_int64), ("C", C.c_int), ("pitch", C.c_int),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("relu", C.c_int),
                ("n_items", C.c_int), ("n_groups", C.c_int), ("item_group", C.c_void_p)]


class BnBwdArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("grad_y", C.c_void_p), ("y", C.c_void_p),
                ("P", C.c_int64), ("C", C.c_int), ("pitch", C.c_int),
                ("mean", C.c_void_p), ("var", C.c_void_p), ("gamma", C.c_void_p), ("eps", C.c_float),
                ("grad_x", C.c_void_p), ("grad_gamma", C.c_void_p), ("grad_beta", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("n_items", C.c_int), ("n_groups", C.c_int), ("item_group", C.c_void_p), ("group_items", C.c_void_p)]


class ReluBwdArgs(C.Structure):
    _fields_ = [("grad_y", C.c_void_p), ("y", C.c_void_p), ("grad_x", C.c_void_p), ("n", C.c_int64)]


class GaussRenderArgs(C.Structure):
    _fields_ = [("kps", C.c_void_p), ("n_people", C.c_void_p),
                ("V", C.c_int), ("B", C.c_int), ("P", C.c_int), ("J", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("inv_scale", C.c_float), ("sigma", C.c_float), ("heatmaps", C.c_void_p)]


class GaussRenderBwdArgs(C.Structure):
    _fields_ = [("fwd", GaussRenderArgs), ("grad_heatmaps", C.c_void_p), ("grad_kps", C.c_void_p)]


class ConvWgradTcArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("grad_out", C.c_void_p),
                ("N", C.c_int), ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int),
                ("cin", C.c_int), ("x_pitch", C.c_int), ("cout", C.c_int), ("g_pitch", C.c_int),
                ("ksize", C.c_int * 3), ("tap_off", C.c_int * 3), ("GX", C.c_int), ("GY", C.c_int), ("GZ", C.c_int),
                ("g_stride", C.c_int * 3), ("g_off", C.c_int * 3),
                ("grad_weight", C.c_void_p), ("gw_cin", C.c_int), ("gw_pitch", C.c_int), ("grad_bias", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class TargetHeatmapsArgs(C.Structure):
    _fields_ = [("joints", C.c_void_p), ("joints_vis", C.c_void_p), ("n_people", C.c_void_p),
                ("n_items", C.c_int), ("P", C.c_int), ("J", C.c_int), ("jstride", C.c_int), ("vstride", C.c_int),
                ("h", C.c_int), ("w", C.c_int), ("stride_x", C.c_double), ("stride_y", C.c_double),
                ("window", C.c_void_p), ("radius", C.c_int), ("target", C.c_void_p), ("target_weight", C.c_void_p)]


class TargetVolumeArgs(C.Structure):
    _fields_ = [("roots", C.c_void_p), ("n_people", C.c_void_p), ("n_items", C.c_int), ("P", C.c_int),
                ("grid_x", C.c_void_p), ("grid_y", C.c_void_p), ("grid_z", C.c_void_p),
                ("X", C.c_int), ("Y", C.c_int), ("Z", C.c_int), ("sigma", C.c_double), ("target", C.c_void_p)]


# every symbol include/sp3d.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sp3d_abi_version": (C.c_int, []),
    "sp3d_strerror": (C.c_char_p, [C.c_int]),
    "sp3d_last_cuda_error": (C.c_char_p, []),
    "sp3d_unproject_fwd": (C.c_int, [C.POINTER(UnprojectArgs), C.c_void_p]),
    "sp3d_heatmaps_to_f16": (C.c_int, [C.POINTER(HeatmapsF16Args), C.c_void_p]),
    "sp3d_unproject_finalize": (C.c_int, [C.POINTER(UnprojectFinalizeArgs), C.c_void_p]),
    "sp3d_nms_topk3d": (C.c_int, [C.POINTER(NmsTopkArgs), C.c_void_p]),
    "sp3d_nms_topk3d_workspace": (C.c_int64, [C.POINTER(NmsTopkArgs)]),
    "sp3d_softargmax3d_workspace": (C.c_int64, [C.POINTER(SoftargmaxArgs)]),
    "sp3d_softargmax3d_fwd": (C.c_int, [C.POINTER(SoftargmaxArgs), C.c_void_p]),
    "sp3d_conv_fwd": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "sp3d_conv_head_workspace": (C.c_int64, [C.POINTER(ConvArgs)]),
    "sp3d_debug_conv_profile": (None, [C.c_void_p]),
    "sp3d_debug_conv_pair": (None, [C.c_int]),
    "sp3d_maxpool_fwd": (C.c_int, [C.POINTER(MaxpoolArgs), C.c_void_p]),
    "sp3d_layout_convert": (C.c_int, [C.POINTER(LayoutArgs), C.c_void_p]),
    "sp3d_space_to_depth": (C.c_int, [C.POINTER(S2DArgs), C.c_void_p]),
    "sp3d_stack_x_shifts": (C.c_int, [C.POINTER(StackArgs), C.c_void_p]),
    "sp3d_split_bf16": (C.c_int, [C.POINTER(SplitArgs), C.c_void_p]),
    "sp3d_merge_bf16": (C.c_int, [C.POINTER(SplitArgs), C.c_void_p]),
    "sp3d_unproject_bwd": (C.c_int, [C.POINTER(UnprojectBwdArgs), C.c_void_p]),
    "sp3d_softargmax3d_bwd": (C.c_int, [C.POINTER(SoftargmaxBwdArgs), C.c_void_p]),
    "sp3d_maxpool_bwd": (C.c_int, [C.POINTER(MaxpoolBwdArgs), C.c_void_p]),
    "sp3d_conv_wgrad": (C.c_int, [C.POINTER(ConvWgradArgs), C.c_void_p]),
    "sp3d_bn_stats": (C.c_int, [C.POINTER(BnStatsArgs), C.c_void_p]),
    "sp3d_bn_apply": (C.c_int, [C.POINTER(BnApplyArgs), C.c_void_p]),
    "sp3d_bn_bwd": (C.c_int, [C.POINTER(BnBwdArgs), C.c_void_p]),
    "sp3d_relu_bwd": (C.c_int, [C.POINTER(ReluBwdArgs), C.c_void_p]),
    "sp3d_gauss_render_fwd": (C.c_int, [C.POINTER(GaussRenderArgs), C.c_void_p]),
    "sp3d_gauss_render_bwd": (C.c_int, [C.POINTER(GaussRenderBwdArgs), C.c_void_p]),
    "sp3d_conv_wgrad_tc": (C.c_int, [C.POINTER(ConvWgradTcArgs), C.c_void_p]),
    "sp3d_conv_wgrad_tc_workspace": (C.c_int64, [C.POINTER(ConvWgradTcArgs)]),
    "sp3d_target_heatmaps": (C.c_int, [C.POINTER(TargetHeatmapsArgs), C.c_void_p]),
    "sp3d_target_volume": (C.c_int, [C.POINTER(TargetVolumeArgs), C.c_void_p]),
}

_lib = None
launch_count = 0  # kernels launched through the C ABI by this process (bench.py reports it)


class Sp3dError(RuntimeError):
    pass


def load():
    """Load ``libsp3d.so`` (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sp3dError(
            "libsp3d.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C selfpose3d_b200/csrc`; there is no CPU or PyTorch fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sp3d_abi_version.restype = C.c_int
    if lib.sp3d_abi_version() != ABI_VERSION:   # checked first: a stale library would otherwise fail on a missing symbol
        raise Sp3dError("libsp3d.so ABI version mismatch (library %d, bindings %d): rebuild with `make -C "
                        "selfpose3d_b200/csrc`" % (lib.sp3d_abi_version(), ABI_VERSION))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    if os.environ.get("SP3D_CONV_PAIR") in ("0", "1"):     # A/B switch of the CTA-pair weight multicast (ops.py)
        lib.sp3d_debug_conv_pair(int(os.environ["SP3D_CONV_PAIR"]))
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.sp3d_strerror(status).decode()
        cuda = lib.sp3d_last_cuda_error().decode()
        raise Sp3dError("%s failed: %s%s" % (what, msg, (" [" + cuda + "]") if cuda else ""))


def call(name, args, stream, launches=1, kind=None, work=0.0, detail=None):
    """Invoke ``sp3d_<name>(&args, stream)`` and raise on a non-zero status.

    ``kind`` / ``work`` tag the launch (kernel family, algorithmic FLOPs or bytes) for
    ``selfpose3d_b200.profiler`` when it is enabled."""
    global launch_count
    from . import profiler
    lib = load()
    start = profiler.begin() if profiler.active() else None
    status = getattr(lib, name)(C.byref(args), C.c_void_p(stream))
    check(status, name)
    launch_count += launches
    if start is not None:
        profiler.end(kind or name, start, work, detail)
