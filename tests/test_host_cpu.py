"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol,
the product's own affine/camera helpers agree with the reference goldens, the camera table packs
what the kernel expects, and the module tree exposes the reference's state-dict keys."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN
from selfpose3d_b200 import _lib, ops, synthetic
from selfpose3d_b200.config import default_config
from selfpose3d_b200.models import multi_person_posenet, multi_person_posenet_ssv, v2v_net
from selfpose3d_b200.utils import cameras, transforms


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sp3d.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(sp3d_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sp3d_abi_version() == _lib.ABI_VERSION == 4
    assert lib.sp3d_strerror(-1) == b"invalid argument"


def test_struct_sizes_match_header_layout():
    # spot checks of the ctypes mirrors against the C layout (LP64): pointer arrays, int64 and double alignment
    assert ctypes.sizeof(_lib.UnprojectArgs) % 8 == 0
    assert _lib.UnprojectArgs.heatmaps.size == 8 * _lib.MAX_VIEWS
    assert _lib.NmsTopkArgs.space_size.offset % 8 == 0
    assert _lib.ConvArgs.ksize.size == 12


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    a = _lib.UnprojectArgs()
    assert lib.sp3d_unproject_fwd(ctypes.byref(a), None) == 0      # n_cubes == 0: nothing to do
    a.n_cubes = 1
    assert lib.sp3d_unproject_fwd(ctypes.byref(a), None) == -1
    n = _lib.NmsTopkArgs()
    assert lib.sp3d_nms_topk3d(ctypes.byref(n), None) == -1
    c = _lib.ConvArgs()
    assert lib.sp3d_conv_fwd(ctypes.byref(c), None) == -1
    with pytest.raises(_lib.Sp3dError):
        _lib.check(-2, "probe")


def test_product_affine_matches_reference_golden(golden):
    g = golden("affine")
    for case, want in zip(g["cases"], g["trans"]):
        got = transforms.get_affine_transform(case[0:2], case[2:4].astype(np.float32), case[4], case[5:7])
        assert np.array_equal(got, want)
    np.testing.assert_allclose(transforms.get_scale((1920, 1080), (288, 384)), [9.6, 12.8], rtol=1e-7)


def test_product_project_pose_matches_reference_golden(golden):
    g = golden("project_pose")
    x = torch.from_numpy(g["points"])
    for v in range(g["pixels"].shape[0]):
        cam = {k[4:]: torch.from_numpy(np.asarray(g[k][v])) for k in g if k.startswith("cam_")}
        got = cameras.project_pose(x, cam).numpy()
        np.testing.assert_allclose(got, g["pixels"][v], rtol=3e-6, atol=2e-3)


def test_pack_cameras_layout(golden):
    g = golden("project_layer_pose")
    V, B = g["center"].shape[:2]
    meta = [{"center": torch.from_numpy(g["center"][c]), "scale": torch.from_numpy(g["scale"][c]),
             "rotation": torch.from_numpy(g["rotation"][c]),
             "camera": {k[4:]: torch.from_numpy(g[k][c]) for k in g if k.startswith("cam_")}} for c in range(V)]
    t = ops.pack_cameras(meta, g["image_size"], torch.from_numpy(g["flip"])).numpy()
    assert t.shape == (B, V, 32) and t.dtype == np.float32
    np.testing.assert_array_equal(t[1, 2, 0:9], g["cam_R"][2, 1].reshape(9).astype(np.float32))
    np.testing.assert_array_equal(t[1, 2, 9:12], g["cam_T"][2, 1].reshape(3).astype(np.float32))
    assert t[0, 0, 27] == 1920.0 and t[0, 0, 28] == 1080.0
    assert t[2, 0, 29] == 1.0 and t[0, 0, 29] == 0.0
    from oracle import geometry
    want = geometry.get_affine_transform(g["center"][3, 1], g["scale"][3, 1], g["rotation"][3, 1], g["image_size"])
    np.testing.assert_array_equal(t[1, 3, 21:27], want.astype(np.float32).reshape(6))


def test_state_dict_keys_match_reference():
    """Key names and shapes recorded from the reference models by tests/golden/make_golden.py."""
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    cfg = default_config()
    cfg.WITH_ATTN = True
    ssv = multi_person_posenet_ssv.get_multi_person_pose_net(cfg, is_train=False)
    ours = {k: list(v.shape) for k, v in ssv.state_dict().items()}
    assert ours == ref["multi_person_posenet_ssv_attn"]
    assert len(ours) == 790
    cfg = default_config()
    cfg.NETWORK.ROOTNET_ROOTHM = False
    sup = multi_person_posenet.get_multi_person_pose_net(cfg, is_train=False)
    assert {k: list(v.shape) for k, v in sup.state_dict().items()} == ref["multi_person_posenet"]


def test_reference_init_statistics():
    net = v2v_net.V2VNet(15, 15)
    w = net.front_layers[0].block[0].weight
    assert abs(float(w.std()) - 1e-3) < 1e-4 and float(net.output_layer.bias.abs().max()) == 0.0


def test_synthetic_is_deterministic():
    a = synthetic.ring_cameras(5, seed=0)
    b = synthetic.ring_cameras(5, seed=0)
    assert all(np.array_equal(x["R"], y["R"]) for x, y in zip(a, b))
    m = v2v_net.V2VNet(1, 1)
    s1 = synthetic.trained_like_state_dict(m, seed=3)
    s2 = synthetic.trained_like_state_dict(m, seed=3)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)


def test_pack_cameras_affine_cache_is_content_keyed():
    """The input affine is cached by the bytes of (center, scale, rotation, input size): a different rig or
    augmentation must never see a stale entry, and a repeated call returns identical tables."""
    from selfpose3d_b200 import ops, synthetic
    from selfpose3d_b200.utils.transforms import get_affine_transform
    cams = synthetic.ring_cameras(3, seed=4)
    m1 = synthetic.make_meta(cams, 2, (96, 128))
    m2 = synthetic.make_meta(cams, 2, (96, 128), rotation=[[7.0, -3.0]] * 3, scale_mul=[[1.2, 0.8]] * 3)
    t1a, t2, t1b = ops.pack_cameras(m1, (96, 128)), ops.pack_cameras(m2, (96, 128)), ops.pack_cameras(m1, (96, 128))
    assert torch.equal(t1a, t1b) and not torch.equal(t1a[..., 21:27], t2[..., 21:27])
    for meta, table in ((m1, t1a), (m2, t2)):
        for c, m in enumerate(meta):
            for i in range(2):
                want = get_affine_transform(np.asarray(m["center"][i]), np.asarray(m["scale"][i]),
                                            np.asarray(m["rotation"][i]), (96, 128)).reshape(6).astype(np.float32)
                assert np.array_equal(table[i, c, 21:27].numpy(), want)
    # same rig at another network input size: other affine
    t3 = ops.pack_cameras(m1, (192, 256))
    assert not torch.equal(t1a[..., 21:27], t3[..., 21:27])


def test_ctypes_struct_layouts_match_the_c_header(tmp_path):
    """Every argument struct of include/sp3d.h has the same size under gcc as its ctypes mirror in _lib.py (field
    order / padding drift would silently corrupt launches)."""
    import ctypes
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no C compiler")
    pairs = {
        "sp3d_unproject_args": _lib.UnprojectArgs, "sp3d_heatmaps_f16_args": _lib.HeatmapsF16Args,
        "sp3d_unproject_finalize_args": _lib.UnprojectFinalizeArgs, "sp3d_nms_topk_args": _lib.NmsTopkArgs,
        "sp3d_softargmax_args": _lib.SoftargmaxArgs, "sp3d_conv_args": _lib.ConvArgs, "sp3d_maxpool_args": _lib.MaxpoolArgs,
        "sp3d_layout_args": _lib.LayoutArgs, "sp3d_s2d_args": _lib.S2DArgs, "sp3d_stack_args": _lib.StackArgs,
        "sp3d_split_args": _lib.SplitArgs, "sp3d_unproject_bwd_args": _lib.UnprojectBwdArgs,
        "sp3d_softargmax_bwd_args": _lib.SoftargmaxBwdArgs, "sp3d_maxpool_bwd_args": _lib.MaxpoolBwdArgs,
        "sp3d_conv_wgrad_args": _lib.ConvWgradArgs, "sp3d_bn_stats_args": _lib.BnStatsArgs,
        "sp3d_bn_apply_args": _lib.BnApplyArgs, "sp3d_bn_bwd_args": _lib.BnBwdArgs, "sp3d_relu_bwd_args": _lib.ReluBwdArgs,
        "sp3d_gauss_render_args": _lib.GaussRenderArgs, "sp3d_gauss_render_bwd_args": _lib.GaussRenderBwdArgs,
        "sp3d_target_heatmaps_args": _lib.TargetHeatmapsArgs, "sp3d_target_volume_args": _lib.TargetVolumeArgs,
        "sp3d_conv_wgrad_tc_args": _lib.ConvWgradTcArgs,
    }
    header = open(os.path.join(ROOT, "include", "sp3d.h")).read()
    declared = set(re.findall(r"}\s*(sp3d_[a-z0-9_]+_args)\s*;", header))
    assert declared == set(pairs), declared ^ set(pairs)
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "sp3d.h"\nint main(void){'
                   + "".join('printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in pairs) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().strip().splitlines()
    for line in out:
        name, size = line.split()
        assert ctypes.sizeof(pairs[name]) == int(size), (name, ctypes.sizeof(pairs[name]), int(size))


def test_tc_case_table_matches_the_kernel_instantiations():
    """ops.TC_CASES (what the host believes is compiled) against the SP3D_TC_CASE list of csrc/conv_tc.cu."""
    src = open(os.path.join(ROOT, "selfpose3d_b200", "csrc", "conv_tc.cu")).read()
    body = src[src.index("int conv_tc(const sp3d_conv_args* a, cudaStream_t st)"):]
    cases = set()
    for m in re.finditer(r"^\s*SP3D_TC_CASE(_F)?\(([^)]*)\)\s*$", body, flags=re.M):
        if not re.fullmatch(r"[\d,\s]+", m.group(2)):
            continue                                   # the macro definitions themselves
        v = [int(t) for t in m.group(2).split(",")]
        cases.add((v[0], v[1], v[2], v[3], v[11] if m.group(1) else 1))
    # SP3D_TC_CASE_P: the same 13 parameters as _W, also compiled as CTA pairs (CL = 2, behind sp3d_debug_conv_pair)
    pairable = [[int(t) for t in m.group(1).split(",")] for m in re.finditer(r"^\s*SP3D_TC_CASE_P\(([\d,\s]+)\)\s*$", body, flags=re.M)]
    cases |= {(v[0], v[1], v[2], v[3], v[11]) for v in pairable if v[12] == 1}
    assert len(cases) >= 20
    assert cases == set(ops.TC_CASES), cases ^ set(ops.TC_CASES)
    # the 2-K-block form of the 3-pair split mode (WD = 2 instantiations) against ops.TC_WIDE_CASES
    wide = set()
    for m in re.finditer(r"^\s*SP3D_TC_CASE_W\(([\d,\s]+)\)\s*$", body, flags=re.M):
        v = [int(t) for t in m.group(1).split(",")]
        assert v[12] == 2
        wide.add((v[0], v[1], v[2], v[3], v[11]))
    wide |= {(v[0], v[1], v[2], v[3], v[11]) for v in pairable if v[12] == 2}
    assert wide == set(ops.TC_WIDE_CASES), wide ^ set(ops.TC_WIDE_CASES)
