import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # keep libsp3d.so in step with its sources (make is a no-op when it is up to date)
    import shutil
    if shutil.which("nvcc") and shutil.which("make"):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    return load


def cams_from_arrays(d, prefix="cam_"):
    """``{key: [V,B,...]}`` camera arrays of a golden file -> nested ``[V][B]`` dicts."""
    keys = [k for k in d if k.startswith(prefix)]
    V, B = d[keys[0]].shape[:2]
    return [[{k[len(prefix):]: d[k][c][i] for k in keys} for i in range(B)] for c in range(V)]


def cam_arrays(d, prefix="cam_"):
    return {k[len(prefix):]: d[k] for k in d if k.startswith(prefix)}


# Modules written against the float32 FMA ("simt") kernels as the bit-level parity path; everything else runs in the
# product's default mode (float32-faithful tensor-core convolutions, ``ops.set_float32_conv("bf16x3")``).
SIMT_MODULES = ("test_gpu_parity", "test_gpu_backward", "test_gpu_tensorcore", "test_gpu_zz_training_step",
                "test_gpu_pair", "test_gpu_split", "test_conv_lowering_cpu", "test_autograd_cpu", "test_training_cpu",
                "test_dist_gloo")


@pytest.fixture(autouse=True)
def _float32_conv_mode(request):
    from selfpose3d_b200 import ops
    import torch
    name = request.module.__name__.rsplit(".", 1)[-1]
    before = ops.float32_conv()
    ops.set_float32_conv("simt" if name in SIMT_MODULES else "bf16x3")
    ops.set_volume_dtype(torch.float32)
    yield
    ops.set_float32_conv(before)
