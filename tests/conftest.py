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
