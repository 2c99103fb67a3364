#!/usr/bin/env python
"""Run one of the reference's UNMODIFIED tools (tools/evaluate.py, validate_3d.py, train_3d.py) against this backend.

    python integration/run_tool.py /path/to/SelfPose3d/tools/evaluate.py --cfg <yaml> --with-ssv --test-file <ckpt>

The tools do `import _init_paths` (which puts <ref>/lib first on sys.path) and then `import models`
(tools/_init_paths.py:15-23, tools/evaluate.py:24-31), so PYTHONPATH alone cannot out-rank the reference's own
`lib/models`.  This launcher pre-registers `selfpose3d_b200.models` (and its sub-modules) in `sys.modules` under
the names `models`, `models.project_layer`, ... before handing control to the tool with `runpy`; everything else
(`core`, `dataset`, `utils`) is the reference's own code, untouched.  Nothing from the reference is copied here.
"""
import importlib
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SUBMODULES = ["pose_resnet", "v2v_net", "project_layer", "cuboid_proposal_net", "cuboid_proposal_net_soft",
              "pose_regression_net", "multi_person_posenet", "multi_person_posenet_ssv"]


def register_backend():
    pkg = importlib.import_module("selfpose3d_b200.models")
    sys.modules["models"] = pkg
    for name in SUBMODULES:
        sys.modules["models." + name] = importlib.import_module("selfpose3d_b200.models." + name)
    return pkg


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        sys.exit(2)
    tool = os.path.abspath(sys.argv[1])
    register_backend()
    sys.argv = [tool] + sys.argv[2:]
    sys.path.insert(0, os.path.dirname(tool))       # so that `import _init_paths` resolves
    runpy.run_path(tool, run_name="__main__")


if __name__ == "__main__":
    main()
