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
    """``run_tool.py [--sp3d-shims] [--sp3d-synthetic-data] <tool.py> <tool arguments...>``

    ``--sp3d-shims``: register stand-ins for the pip packages the tools import and the image lacks
    (``integration/shims.py``); ``--sp3d-synthetic-data``: make ``dataset.panoptic_synth[_ssv]`` resolvable
    (``integration/synth_panoptic.py``) for a YAML whose ``DATASET.*_DATASET`` names them.  Both are harness switches
    for boxes without the pip packages / the Panoptic data; the tool itself runs unmodified."""
    argv = sys.argv[1:]
    shims = synthetic_data = False
    while argv and argv[0].startswith("--sp3d-"):
        flag = argv.pop(0)
        shims |= flag == "--sp3d-shims"
        synthetic_data |= flag == "--sp3d-synthetic-data"
    if not argv:
        print(__doc__)
        sys.exit(2)
    tool = os.path.abspath(argv[0])
    if shims:
        from integration import shims as _shims
        _shims.install()
    register_backend()
    sys.argv = [tool] + argv[1:]
    sys.path.insert(0, os.path.dirname(tool))       # so that `import _init_paths` resolves
    if synthetic_data:
        lib = os.path.join(os.path.dirname(os.path.dirname(tool)), "lib")     # what tools/_init_paths.py will add
        if lib not in sys.path:
            sys.path.insert(0, lib)
        import dataset                                # the reference's own package (unmodified)
        from integration import synth_panoptic
        synth_panoptic.register(dataset)
    runpy.run_path(tool, run_name="__main__")


if __name__ == "__main__":
    main()
