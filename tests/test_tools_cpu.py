"""The reference's UNMODIFIED tools against this backend, on CPU with every kernel entry point emulated
(tests/run_tool_emulated.py): ``tools/evaluate.py`` and one epoch of ``tools/train_3d.py`` (self-supervised
configuration: two training iterations, validation, checkpoint) through ``integration/run_tool.py`` -- the launcher
that registers ``selfpose3d_b200.models`` as the tools' ``models`` package -- with the harness stand-ins for the
missing pip packages (integration/shims.py) and the synthetic ``panoptic_synth*`` datasets
(integration/synth_panoptic.py).  The tool sources come from the staged reference (oracle/_ref, git-ignored, staged
by oracle/make_ref.sh); the GPU twin is tests/test_gpu_tools.py."""
import os
import subprocess

import pytest

import tools_harness as H

pytestmark = pytest.mark.skipif(not H.have_reference(), reason="oracle/_ref not staged (sh oracle/make_ref.sh)")


def _run(cmd, cwd):
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, timeout=1500)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    return res.stdout + res.stderr


def test_evaluate_py_runs_unchanged(tmp_path):
    y = H.write_yaml(str(tmp_path / "synth_ssv.yaml"), str(tmp_path / "out"), ssl=True)
    H.write_checkpoint(str(tmp_path / "ckpt.pth"), y)
    out = _run(H.command("evaluate.py", "--cfg", str(tmp_path / "synth_ssv.yaml"), "--with-ssv", "--test-file",
                         str(tmp_path / "ckpt.pth"), emulate=True), str(tmp_path))
    assert "=> load models state" in out and "Type: pose" in out and "Type: root" in out and "MPJPE" in out


def test_train_3d_py_runs_one_epoch_unchanged(tmp_path):
    H.write_yaml(str(tmp_path / "synth_ssv.yaml"), str(tmp_path / "out"), ssl=True)
    out = _run(H.command("train_3d.py", "--cfg", str(tmp_path / "synth_ssv.yaml"), emulate=True), str(tmp_path))
    assert "Epoch: [0][1/2]" in out and "loss_pose3d_ssv" in out          # two training iterations were logged
    assert "Test: [1/2]" in out and "mpjpe@500mm" in out                    # validate_3d ran on the synthetic frames
    assert os.path.isfile(str(tmp_path / "out" / "synth_ssv" / "final_state.pth.tar"))
