"""GPU: the reference's UNMODIFIED ``tools/evaluate.py`` and one epoch of ``tools/train_3d.py`` (self-supervised
configuration) run against this backend on the B200 through ``integration/run_tool.py`` -- under the tools' own
``nn.DataParallel(model, device_ids=[0]).cuda()`` wrapping (``meta`` arrives on the device), with the harness stand-ins
for the missing pip packages and the synthetic ``panoptic_synth*`` datasets (see tests/test_tools_cpu.py)."""
import os
import subprocess

import pytest

import tools_harness as H

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900),
              pytest.mark.skipif(not H.have_reference(), reason="oracle/_ref not staged (sh oracle/make_ref.sh)")]


def _run(cmd, cwd):
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, timeout=800)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    return res.stdout + res.stderr


def test_evaluate_py_runs_unchanged_on_the_gpu(tmp_path):
    y = H.write_yaml(str(tmp_path / "synth_ssv.yaml"), str(tmp_path / "out"), ssl=True)
    H.write_checkpoint(str(tmp_path / "ckpt.pth"), y)
    out = _run(H.command("evaluate.py", "--cfg", str(tmp_path / "synth_ssv.yaml"), "--with-ssv", "--test-file",
                         str(tmp_path / "ckpt.pth")), str(tmp_path))
    assert "=> load models state" in out and "Type: pose" in out and "Type: root" in out and "MPJPE" in out


@pytest.mark.parametrize("with_attn", [False, True])
def test_train_3d_py_runs_one_epoch_unchanged_on_the_gpu(tmp_path, with_attn):
    H.write_yaml(str(tmp_path / "synth_ssv.yaml"), str(tmp_path / "out"), ssl=True, with_attn=with_attn)
    out = _run(H.command("train_3d.py", "--cfg", str(tmp_path / "synth_ssv.yaml")), str(tmp_path))
    assert "Epoch: [0][1/2]" in out and "loss_pose3d_ssv" in out
    assert "Test: [1/2]" in out and "mpjpe@500mm" in out
    assert os.path.isfile(str(tmp_path / "out" / "synth_ssv" / "final_state.pth.tar"))
