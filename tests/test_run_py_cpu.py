"""CPU: tools/run_reference.py drives the reference's unmodified run.py.

  * with the reference's own packages it reproduces tests/golden/run_py_traj.json (so the launcher, its stubs and the
    synthetic pickles are pinned, and the fixture is what the reference computes);
  * with THIS repo's drop-in packages, run.py's imports, `Trainer.__init__` (model / criterion dispatch of run.py:59-102,
    `optim.Adam(model.parameters(), ...)`, SummaryWriter) and the DataLoader all work, and the run stops exactly where
    it must without a GPU: at the first `self.model(X_train)` (run.py:125), with the drop-in's "no CPU path" error.
Both need the reference tree (/root/reference) and are skipped where it is absent (the GPU box)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "run.py").exists(), reason="reference tree not present")


def _launch(packages, model, tmp_path, epochs=2):
    out = tmp_path / "traj.json"
    p = subprocess.run([sys.executable, str(ROOT / "tools" / "run_reference.py"), "--packages", packages, "--model-name",
                        model, "--epochs", str(epochs), "--seed", "7", "--dropout", "0", "--criterion", "f1", "--out",
                        str(out)], cwd=tmp_path, capture_output=True, text=True)
    return p, out


def test_launcher_with_reference_packages_reproduces_the_fixture(tmp_path):
    gold = json.loads((ROOT / "tests" / "golden" / "run_py_traj.json").read_text())["bicut"]
    p, out = _launch("reference", "bicut", tmp_path)
    assert p.returncode == 0, p.stderr[-2000:]
    got = json.loads(out.read_text())
    assert got["packages"] == "reference"
    for tag, ref in gold["scalars"].items():
        vals = [v for _, v in got["scalars"][tag]]
        assert len(vals) == len(ref), tag
        for a, b in zip(vals, ref):
            assert a == pytest.approx(b, rel=1e-5, abs=1e-7), tag


@pytest.mark.parametrize("model", ["bicut", "mmoecut", "mtple"])
def test_run_py_constructs_the_drop_in_and_stops_at_the_first_forward_without_a_gpu(model, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_zzzz_run_py_gpu.py runs the whole trajectory")
    p, _ = _launch("b200", model, tmp_path, epochs=1)
    assert p.returncode != 0
    assert "rlt_b200 has no CPU path" in p.stderr, p.stderr[-2000:]
    assert "Training for epoch_0" in p.stderr          # tqdm banner of run.py:120: the loader and the Trainer were built
    assert "output = self.model(X_train)" in p.stderr  # ... and the failure is run.py:125
