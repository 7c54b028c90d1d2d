"""GPU: the reference's UNMODIFIED run.py (Trainer.run, run.py:113-240: DataLoader -> model -> criterion -> backward ->
torch.optim.Adam -> np.argmax cut -> Metric.f1 / Metric.dcg -> test pass) drives this repo's drop-in `models` / `utils`
packages for two epochs on synthetic robust04-shaped pickles (BASELINE config 1) and lands on the trajectory of the
reference's own packages (tests/golden/run_py_traj.json, written by oracle/make_golden_run_py.py on the CPU).

run.py is reference source: it is not part of this repository and does not travel to the GPU box with it.  The test
looks for it under $RLT_REFERENCE_ROOT, /root/reference and oracle/_ref/reference_stage (a scratch copy placed there
for one verification call, never committed) and is skipped when none exists; profiles/r02_run_py_gpu.txt holds the log
of the verified run."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _reference_root():
    for c in (os.environ.get("RLT_REFERENCE_ROOT"), "/root/reference", str(ROOT / "oracle" / "_ref" / "reference_stage")):
        if c and (Path(c) / "run.py").exists() and (Path(c) / "dataloader" / "__init__.py").exists():
            return Path(c)
    return None


@pytest.mark.parametrize("model", ["bicut", "choopy", "attncut", "mtchoopy", "mtattncut", "mmoecut"])
def test_unmodified_run_py_drives_the_drop_in_along_the_reference_trajectory(model, tmp_path):
    ref_root = _reference_root()
    if ref_root is None:
        pytest.skip("the reference's run.py is not reachable on this machine")
    gold = json.loads((ROOT / "tests" / "golden" / "run_py_traj.json").read_text())[model]
    out = tmp_path / "traj.json"
    p = subprocess.run([sys.executable, str(ROOT / "tools" / "run_reference.py"), "--packages", "b200", "--model-name",
                        model, "--epochs", str(gold["epochs"]), "--seed", str(gold["seed"]), "--dropout", "0",
                        "--criterion", gold["criterion"], "--reference-root", str(ref_root), "--out", str(out)],
                       cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    got = json.loads(out.read_text())
    assert got["packages"] == "b200" and got["cuda"]
    report = {}
    for tag, ref in gold["scalars"].items():
        vals = [v for _, v in got["scalars"][tag]]
        assert len(vals) == len(ref), tag
        report[tag] = (vals, ref)
    print(model, json.dumps({k: v for k, v in report.items() if k.endswith("_epoch")}))
    # losses: every step and every epoch within 1e-3 (relative to the largest loss of the run)
    scale = max(1e-2, max(abs(v) for v in gold["scalars"]["train/loss_step"]))
    for tag in ("train/loss_step", "train/loss_epoch", "test/loss_epoch"):
        for a, b in zip(*report[tag]):
            assert abs(a - b) <= 1e-3 * scale, (model, tag, a, b)
    # F1 / DCG of the argmax cuts: identical cut positions give bit-identical metrics (Metric.f1 / Metric.dcg are
    # bit-exact, tests/test_heads_gpu.py); a list whose two best positions are within the 1e-3 tolerance may flip
    for tag in ("train/F1_epoch", "train/DCG_epoch", "test/F1_epoch", "test/DCG_epoch"):
        for a, b in zip(*report[tag]):
            assert a == pytest.approx(b, rel=2e-2, abs=2e-3), (model, tag, a, b)
