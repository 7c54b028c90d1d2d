"""GPU, no reference tree needed: `rlt_b200.trainer.DeviceTrainer` -- the device-resident twin of run.py's loop (HBM-resident
split, gather-collated batches, on-device cut metrics, fused Adam, one host synchronisation per epoch) -- lands on the
trajectory of the reference's UNMODIFIED run.py with the reference's own packages (tests/golden/run_py_traj.json, written
on the CPU by oracle/make_golden_run_py.py): BASELINE config 1, 199 train / 50 test synthetic robust04-shaped lists in the
reference's pickle formats, batch 63, two epochs, all six families.  The test with run.py itself in the loop is
tests/test_zzzz_run_py_gpu.py (it needs the reference checkout and is skipped where that is absent)."""
import json
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["bicut", "choopy", "attncut", "mtchoopy", "mtattncut", "mmoecut"])
def test_device_trainer_lands_on_the_run_py_trajectory(model, tmp_path):
    from rlt_b200.data import write_synthetic_pickles
    from rlt_b200.trainer import DeviceTrainer
    gold = json.loads((ROOT / "tests" / "golden" / "run_py_traj.json").read_text())[model]
    db = tmp_path / "dataset" / "robust04"
    write_synthetic_pickles(db, "drmm_tks", n_train=199, n_test=50, seq_len=300)        # what tools/run_reference.py writes
    torch.manual_seed(gold["seed"])                                                      # run_reference.py: before main()
    trainer = DeviceTrainer.from_pickles(model, db, "drmm_tks", criterion=gold["criterion"], dropout=gold["dropout"])
    got = trainer.run(gold["epochs"])
    scale = max(1e-2, max(abs(v) for v in gold["scalars"]["train/loss_step"]))
    for tag in ("train/loss_step", "train/loss_epoch", "test/loss_epoch"):
        assert len(got[tag]) == len(gold["scalars"][tag]), tag
        for a, b in zip(got[tag], gold["scalars"][tag]):
            assert abs(a - b) <= 1e-3 * scale, (model, tag, a, b)
    for tag in ("train/F1_epoch", "train/DCG_epoch", "test/F1_epoch", "test/DCG_epoch"):
        for a, b in zip(got[tag], gold["scalars"][tag]):
            assert a == pytest.approx(b, rel=2e-2, abs=2e-3), (model, tag, a, b)
    print(model, {k: [round(v, 6) for v in got[k]] for k in got if k.endswith("_epoch")})
