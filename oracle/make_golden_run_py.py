"""Write tests/golden/run_py_traj.json: the trajectory of the reference's UNMODIFIED run.py (Trainer.run, run.py:113-240)
with the reference's OWN models / utils packages on CPU, driven by tools/run_reference.py (synthetic robust04-shaped
pickles, 199 train / 50 test lists, batch 63 from the conf, dropout patched to 0, torch.manual_seed(7), criterion f1).

    python -m oracle.make_golden_run_py [model ...]          (each model runs in its own process)

TEST INFRASTRUCTURE: the GPU test (tests/test_zzzz_run_py_gpu.py) drives the same run.py against this repo's drop-in
packages and compares per-step loss and per-epoch loss / F1 / DCG with this fixture."""
from __future__ import annotations

import json
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden" / "run_py_traj.json"
MODELS = ["bicut", "choopy", "attncut", "mtchoopy", "mtattncut", "mmoecut"]
EPOCHS = 2


def main():
    names = sys.argv[1:] or MODELS
    rec = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in names:
        with tempfile.TemporaryDirectory() as td:
            out = Path(td) / "t.json"
            subprocess.run([sys.executable, str(ROOT / "tools" / "run_reference.py"), "--packages", "reference",
                            "--model-name", name, "--epochs", str(EPOCHS), "--seed", "7", "--dropout", "0", "--criterion",
                            "f1", "--out", str(out)], check=True, cwd=td)
            r = json.loads(out.read_text())
        rec[name] = {"epochs": EPOCHS, "seed": 7, "dropout": 0.0, "criterion": "f1",
                     "scalars": {k: [v for _, v in vals] for k, vals in r["scalars"].items()}}
        print(name, {k: [round(x, 5) for x in v] for k, v in rec[name]["scalars"].items() if k.endswith("_epoch")})
    OUT.write_text(json.dumps(rec, indent=0))


if __name__ == "__main__":
    main()
