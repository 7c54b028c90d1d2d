#!/usr/bin/env python
"""Drive the reference's UNMODIFIED run.py (Trainer.run: run.py:113-240) against either the reference's own
`models` / `utils` packages or this repo's drop-in packages, on synthetic robust04-shaped pickles (BASELINE config 1:
249 queries x 300 documents -> 199 train / 50 test, batch 63).

    python tools/run_reference.py --packages reference --model-name bicut --epochs 2 --out traj.json    # CPU, here
    python tools/run_reference.py --packages b200      --model-name bicut --epochs 2 --out traj.json    # B200 box

What is done around run.py (nothing inside it is edited; the file is executed from where it lies):
  * `tensorboardX` and `matplotlib` are not installed here: stub modules are registered before the import; the stub
    SummaryWriter records every `add_scalar`, which is how the trajectory (loss / F1 / DCG per step and epoch) is read;
  * `numpy.lib.financial` (dead import of utils/metrics.py:3, removed from numpy) gets a stub when the reference's
    `utils` is used;
  * run.py:22 `RUNNING_PATH` and `dataloader/*_dataloader.py:10` `DATASET_BASE` are hard-coded paths of the author's
    machine: the two module globals are re-pointed at a work directory that holds the reference's two `.conf` files
    (dropout optionally patched to 0 so that the run is deterministic) and the synthetic pickles
    (`rlt_b200.data.write_synthetic_pickles`, formats of SURVEY.md section 8(c));
  * `sys.path` decides which `models` / `utils` run.py's `from models import *` / `from utils import losses` resolve to:
    the reference tree (`--packages reference`) or `ranked-list-truncation_b200/` first (`--packages b200`); the
    `dataloader` package is the reference's in both cases (L4 stays as it is);
  * `torch.manual_seed(--seed)` before `main()`: both package sets build the same torch submodules in the same order,
    so the initial weights, the DataLoader shuffles and MtCutLoss's unused random Parameter are identical.

run.py is reference source and does not travel to the GPU box with the repository: `--run-py` / `--reference-root`
name where it lies (default /root/reference).  tests/test_run_py_cpu.py runs the `reference` side here and pins the
trajectory fixture; tests/test_zzzz_run_py_gpu.py runs the `b200` side wherever a run.py is reachable.
"""
from __future__ import annotations

import argparse
import configparser
import importlib
import importlib.util
import json
import os
import sys
import tempfile
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ranked-list-truncation_b200"


class _Recorder:
    """Stand-in for tensorboardX.SummaryWriter (run.py:111,146,154-156,196-198)."""
    scalars: dict = {}

    def __init__(self, *a, **k):
        pass

    def add_scalar(self, tag, value, step=None):
        _Recorder.scalars.setdefault(tag, []).append((None if step is None else int(step), float(value)))

    def close(self):
        pass


def _install_stubs(reference_utils: bool):
    tb = types.ModuleType("tensorboardX")
    tb.SummaryWriter = _Recorder
    sys.modules["tensorboardX"] = tb
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "plot", "title", "xlabel", "ylabel", "legend", "savefig", "close", "clf", "subplot", "bar"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    if reference_utils and "numpy.lib.financial" not in sys.modules:
        fin = types.ModuleType("numpy.lib.financial")
        fin.irr = None
        sys.modules["numpy.lib.financial"] = fin


def prepare_workdir(work: Path, reference_root: Path, dataset_name: str, dropout, lr, n_train: int, n_test: int,
                    seq_len: int = 300):
    """Conf files (from the reference tree, values optionally overridden in every section) + synthetic pickles."""
    sys.path.insert(0, str(PKG))
    from rlt_b200.data import write_synthetic_pickles
    work.mkdir(parents=True, exist_ok=True)
    for conf in sorted(reference_root.glob("hyper_parameter_*.conf")):
        cp = configparser.ConfigParser()
        cp.read(conf)
        for section in cp.sections():
            if dropout is not None:
                cp.set(section, "dropout", str(dropout))
            if lr is not None:
                cp.set(section, "lr", str(lr))
        with open(work / conf.name, "w") as f:
            cp.write(f)
    write_synthetic_pickles(work / "dataset" / "robust04", dataset_name, n_train=n_train, n_test=n_test, seq_len=seq_len)


def run(packages: str, model_name: str, epochs: int, reference_root: Path, run_py: Path | None = None, seed: int = 7,
        dropout=0.0, lr=None, dataset_name: str = "drmm_tks", criterion: str = "f1", num_tasks: float = 3,
        n_train: int = 199, n_test: int = 50, work: Path | None = None, extra_args=()) -> dict:
    """Execute run.py's main() once; returns {'scalars': {tag: [(step, value), ...]}, 'packages': ..., ...}."""
    import torch
    reference_root = Path(reference_root)
    run_py = Path(run_py) if run_py else reference_root / "run.py"
    if not run_py.exists():
        raise FileNotFoundError(f"{run_py}: the reference's run.py is not reachable from here")
    work = Path(work) if work else Path(tempfile.mkdtemp(prefix="rlt_run_py_"))
    prepare_workdir(work, reference_root, dataset_name, dropout, lr, n_train, n_test)
    _install_stubs(reference_utils=(packages == "reference"))
    sys.dont_write_bytecode = True                       # the reference tree is read-only
    # import order decides the package set: drop any earlier `models` / `utils` / `dataloader` / `run`
    for name in [m for m in sys.modules if m.split(".")[0] in ("models", "utils", "dataloader", "run")]:
        del sys.modules[name]
    paths = [str(reference_root)] if packages == "reference" else [str(PKG), str(reference_root)]
    saved_path = list(sys.path)
    sys.path[:0] = paths
    try:
        dl = importlib.import_module("dataloader")
        for sub in ("bicut_dataloader", "choopy_dataloader", "attncut_dataloader", "mtcut_dataloader"):
            mod = sys.modules.get(f"dataloader.{sub}")
            if mod is not None and hasattr(mod, "DATASET_BASE"):
                mod.DATASET_BASE = str(work / "dataset")
        spec = importlib.util.spec_from_file_location("run", run_py)
        runmod = importlib.util.module_from_spec(spec)
        sys.modules["run"] = runmod
        spec.loader.exec_module(runmod)                  # executes run.py's top level (imports, class Trainer, main)
        origin = Path(sys.modules["models"].__file__).resolve().parent
        expect = (reference_root / "models").resolve() if packages == "reference" else (PKG / "models").resolve()
        if origin != expect:
            raise RuntimeError(f"run.py resolved `models` to {origin}, expected {expect}")
        runmod.RUNNING_PATH = str(work)
        _Recorder.scalars = {}
        argv = ["run.py", "--model-name", model_name, "--dataset-name", dataset_name, "--epochs", str(epochs),
                "--criterion", criterion, "--num-tasks", str(num_tasks), "--save-path", str(work / "best_model") + "/",
                *extra_args]
        saved_argv = sys.argv
        sys.argv = argv
        torch.manual_seed(seed)
        try:
            runmod.main()
        finally:
            sys.argv = saved_argv
        del dl
    finally:
        sys.path[:] = saved_path
    return {"packages": packages, "model_name": model_name, "epochs": epochs, "seed": seed, "dropout": dropout,
            "cuda": bool(torch.cuda.is_available()), "scalars": {k: v for k, v in _Recorder.scalars.items()}}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--packages", choices=["reference", "b200"], default="b200")
    ap.add_argument("--model-name", default="bicut")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--dropout", type=float, default=0.0, help="written into every conf section (negative: keep the conf's)")
    ap.add_argument("--lr", type=float, default=None)
    ap.add_argument("--criterion", default="f1")
    ap.add_argument("--num-tasks", type=float, default=3)
    ap.add_argument("--reference-root", default=os.environ.get("RLT_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--run-py", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rec = run(a.packages, a.model_name, a.epochs, Path(a.reference_root), a.run_py, a.seed,
              None if a.dropout < 0 else a.dropout, a.lr, criterion=a.criterion, num_tasks=a.num_tasks)
    text = json.dumps(rec)
    if a.out:
        Path(a.out).write_text(text)
    for tag in ("train/loss_epoch", "train/F1_epoch", "train/DCG_epoch", "test/loss_epoch", "test/F1_epoch", "test/DCG_epoch"):
        print(tag, [round(v, 6) for _, v in rec["scalars"].get(tag, [])])


if __name__ == "__main__":
    main()
