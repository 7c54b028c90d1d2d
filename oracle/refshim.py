"""Import the UNMODIFIED reference (read-only tree at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so this module is used in
exactly two places: `oracle/make_golden.py` (writes tests/golden/*.npz) and the CPU-side tests that
validate `oracle/rlt_oracle.py` against the live reference when the tree happens to be present.

Shims (nothing under /root/reference is touched):
  * `numpy.lib.financial` was removed from numpy; reference utils/metrics.py:3 imports `irr` from it
    (never used) -> a stub module provides the name.
  * the reference packages are called `models` / `utils` / `dataloader`, the same names our drop-in
    packages use, so they are loaded under private aliases via importlib.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("RLT_REFERENCE_ROOT", "/root/reference"))


def available() -> bool:
    return (REFERENCE_ROOT / "models" / "__init__.py").exists()


def _load_package(alias: str, directory: Path):
    """Load `directory` (a package) under the module name `alias` without touching sys.path order."""
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, directory / "__init__.py",
                                                  submodule_search_locations=[str(directory)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns (ref_models, ref_losses, ref_metrics) — the reference's own modules."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if "numpy.lib.financial" not in sys.modules:
        stub = types.ModuleType("numpy.lib.financial")
        stub.irr = None
        sys.modules["numpy.lib.financial"] = stub
    ref_models = _load_package("_rlt_reference_models", REFERENCE_ROOT / "models")
    ref_utils = _load_package("_rlt_reference_utils", REFERENCE_ROOT / "utils")
    ref_losses = importlib.import_module("_rlt_reference_utils.losses")
    ref_metrics = importlib.import_module("_rlt_reference_utils.metrics")
    return ref_models, ref_losses, ref_metrics
