"""ctypes binding of librlt_b200.so (the C ABI declared in include/rlt_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent.parent          # ranked-list-truncation_b200/
REPO_ROOT = _PKG_DIR.parent
LIB_PATH = Path(os.environ.get("RLT_B200_LIB", _PKG_DIR / "librlt_b200.so"))   # override: kernel A/B experiments only
HEADER_PATH = REPO_ROOT / "include" / "rlt_b200.h"


class RltError(RuntimeError):
    """A C-ABI call returned a negative rlt_status."""


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: build it first (python -c 'import __graft_entry__ as g; g.build()' "
                "or make -C ranked-list-truncation_b200/csrc). There is no CPU fallback.")
        _lib = ctypes.CDLL(str(LIB_PATH))
        _lib.rlt_version.restype = ctypes.c_char_p
        _lib.rlt_last_error.restype = ctypes.c_char_p
    return _lib


def declared_symbols() -> list[str]:
    """Every function name declared in include/rlt_b200.h (used by the CPU-side export test)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rlt_[a-z0-9_]+)\s*\(", text)))


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rlt_last_error().decode(errors="replace")
        raise RltError(f"{what} failed with status {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_option(key: str, value: int) -> None:
    check(load().rlt_set_option(key.encode(), int(value)), f"rlt_set_option({key})")


def get_option(key: str) -> int:
    return int(load().rlt_get_option(key.encode()))
