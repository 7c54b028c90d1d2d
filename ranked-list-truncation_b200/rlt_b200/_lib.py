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
        bind_prototypes(_lib)
    return _lib


_SCALARS = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "uint32_t": ctypes.c_uint32, "int64_t": ctypes.c_int64,
            "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t, "float": ctypes.c_float, "double": ctypes.c_double,
            "long long": ctypes.c_longlong, "unsigned long long": ctypes.c_ulonglong,
            "rlt_stream_t": ctypes.c_void_p}      # typedef void* rlt_stream_t


def _ctype(decl: str):
    """ctypes type of one C parameter / return declaration of include/rlt_b200.h (names and `const` dropped)."""
    decl = re.sub(r"\bconst\b", " ", decl).strip()
    if "*" in decl:
        return ctypes.c_char_p if re.match(r"char\s*\*", decl) else ctypes.c_void_p
    words = decl.split()
    for n in (len(words), len(words) - 1):          # with or without a trailing parameter name
        t = " ".join(words[:n])
        if t in _SCALARS:
            return _SCALARS[t]
    if decl == "void":
        return None
    raise TypeError(f"include/rlt_b200.h: no ctypes mapping for '{decl}'")


def prototypes() -> dict:
    """{function: (restype, [argtypes])} parsed from include/rlt_b200.h -- the header is the single source of truth for
    the binding, so a call whose arguments do not convert to the declared types raises ctypes.ArgumentError instead of
    passing garbage to the GPU."""
    text = re.sub(r"/\*.*?\*/", "", HEADER_PATH.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for ret, name, params in re.findall(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(rlt_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        params = params.strip()
        args = [] if params in ("", "void") else [_ctype(a) for a in params.split(",")]
        out[name] = (_ctype(ret.strip()), args)
    return out


def bind_prototypes(lib) -> None:
    for name, (restype, argtypes) in prototypes().items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise ImportError(f"{LIB_PATH} does not export {name} (declared in include/rlt_b200.h): rebuild the library")
        fn.restype = restype
        fn.argtypes = argtypes


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
