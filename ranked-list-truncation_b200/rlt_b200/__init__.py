"""rlt_b200: host side of the B200-native hot path (ctypes binding + autograd wrappers).

`models/` and `utils/` next to this package mirror the reference's Python API; this package holds
the plumbing they share.  Importing it does not load the CUDA library; the first op call does, and
raises if librlt_b200.so is missing (no CPU fallback).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
