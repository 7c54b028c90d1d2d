"""Drop-in replacement of the reference `models` package (models/__init__.py:1-12 of the reference).

Same class names, constructor arguments, forward signatures, attribute names and state_dict keys;
the forward/backward math runs in librlt_b200.so (sm_100a).  `from models import *` in the
reference's run.py resolves here when this directory precedes the reference on sys.path.
"""
from .truncation import (AttnCut, BiCut, Choopy, MMOECut, MOECut, MtAttnCut, MtChoopy, PLECut, Probe, ProbeBase, TaskC, TaskR,
                         TowerClass, TowerCut, TowerRerank)

__all__ = ["BiCut", "Choopy", "AttnCut", "MtChoopy", "MtAttnCut", "MMOECut", "MOECut", "PLECut", "ProbeBase", "Probe",
           "TowerClass", "TowerRerank", "TowerCut", "TaskC", "TaskR"]
