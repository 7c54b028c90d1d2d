"""Shared test helpers: golden loading, seeded weights, error metrics."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
WEIGHT_SEED = 1234

MODEL_KW = {
    "bicut": ("BiCut", dict(input_size=3, dropout=0.0)),
    "choopy": ("Choopy", dict(seq_len=300, dropout=0.0)),
    "attncut": ("AttnCut", dict(input_size=3, dropout=0.0)),
    "mtchoopy": ("MtChoopy", dict(seq_len=300, num_tasks=3, dropout=0.0)),
    "mtattncut": ("MtAttnCut", dict(input_size=3, num_tasks=3, dropout=0.0)),
    "mmoecut": ("MMOECut", dict(seq_len=300, num_tasks=3, input_size=3, dropout=0.0, num_experts=3)),
    "moecut": ("MOECut", dict(seq_len=300, num_tasks=3, input_size=3, dropout=0.0)),
    "plecut": ("PLECut", dict(seq_len=300, input_size=3, dropout=0.0, num_experts=3)),
    "probebase": ("ProbeBase", dict(seq_len=300, num_tasks=3, input_size=3, dropout=0.0, num_experts=2)),
    # run.py --num_tasks 2.1 (class + cut) / 2.2 (rerank + cut): goldens at B = 5 only
    "mtchoopy_t21": ("MtChoopy", dict(seq_len=300, num_tasks=2.1, dropout=0.0)),
    "mtchoopy_t22": ("MtChoopy", dict(seq_len=300, num_tasks=2.2, dropout=0.0)),
    "mtattncut_t21": ("MtAttnCut", dict(input_size=3, num_tasks=2.1, dropout=0.0)),
    "mtattncut_t22": ("MtAttnCut", dict(input_size=3, num_tasks=2.2, dropout=0.0)),
    "mmoecut_t21": ("MMOECut", dict(seq_len=300, num_tasks=2.1, input_size=3, dropout=0.0, num_experts=3)),
    "mmoecut_t22": ("MMOECut", dict(seq_len=300, num_tasks=2.2, input_size=3, dropout=0.0, num_experts=3)),
}
TWO_TASK = [n for n in MODEL_KW if n.endswith(("_t21", "_t22"))]
PROBE_SEED = 4321


def probe_inputs(B: int, L: int = 300, d: int = 256):
    """The seeded (experts_in, experts_o) of tests/golden/probe_B*.npz (oracle/make_golden.py::probe_inputs)."""
    g = torch.Generator().manual_seed(PROBE_SEED + B)
    return torch.randn(B, L, d, generator=g), [torch.randn(B, L, d, generator=g), torch.randn(B, L, d, generator=g)]


def output_error(o, g, key: str):
    """(max |o - ref|, max |ref|) of one model output against the golden entry `key`, stored either whole or as a
    shape + sampled digest (oracle/make_golden.py::store_output).  Asserts the shape."""
    a = o.detach().double().cpu().numpy()
    if key in g.files:
        ref = g[key]
        assert a.shape == ref.shape, (key, a.shape, ref.shape)
        return float(np.abs(a - ref).max()), float(np.abs(ref).max())
    assert a.shape == tuple(int(v) for v in g[key + "/shape"]), (key, a.shape)
    assert a.size == int(g[key + "/size"])
    return float(np.abs(a.ravel()[g[key + "/idx"]] - g[key + "/val"]).max()), float(g[key + "/absmax"])



def load_golden(name: str):
    return np.load(GOLDEN / name, allow_pickle=False)


def build_model(name: str):
    """Our drop-in module with the golden's weights (seeded init = the reference's init)."""
    import models
    cls, kw = MODEL_KW[name]
    torch.manual_seed(WEIGHT_SEED)
    return getattr(models, cls)(**kw)


def check_weights(model, g):
    for pname, p in model.named_parameters():
        assert abs(p.detach().double().sum().item() - float(g[f"wsum/{pname}"])) <= 1e-9 * max(1.0, float(g[f"wabs/{pname}"])), pname
        assert abs(p.detach().double().abs().sum().item() - float(g[f"wabs/{pname}"])) <= 1e-9 * max(1.0, float(g[f"wabs/{pname}"])), pname


def grad_errors(named_grads: dict, g):
    """Compares gradients with the golden digests.  Returns (global rel-L2 over the sampled entries,
    max|delta| / global max|g_ref|, worst per-tensor l2-norm relative error)."""
    num = den = 0.0
    gmax = max(float(g[f"grad/{n}/absmax"]) for n in g["param_names"])
    worst_abs = worst_norm = 0.0
    for n in g["param_names"]:
        n = str(n)
        ref_idx, ref_val = g[f"grad/{n}/idx"], g[f"grad/{n}/val"]
        got = named_grads[n].detach().double().cpu().numpy().ravel()
        assert got.size == int(g[f"grad/{n}/size"]), n
        d = got[ref_idx] - ref_val
        num += float((d * d).sum())
        den += float((ref_val * ref_val).sum())
        worst_abs = max(worst_abs, float(np.abs(d).max()))
        l2 = float(np.sqrt((got * got).sum()))
        ref_l2 = float(g[f"grad/{n}/l2"])
        if ref_l2 > 1e-3 * gmax:
            worst_norm = max(worst_norm, abs(l2 - ref_l2) / ref_l2)
    return (num / max(den, 1e-300)) ** 0.5, worst_abs / gmax, worst_norm
