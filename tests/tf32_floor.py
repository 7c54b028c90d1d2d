"""Emulated-TF32 error floors of the gradient parity metrics (tests/helpers.py::grad_errors), computed on the CPU by
`oracle/tf32_emulation.py` (every matmul of the oracle with both operands rounded to TF32, float64 accumulation, forward
and backward) against the reference goldens.  A fixture whose floor exceeds the 1e-3 contract of SURVEY.md section 8(c)
cannot meet it with ANY TF32 tensor-core implementation; its GPU test asserts `max(1e-3, HEADROOM * floor)` instead.

Regenerate a row:  python tests/tf32_floor.py <fixture>     (tests/test_tf32_floor_cpu.py re-derives the pinned row)
"""
from __future__ import annotations

import sys
from pathlib import Path

# fixture -> (global rel-L2, max|d| / max|g_ref|) of the emulated-TF32 oracle gradients vs the golden (fp32 reference)
FLOOR = {
    # MtAttnCut num_tasks = 2.2 (rerank + cut): the total loss is 5.6e-3 and its gradient is dominated by the rerank
    # hinge, whose d(loss)/d(score) = +-0.5/n is the same on every token -- operand rounding errors add coherently
    # over B*L tokens.  Measured on B200 (round 2, tools/diag_two_task.py): rel_l2 1.05e-3, rel_max 1.20e-3.
    "mtattncut_t22_B5": (1.111e-03, 1.609e-03),
    "mtattncut_B5": (5.554e-04, 7.236e-04),
}
HEADROOM = 1.25


def bounds(fixture: str, l2: float = 2e-3, mx: float = 1e-3):
    """(rel_l2 bound, rel_max bound) of a model fixture: the contract, or HEADROOM x the emulated floor if that is larger."""
    f = FLOOR.get(fixture)
    if f is None:
        return l2, mx
    return max(l2, HEADROOM * f[0]), max(mx, HEADROOM * f[1])


def emulated_floor(fixture: str):
    import torch
    ROOT = Path(__file__).resolve().parent.parent
    for p in (str(ROOT), str(ROOT / "ranked-list-truncation_b200"), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from helpers import MODEL_KW, build_model, grad_errors, load_golden
    from oracle import rlt_oracle as O
    from oracle.tf32_emulation import tf32_matmuls
    name = fixture.rsplit("_B", 1)[0]
    g = load_golden(f"model_{fixture}.npz")
    model = build_model(name)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
    x, y = torch.from_numpy(g["x"]).double(), torch.from_numpy(g["y"]).double()
    fwd = getattr(O, name.split("_")[0] + "_forward")
    kw = {"num_tasks": MODEL_KW[name][1]["num_tasks"]} if "num_tasks" in MODEL_KW[name][1] else {}
    with tf32_matmuls():
        loss = O.criterion_for(name)(O.loss_input(fwd(sd, x, **kw)), y)
        loss.backward()
    named = {n: (sd[n].grad if sd[n].grad is not None else torch.zeros_like(sd[n])) for n, _ in model.named_parameters()}
    rel_l2, rel_max, _ = grad_errors(named, g)
    return rel_l2, rel_max, float(loss.item())


if __name__ == "__main__":
    for fx in sys.argv[1:]:
        print(fx, "rel_l2 %.3e rel_max %.3e loss %.8f" % emulated_floor(fx))
