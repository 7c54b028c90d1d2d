"""Multi-step parity (run.py:104,120-129: model -> criterion -> backward -> Adam step, batch after batch).

Goldens: tests/golden/traj_<family>.npz -- five steps of the UNMODIFIED reference with torch.optim.Adam(lr=1e-3,
weight_decay=1e-3) on five different batches of 16 lists (`python -m oracle.make_golden traj`).

What is asserted
  * module + FusedAdam and Engine + FusedAdam: the loss of every step within 2e-3 of the reference trajectory.  The
    loss of step k depends on the parameters after k - 1 optimizer steps, so operands that fail to follow the parameters
    (round 1's Engine kept private copies of the head and gate weights) show up from step 2 on;
  * the total parameter displacement after five steps against the reference's.  Adam normalises every element's update
    to ~lr * sign(g), so an element whose gradient is at rounding-noise level moves by +-lr whichever way its noise
    points: the comparison is the cosine between our displacement and the reference's over all parameters with the
    softmax-shift-invariant ones (zero gradient by construction) left out, plus a direct check that the head / tower /
    gate parameters moved like the reference's;
  * Engine + FusedAdam against module + FusedAdam (same kernels, different orchestration): losses within 1e-4.
"""
import numpy as np
import pytest
import torch

from helpers import build_model, load_golden

pytestmark = pytest.mark.gpu

FAMILIES = ["bicut", "choopy", "attncut", "mtchoopy", "mtattncut", "mmoecut"]


def _criterion(name):
    from utils import losses
    torch.manual_seed(0)
    if name == "bicut":
        return losses.BiCutLoss(metric="f1")
    if name == "choopy":
        return losses.ChoopyLoss(metric="f1")
    if name == "attncut":
        return losses.DivLoss(metric="f1", div_type="js", augmented=True)
    if name == "mmoecut":
        return losses.MtCutLoss(metric="f1", num_tasks=3)
    return losses.MtCutLoss(metric="f1", rerank_weight=0.5, classi_weight=0.5, num_tasks=3)


def _run_module(name, g):
    from rlt_b200.optim import FusedAdam
    model = build_model(name).cuda().train()
    init = {n: p.detach().clone() for n, p in model.named_parameters()}
    crit = _criterion(name).cuda()
    opt = FusedAdam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    losses = []
    for step in range(int(g["steps"])):
        x, y = torch.from_numpy(g[f"x{step}"]).cuda(), torch.from_numpy(g[f"y{step}"]).cuda()
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return model, init, losses


def _run_engine(name, g):
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    model = build_model(name).cuda().train()
    init = {n: p.detach().clone() for n, p in model.named_parameters()}
    eng = Engine(model, n_groups=1, group_size=g["x0"].shape[0], seq_len=300)
    opt = FusedAdam.for_engine(eng, lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    losses = []
    for step in range(int(g["steps"])):
        x, y = torch.from_numpy(g[f"x{step}"]).cuda(), torch.from_numpy(g[f"y{step}"]).cuda()
        losses.append(eng.train_step(x, y).item())
        opt.step()
    return model, init, losses


def _displacement_cosine(model, init, g, subset=None):
    """cos(our total parameter displacement, the reference's) over the sampled entries stored in the golden."""
    dot = na = nb = 0.0
    for n, p in model.named_parameters():
        if subset is not None and not any(s in n for s in subset):
            continue
        if subset is None and (n.endswith(("decison_layer.0.bias", "cut_layer.0.bias")) or ".norm2.bias" in n):
            continue          # zero-gradient parameters (softmax shift invariance): their Adam step is sign(noise)
        idx = g[f"delta/{n}/idx"]
        ours = (p.detach() - init[n]).double().cpu().numpy().ravel()[idx]
        ref = g[f"delta/{n}/val"]
        dot += float((ours * ref).sum()); na += float((ours * ours).sum()); nb += float((ref * ref).sum())
    return dot / max((na * nb) ** 0.5, 1e-300)


HEADS = ("decison_layer.0.weight", "classi.0.weight", "rerank.weight", "towers.", "w_gates", "softmax.1.weight")


@pytest.mark.parametrize("runner", ["module", "engine"])
@pytest.mark.parametrize("name", FAMILIES)
def test_five_adam_steps_follow_the_reference_trajectory(name, runner):
    g = load_golden(f"traj_{name}.npz")
    model, init, losses = (_run_module if runner == "module" else _run_engine)(name, g)
    ref = g["losses"]
    scale = max(1e-2, float(np.abs(ref).max()))
    for k, (a, b) in enumerate(zip(losses, ref)):
        assert abs(a - b) <= 2e-3 * scale, (name, runner, k, losses, ref.tolist())
    cos_all = _displacement_cosine(model, init, g)
    cos_heads = _displacement_cosine(model, init, g, HEADS)
    assert cos_all >= 0.98, (name, runner, cos_all)
    assert cos_heads >= 0.98, (name, runner, cos_heads)


@pytest.mark.parametrize("name", FAMILIES)
def test_engine_trajectory_equals_module_trajectory(name):
    g = load_golden(f"traj_{name}.npz")
    ma, ia, la = _run_module(name, g)
    mb, ib, lb = _run_engine(name, g)
    scale = max(1e-2, max(abs(v) for v in la))
    for k, (a, b) in enumerate(zip(la, lb)):
        assert abs(a - b) <= 1e-4 * scale, (name, k, la, lb)
    # the stacked operands of the Engine (head weights, MMOECut gates) must have followed the optimizer
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        if any(s in n for s in HEADS):
            da, db = (pa - ia[n]).flatten().double(), (pb - ib[n]).flatten().double()
            cos = float((da * db).sum() / (da.norm() * db.norm() + 1e-300))
            assert cos >= 0.99, (n, cos)
