"""Optimizer row N1: the oracle's Adam restatement against goldens produced by torch.optim.Adam (CPU, the
reference's construction at run.py:104), and the fused CUDA step against the oracle.
Tolerance: the update is a chain of ~10 fp32 operations whose fusion (FMA contraction) differs between torch's CPU
kernels, numpy and our kernel: parameters agree to 2e-7 relative to max|p| per step, moments to 1e-6 relative."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import rlt_oracle as O

GOLD = Path(__file__).resolve().parent / "golden" / "adam.npz"
N_TENSORS = 9


def _case(z, case):
    lr, wd, steps = z[f"{case}_hyper"]
    steps = int(steps)
    p0 = [z[f"{case}_p0_{i}"].copy() for i in range(N_TENSORS)]
    grads = [[z[f"{case}_g{t}_{i}"] for i in range(N_TENSORS)] for t in range(steps)]
    pT = [z[f"{case}_p{steps}_{i}"] for i in range(N_TENSORS)]
    mT = [z[f"{case}_m_{i}"] for i in range(N_TENSORS)]
    vT = [z[f"{case}_v_{i}"] for i in range(N_TENSORS)]
    return float(lr), float(wd), steps, p0, grads, pT, mT, vT


@pytest.mark.parametrize("case", ["wd", "nowd"])
def test_oracle_adam_matches_torch_golden(case):
    lr, wd, steps, p, grads, pT, mT, vT = _case(np.load(GOLD), case)
    m = [np.zeros_like(a) for a in p]
    v = [np.zeros_like(a) for a in p]
    for t in range(steps):
        O.adam_step(p, grads[t], m, v, t + 1, lr=lr, weight_decay=wd)
    for i in range(N_TENSORS):
        assert np.abs(p[i] - pT[i]).max() <= 2e-7 * steps * np.abs(pT[i]).max() + 1e-12, ("p", i)
        assert np.abs(m[i] - mT[i]).max() <= 1e-6 * np.abs(mT[i]).max() + 1e-30, ("m", i)
        assert np.abs(v[i] - vT[i]).max() <= 1e-6 * np.abs(vT[i]).max() + 1e-30, ("v", i)


def test_fused_adam_refuses_cpu_parameters_and_bad_hyperparameters():
    from rlt_b200.optim import FusedAdam
    p = [torch.nn.Parameter(torch.zeros(4))]
    with pytest.raises(RuntimeError, match="no CPU"):
        FusedAdam(p)
    with pytest.raises(ValueError):
        FusedAdam(p, lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([])
    with pytest.raises(NotImplementedError):
        FusedAdam(p, amsgrad=True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["wd", "nowd"])
def test_fused_adam_matches_golden_and_oracle(case):
    from rlt_b200.optim import FusedAdam
    lr, wd, steps, p0, grads, pT, mT, vT = _case(np.load(GOLD), case)
    params = [torch.nn.Parameter(torch.from_numpy(a).cuda()) for a in p0]
    opt = FusedAdam(params, lr=lr, weight_decay=wd)
    for t in range(steps):
        for p, g in zip(params, grads[t]):
            p.grad = torch.from_numpy(g).cuda()
        opt.step()
    for i, p in enumerate(params):
        got = p.detach().cpu().numpy()
        assert np.abs(got - pT[i]).max() <= 2e-7 * steps * np.abs(pT[i]).max() + 1e-12, ("p", i)
        m, v = opt.moments(i)
        assert np.abs(m.cpu().numpy() - mT[i]).max() <= 1e-6 * np.abs(mT[i]).max() + 1e-30, ("m", i)
        assert np.abs(v.cpu().numpy() - vT[i]).max() <= 1e-6 * np.abs(vT[i]).max() + 1e-30, ("v", i)


@pytest.mark.gpu
def test_fused_adam_steps_a_model_like_torch_adam():
    """Whole-model check on the drop-in path: Choopy forward + loss + backward, then FusedAdam vs torch.optim.Adam
    (CUDA) on the same gradients from the same state for three steps, with the reference's hyper-parameters (lr 3e-5, weight decay)."""
    import copy
    import models
    from utils import losses
    torch.manual_seed(5)
    a = models.Choopy(seq_len=40, dropout=0.0).cuda()
    b = copy.deepcopy(a)
    oa = __import__("rlt_b200.optim", fromlist=["FusedAdam"]).FusedAdam(a.parameters(), lr=3e-5, weight_decay=1e-3)
    ob = torch.optim.Adam(b.parameters(), lr=3e-5, weight_decay=1e-3)
    crit = losses.ChoopyLoss()
    for step in range(3):
        x = torch.rand(16, 40, 1, device="cuda")
        y = (torch.rand(16, 40, device="cuda") < 0.2).float()
        oa.zero_grad()
        crit(a(x), y).backward()
        # both optimizers see the SAME gradient: Adam's first steps are sign-like (m / sqrt(v) = +-1), so the ulp-level
        # run-to-run differences of the backward's atomics would otherwise flip updates of near-zero gradients
        for pa, pb in zip(a.parameters(), b.parameters()):
            pb.grad = pa.grad.detach().clone()
        oa.step()
        ob.step()
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        d = (pa - pb).abs().max().item()
        assert d <= 2e-6 * pb.abs().max().item() + 1e-9, (n, d)


@pytest.mark.gpu
def test_fused_adam_from_engine_bucket():
    """Engine path: gradients are read from the flat bucket; grad_scale reproduces the averaged all-reduce."""
    import models
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    torch.manual_seed(9)
    model = models.Choopy(seq_len=300, dropout=0.0).cuda()
    eng = Engine(model, n_groups=1, group_size=16, seq_len=300, training=True)
    x, y = synthetic_lists(16, 300, 1, seed=3, device="cuda")
    before = [p.detach().clone() for _, p in eng.named_params]
    opt = FusedAdam.for_engine(eng, lr=1e-3, weight_decay=1e-3, grad_scale=0.5)
    eng.train_step(x, y)
    grads = [eng.grads[n].detach().cpu().numpy().copy() for n, _ in eng.named_params]
    opt.step()
    p = [t.cpu().numpy().copy() for t in before]
    m = [np.zeros_like(t) for t in p]
    v = [np.zeros_like(t) for t in p]
    O.adam_step(p, grads, m, v, 1, lr=1e-3, weight_decay=1e-3, grad_scale=0.5)
    for (n, q), ref in zip(eng.named_params, p):
        d = np.abs(q.detach().cpu().numpy() - ref).max()
        assert d <= 2e-7 * np.abs(ref).max() + 1e-9, (n, d)


@pytest.mark.gpu
def test_fused_adam_skips_parameters_without_gradient_like_torch():
    """torch.optim.Adam leaves a parameter whose .grad is None untouched (no L2 decay, no moment decay) and counts its
    steps separately (state['step'] per parameter).  run.py:121 resets gradients to None every batch and the rerank
    head has none while its hinge is inactive (utils/losses.py:141), so the reference freezes it for those batches."""
    from rlt_b200.optim import FusedAdam
    torch.manual_seed(3)
    shapes = [(7, 5), (300,), (4097,), (1,)]
    pa = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedAdam(pa, lr=1e-2, weight_decay=1e-2)
    ob = torch.optim.Adam(pb, lr=1e-2, weight_decay=1e-2)
    present = [(1, 1, 1, 1), (1, 0, 1, 0), (1, 0, 0, 1), (1, 1, 1, 0), (0, 1, 1, 1)]
    for step, mask in enumerate(present):
        for i, (a, b) in enumerate(zip(pa, pb)):
            g = torch.randn_like(a) if mask[i] else None
            a.grad = g
            b.grad = None if g is None else g.clone()
        oa.step()
        ob.step()
        for i, (a, b) in enumerate(zip(pa, pb)):
            d = (a - b).abs().max().item()
            assert d <= 4e-7 * (step + 1) * b.abs().max().item() + 1e-9, (step, i, d)
    assert oa.tensor_steps.tolist() == [4, 3, 4, 3]
    assert [int(ob.state[p]["step"]) for p in pb] == [4, 3, 4, 3]
