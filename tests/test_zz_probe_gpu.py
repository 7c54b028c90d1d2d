"""GPU parity of the probing models (models/Probe.py, Classification.py, Rerank.py)
against goldens generated from the unmodified reference (oracle/make_golden.py probe probebase).
Tolerances as in test_lstm_models_gpu.py: outputs 1e-3 of the reference maximum, loss 1e-3, gradients 2e-3 (rel-L2)."""
import numpy as np
import pytest
import torch

from helpers import build_model, check_weights, grad_errors, load_golden, output_error, probe_inputs
from oracle import rlt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [5, 16])
def test_probebase_vs_reference_golden(B):
    from utils import losses
    g = load_golden(f"model_probebase_B{B}.npz")
    model = build_model("probebase")
    check_weights(model, g)
    model = model.cuda().train()
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    out = model(x)
    assert isinstance(out, tuple) and len(out) == 3 and len(out[1]) == 2 and len(out[2]) == 3   # Probe.py:99
    outs = O.flat_outputs(out)
    assert len(outs) == int(g["n_out"])
    for i, o in enumerate(outs):
        err, ref_max = output_error(o, g, f"out{i}")
        assert err <= 1e-3 * ref_max, (i, err, ref_max)
    cut = outs[-1].detach().cpu().numpy()
    assert np.array_equal(np.argmax(cut[..., 0], 1), np.argmax(g[f"out{len(outs) - 1}"][..., 0], 1))
    torch.manual_seed(0)
    loss = losses.MtCutLoss(metric="f1", num_tasks=3).cuda()(out[-1], y)           # verify_probe.py:62, :107
    ref_loss = float(g["loss"])
    assert abs(loss.item() - ref_loss) <= 1e-3 * max(abs(ref_loss), 1e-2), (loss.item(), ref_loss)
    loss.backward()
    named = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    rel_l2, rel_max, rel_norm = grad_errors(named, g)
    assert rel_l2 <= 2e-3 and rel_max <= 1e-3, (rel_l2, rel_max, rel_norm)


@pytest.mark.parametrize("B", [4, 9])
def test_probe_vs_reference_golden(B):
    """Probe on seeded representations with the criteria of verify_probe.py:82-83, :193-211."""
    import models
    from utils import losses
    g = load_golden(f"probe_B{B}.npz")
    torch.manual_seed(1234)
    model = models.Probe()
    check_weights(model, g)
    model = model.cuda()
    e_in, e_o = probe_inputs(B)
    leaves = [e_in.cuda().requires_grad_(True)] + [t.cuda().requires_grad_(True) for t in e_o]
    y = torch.from_numpy(g["y"]).cuda()
    outs = model(leaves[0], leaves[1:])
    assert len(outs) == int(g["n_out"])
    for i, o in enumerate(outs):
        err, ref_max = output_error(o, g, f"out{i}")
        assert err <= 1e-4 * ref_max, (i, err, ref_max)
    bce, rr = torch.nn.BCELoss(), losses.RerankLoss()
    parts = [bce(outs[0].squeeze(), y), rr(outs[1].squeeze(), y), bce(outs[2].squeeze(), y), bce(outs[3].squeeze(), y),
             rr(outs[4].squeeze(), y), rr(outs[5].squeeze(), y)]
    assert np.abs(np.array([p.item() for p in parts]) - g["losses"]).max() <= 1e-4
    sum(parts).backward()
    rel_l2, rel_max, _ = grad_errors({n: p.grad for n, p in model.named_parameters()}, g)
    assert rel_l2 <= 1e-3 and rel_max <= 1e-3, (rel_l2, rel_max)
    for tag, t in zip(("in", "o0", "o1"), leaves):
        d = t.grad.double().cpu().numpy().ravel()[g[f"dx/{tag}/idx"]] - g[f"dx/{tag}/val"]
        assert np.abs(d).max() <= 1e-3 * float(g[f"dx/{tag}/absmax"]), tag


@pytest.mark.parametrize("cls,d", [("TaskC", 128), ("TaskR", 128), ("TowerClass", 256), ("TowerRerank", 256)])
def test_standalone_towers_vs_torch(cls, d):
    """The stand-alone towers (Classification.py:3-13, Rerank.py:3-13, verify_probe.py:66-71) against the torch
    submodules they hold, forward and backward."""
    import models
    torch.manual_seed(d)
    m = getattr(models, cls)(d_model=d).cuda()
    x = torch.randn(7, 300, d, device="cuda")
    w = torch.randn(7, 300, 1, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out = m(xa)
    ref = getattr(m, m.attr)(xb)
    assert out.shape == ref.shape == (7, 300, 1)
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    (out * w).sum().backward()
    got = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    (ref * w).sum().backward()
    gmax = max(p.grad.abs().max().item() for p in m.parameters())   # the softmax towers' bias gradient is 0 +- rounding
    for n, p in m.named_parameters():
        assert (got[n] - p.grad).abs().max().item() <= 1e-3 * gmax, n
    assert (xa.grad - xb.grad).abs().max().item() <= 1e-3 * xb.grad.abs().max().item()
