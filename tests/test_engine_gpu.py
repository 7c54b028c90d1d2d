"""GPU: the autograd-free Engine (G groups per call, fused logits->loss->dlogits) equals the nn.Module + criterion
path run group by group and averaged (= data-parallel training of the reference with per-replica batch S)."""
import pytest
import torch

from helpers import build_model

pytestmark = pytest.mark.gpu


def _criterion(name):
    from utils import losses
    torch.manual_seed(0)
    return {"bicut": lambda: losses.BiCutLoss(metric="f1"), "choopy": lambda: losses.ChoopyLoss(metric="f1"),
            "attncut": lambda: losses.DivLoss(metric="f1", div_type="js", augmented=True)}.get(
        name, lambda: losses.MtCutLoss(metric="f1", num_tasks=3))()


@pytest.mark.parametrize("name", ["choopy", "mtchoopy", "bicut", "attncut", "mtattncut", "mmoecut"])
def test_engine_matches_module_path(name):
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from utils.metrics import Metric
    G, S, L = 2, 6, 300
    feats = 1 if "choopy" in name else 3
    model = build_model(name).cuda().train()
    x, y = synthetic_lists(G * S, L, feats, seed=11, device="cuda")
    crit = _criterion(name).cuda()
    ref_loss = 0.0
    ref_grads = {n: torch.zeros_like(p) for n, p in model.named_parameters()}
    ks = []
    for g in range(G):
        model.zero_grad(set_to_none=True)
        out = model(x[g * S:(g + 1) * S])
        loss = crit(out, y[g * S:(g + 1) * S])
        loss.backward()
        ref_loss += loss.item() / G
        for n, p in model.named_parameters():
            if p.grad is not None:
                ref_grads[n] += p.grad / G
        last = out[-1] if isinstance(out, list) else out
        ks.append(last.detach())
    eng = Engine(model, n_groups=G, group_size=S, seq_len=L)
    loss = eng.train_step(x, y).item()
    assert abs(loss - ref_loss) <= 1e-4 * max(abs(ref_loss), 1e-2), (loss, ref_loss)
    gmax = max(v.abs().max().item() for v in ref_grads.values())
    num = den = 0.0
    for n, ref in ref_grads.items():
        d = (eng.grads[n] - ref)
        assert d.abs().max().item() <= 1e-3 * gmax, (n, d.abs().max().item(), gmax)
        num += float((d.double() ** 2).sum()); den += float((ref.double() ** 2).sum())
    assert (num / den) ** 0.5 <= 1e-3, (num / den) ** 0.5
    # inference path: same cut positions and metrics as the module outputs
    k, f1, dcg = eng.infer(x, y)
    probs = torch.cat(ks)
    import numpy as np
    from oracle import rlt_oracle as O
    if name == "bicut":
        ref_k = O.bicut_cut_positions(probs.cpu().numpy())
    else:
        ref_k = O.cut_positions(probs.cpu().numpy())
    assert [int(v) for v in k.cpu().numpy()] == [int(v) for v in ref_k]
    assert float(np.mean(f1.cpu().numpy())) == pytest.approx(float(Metric.f1(y.cpu().numpy(), ref_k)), abs=1e-12)
    assert np.array_equal(dcg.cpu().numpy(), np.array(O.dcg_per_list(y.cpu().numpy(), ref_k)))
