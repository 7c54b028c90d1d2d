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


@pytest.mark.parametrize("name", ["choopy", "bicut", "attncut"])
def test_cuda_graph_replay_matches_eager_step(name):
    """Row N2: the captured train step (one graph launch) reproduces the eager step's loss and gradient bucket on new
    data copied into the static buffers (bit-exact up to the order of the weight-gradient atomics)."""
    import models
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    feats = {"choopy": 1, "bicut": 3, "attncut": 3}[name]
    kw = {"choopy": dict(seq_len=300, dropout=0.0), "bicut": dict(input_size=3, dropout=0.0),
          "attncut": dict(input_size=3, dropout=0.0)}[name]
    torch.manual_seed(1234)
    model = getattr(models, {"choopy": "Choopy", "bicut": "BiCut", "attncut": "AttnCut"}[name])(**kw).cuda()
    eng = Engine(model, n_groups=1, group_size=63, seq_len=300, training=True)
    xs, ys = synthetic_lists(63, 300, feats, seed=5, device="cuda")
    replay = eng.capture_train_step(xs, ys)
    x2, y2 = synthetic_lists(63, 300, feats, seed=6, device="cuda")
    eng.train_step(x2, y2)
    ref_loss, ref_bucket = eng.loss.item(), eng.grad_bucket.clone()
    xs.copy_(x2)
    ys.copy_(y2)
    eng.grad_bucket.fill_(123.0)          # the replay must zero and refill the bucket itself
    replay()
    torch.cuda.synchronize()
    assert abs(eng.loss.item() - ref_loss) <= 1e-6 * max(1.0, abs(ref_loss))
    d = (eng.grad_bucket - ref_bucket).abs().max().item()
    assert d <= 1e-5 * ref_bucket.abs().max().item(), d


@pytest.mark.parametrize("name", ["choopy", "attncut", "mmoecut"])
def test_whole_step_graph_equals_eager_steps(name):
    """forward + criterion + backward + cut metrics + fused Adam as ONE CUDA graph: five replays on five batches leave the
    same parameters, moments, losses and per-list metrics as five eager Engine steps.  Not bitwise: the weight-gradient
    kernels accumulate with atomics, two EAGER runs already differ by ~1e-6 relative from the second step on
    (tools/diag_graph.py); the first loss (no update yet) is bitwise."""
    import copy
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    from rlt_b200.optim import FusedAdam
    from rlt_b200 import ops
    feats = 1 if name == "choopy" else 3
    S, L = 63, 300
    model_a = build_model(name).cuda().train()
    model_b = copy.deepcopy(model_a)
    start = [p.detach().clone() for p in model_a.parameters()]
    batches = [synthetic_lists(S, L, feats, seed=40 + i, device="cuda") for i in range(5)]
    # eager
    eng_a = Engine(model_a, n_groups=1, group_size=S, seq_len=L)
    opt_a = FusedAdam.for_engine(eng_a, lr=1e-3, weight_decay=1e-3)
    ref = []
    for x, y in batches:
        loss = eng_a.train_step(x, y).item()
        m = ops.eval_cut(eng_a.z[eng_a.H - 1], y, mode=0)
        opt_a.step()
        ref.append((loss, m[0].clone(), m[3].clone(), m[4].clone()))
    # one graph
    eng_b = Engine(model_b, n_groups=1, group_size=S, seq_len=L)
    opt_b = FusedAdam.for_engine(eng_b, lr=1e-3, weight_decay=1e-3)
    xs, ys = batches[0][0].clone(), batches[0][1].clone()
    replay = eng_b.capture_train_step(xs, ys, optimizer=opt_b, metrics=True)
    assert opt_b.step_count == 0 and int(opt_b.tensor_steps.max()) == 0          # capture and warm-up did not train
    for (x, y), (loss, k, f1, dcg) in zip(batches, ref):
        xs.copy_(x)
        ys.copy_(y)
        replay()
        gk, _, _, gf1, gdcg = eng_b.graph_metrics
        if loss is ref[0][0]:
            assert eng_b.loss.item() == loss
        assert abs(eng_b.loss.item() - loss) <= 2e-5 * max(abs(loss), 1e-2), (eng_b.loss.item(), loss)
        same = (gk == k)
        assert int(same.sum()) >= (3 * S) // 4                            # near-tied arg-maxima of an untrained model move with O(lr) weight noise
        assert torch.equal(gf1[same], f1[same]) and torch.equal(gdcg[same], dcg[same])
    assert opt_b.step_count == 5
    assert torch.equal(opt_a.tensor_steps, opt_b.tensor_steps)
    # Adam divides by sqrt(v): where a gradient is ~0 its rounding noise is amplified to O(lr), so single weights are
    # compared loosely and the displacement of all parameters by its direction (as tests/test_zzzz_trajectory_gpu.py)
    num = na = nb = 0.0
    for (n, pa), (_, pb), p0 in zip(model_a.named_parameters(), model_b.named_parameters(), start):
        da, db = (pa - p0).double(), (pb - p0).double()
        num += float((da * db).sum()); na += float((da * da).sum()); nb += float((db * db).sum())
        assert (pa - pb).abs().max().item() <= 2e-3, (n, (pa - pb).abs().max().item())
    assert num / (na * nb) ** 0.5 >= 0.9999, num / (na * nb) ** 0.5
    assert (opt_a.exp_avg - opt_b.exp_avg).abs().max().item() <= 1e-2 * opt_a.exp_avg.abs().max().item()   # weights already differ by O(lr)
