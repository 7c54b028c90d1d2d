"""CPU: pins oracle/rlt_oracle.py against the reference's known-answer vector and the golden fixtures
produced by the unmodified reference (oracle/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

from helpers import MODEL_KW, TWO_TASK, build_model, check_weights, grad_errors, load_golden, output_error, probe_inputs
from oracle import rlt_oracle as O


def test_reference_known_answer_metric_vector():
    # reference utils/metrics.py:104-109 prints these two numbers
    x = np.array([[1, 0, 1], [0, 0, 1], [1, 0, 0]])
    k = np.array([1, 2, 1])
    assert O.metric_f1(x, k) == 0.5555555555555555
    assert O.metric_dcg(x, k) == 0.1230234154761809
    g = load_golden("metrics.npz")
    assert float(g["known/f1"]) == 0.5555555555555555 and float(g["known/dcg"]) == 0.1230234154761809


@pytest.mark.parametrize("L", [300, 40])
def test_metrics_bit_exact_vs_golden(L):
    g = load_golden("metrics.npz")
    y, ks = g[f"y_{L}"], g[f"k_{L}"]
    assert np.array_equal(O.cut_positions(g[f"p_{L}"])[2:], ks[2:])
    assert np.array_equal(np.array(O.f1_per_list(y, ks), dtype=np.float64), g[f"f1_{L}"])
    assert np.array_equal(np.array(O.dcg_per_list(y, ks), dtype=np.float64), g[f"dcg_{L}"])
    assert O.metric_f1(y, ks) == float(g[f"f1_mean_{L}"])
    assert O.metric_dcg(y, ks) == float(g[f"dcg_mean_{L}"])
    all_k = np.arange(1, L + 1)
    assert np.array_equal(np.array([O.metric_dcg(y[2:3], all_k[j:j + 1]) for j in range(L)]), g[f"dcg_allk_{L}"])
    assert np.array_equal(np.array([O.metric_f1(y[2:3], all_k[j:j + 1]) for j in range(L)]), g[f"f1_allk_{L}"])


def test_pairwise_sum_model_matches_numpy_bit_for_bit():
    rng = np.random.default_rng(0)
    for n in list(range(1, 140)) + [255, 256, 257, 300, 511, 777, 1000]:
        a = rng.standard_normal(n)
        assert O.numpy_pairwise_sum(list(a)) == a.sum(), n
    # and on the actual DCG terms
    for k in (1, 7, 8, 9, 128, 129, 300, 1000):
        t = O.DCG_TERM64[:k] * np.where(rng.random(k) < 0.3, 1.0, -1.0)
        assert O.numpy_pairwise_sum(list(t)) == t.sum()


@pytest.mark.parametrize("L", [300, 40])
def test_reward_matrices_vs_reference(L):
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"])
    for metric in ("f1", "dcg"):
        ref = g[f"reward_{metric}_{L}"]
        vec = O.reward_matrix(y, metric).numpy()
        tol = 0 if metric == "f1" else 2e-5
        assert np.abs(vec - ref).max() <= tol * max(1.0, np.abs(ref).max()), metric
    loop = O.reward_matrix_loop(y[:2], "f1").numpy()
    assert np.array_equal(loop, g[f"reward_f1_{L}"][:2])
    loop = O.reward_matrix_loop(y[:1], "dcg").numpy()
    assert np.array_equal(loop, g[f"reward_dcg_{L}"][:1])


LOSS_FNS = {
    "choopy": lambda p, y, m: O.choopy_loss(p, y, metric=m),
    "raml": lambda p, y, m: O.attncut_loss(p, y, metric=m),
    "kl": lambda p, y, m: O.div_loss(p, y, metric=m, div_type="kl"),
    "js": lambda p, y, m: O.div_loss(p, y, metric=m, div_type="js"),
    "js_noaug": lambda p, y, m: O.div_loss(p, y, metric=m, div_type="js", augmented=False),
}


@pytest.mark.parametrize("L", [300, 40])
@pytest.mark.parametrize("metric", ["f1", "dcg"])
@pytest.mark.parametrize("kind", list(LOSS_FNS))
def test_cut_losses_vs_reference(kind, metric, L):
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"])
    z = torch.from_numpy(g[f"z_{L}"]).requires_grad_(True)
    p = torch.softmax(z, dim=1)
    p.retain_grad()
    loss = LOSS_FNS[kind](p, y, metric)
    loss.backward()
    key = f"{kind}_{metric}_{L}"
    assert abs(loss.item() - float(g[key + "/loss"])) <= 2e-5 * max(1.0, abs(float(g[key + "/loss"])))
    ref_dp = g[key + "/dp"]
    assert np.abs(p.grad.numpy() - ref_dp).max() <= 2e-4 * np.abs(ref_dp).max()
    ref_dz = g[key + "/dz"]
    assert np.abs(z.grad.numpy() - ref_dz).max() <= 2e-4 * max(np.abs(ref_dz).max(), 1e-6)


@pytest.mark.parametrize("L", [300, 40])
def test_aux_and_bicut_losses_vs_reference(L):
    g = load_golden("losses.npz")
    y = torch.from_numpy(g[f"y_{L}"])
    for tag in ("active", "inactive"):
        s = torch.from_numpy(g[f"rerank_{tag}_{L}/s"]).requires_grad_(True)
        loss = O.rerank_loss(s, y)
        assert abs(loss.item() - float(g[f"rerank_{tag}_{L}/loss"])) <= 1e-6
        if tag == "active":
            loss.backward()
            assert np.abs(s.grad.numpy() - g[f"rerank_{tag}_{L}/ds"]).max() <= 1e-9
    with pytest.raises(RuntimeError):
        O.rerank_loss(torch.zeros(2, L, 1), torch.zeros(2, L))
    u = torch.from_numpy(g[f"bicut_{L}/u"])
    for metric in ("f1", "nci"):
        o = torch.softmax(u, dim=2).requires_grad_(True)
        loss = O.bicut_loss(o, y, metric=metric)
        loss.backward()
        ref = float(g[f"bicut_{metric}_{L}/loss"])
        assert abs(loss.item() - ref) <= 1e-5 * max(1.0, abs(ref))
        assert np.abs(o.grad.numpy() - g[f"bicut_{metric}_{L}/do"]).max() <= 1e-5 * np.abs(g[f"bicut_{metric}_{L}/do"]).max()


@pytest.mark.parametrize("name,B", [(n, B) for n in MODEL_KW for B in ((5,) if n in TWO_TASK else (5, 16))])
def test_model_oracle_vs_reference(name, B):
    """Weights reproduced from the seed, oracle forward (fp32 and fp64), loss and gradients vs the reference."""
    g = load_golden(f"model_{name}_B{B}.npz")
    model = build_model(name)
    check_weights(model, g)   # same submodules in the same order => same seeded init as the reference
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    out = O.FORWARDS[name](sd, x)
    outs = O.flat_outputs(out)
    assert len(outs) == int(g["n_out"])
    for i, o in enumerate(outs):
        err, ref_max = output_error(o, g, f"out{i}")
        assert err <= 5e-5 * ref_max, (name, i)
    loss = O.criterion_for(name)(O.loss_input(out), y)
    assert abs(loss.item() - float(g["loss"])) <= 5e-5 * max(1e-2, abs(float(g["loss"])))
    loss.backward()
    named = {k: (sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])) for k in map(str, g["param_names"])}
    rel_l2, rel_max, rel_norm = grad_errors(named, g)
    assert rel_l2 <= 2e-3 and rel_max <= 1e-3, (name, rel_l2, rel_max, rel_norm)
    # float64 restatement vs the reference module run in float64
    sd64 = {k: v.detach().double() for k, v in model.state_dict().items()}
    out64 = O.FORWARDS[name](sd64, x.double())
    for i, o in enumerate(O.flat_outputs(out64)):
        err, ref_max = output_error(o, g, f"out{i}_f64")
        assert err <= 1e-9 * ref_max, (name, i)


@pytest.mark.parametrize("B", [4, 9])
def test_probe_oracle_vs_reference(B):
    """models/Probe.py:102-122 with the criteria of verify_probe.py:82-83 (BCELoss / RerankLoss per probe)."""
    import models
    g = load_golden(f"probe_B{B}.npz")
    torch.manual_seed(1234)
    model = models.Probe()
    check_weights(model, g)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    e_in, e_o = probe_inputs(B)
    leaves = [e_in.requires_grad_(True)] + [t.requires_grad_(True) for t in e_o]
    y = torch.from_numpy(g["y"])
    outs = O.probe_forward(sd, e_in, e_o)
    assert len(outs) == int(g["n_out"])
    for i, o in enumerate(outs):
        err, ref_max = output_error(o, g, f"out{i}")
        assert err <= 2e-6 * ref_max, i
    parts = [O.bce_loss(outs[0], y), O.rerank_loss(outs[1], y), O.bce_loss(outs[2], y), O.bce_loss(outs[3], y),
             O.rerank_loss(outs[4], y), O.rerank_loss(outs[5], y)]
    assert np.abs(np.array([p.item() for p in parts]) - g["losses"]).max() <= 1e-6
    sum(parts).backward()
    rel_l2, rel_max, _ = grad_errors({k: v.grad for k, v in sd.items()}, g)
    assert rel_l2 <= 1e-5 and rel_max <= 1e-5
    for tag, t in zip(("in", "o0", "o1"), leaves):
        d = t.grad.double().numpy().ravel()[g[f"dx/{tag}/idx"]] - g[f"dx/{tag}/val"]
        assert np.abs(d).max() <= 1e-5 * float(g[f"dx/{tag}/absmax"]), tag


@pytest.mark.parametrize("L", [300, 40])
def test_rank_metrics_vs_reference(L):
    """Metric.taskr_metric bit for bit (tie-free predictions), Metric.taskc_metric to 1e-12 (sklearn integrates the ROC
    curve by trapezoids; the oracle divides the exact pair count once)."""
    g = load_golden("rank_metrics.npz")
    y, p = g[f"y_{L}"], g[f"p_{L}"]
    assert np.array_equal(np.array(O.taskr_dcg_per_list(y, p), dtype=np.float64), g[f"taskr_{L}"])
    assert O.taskr_metric(y, p) == float(g[f"taskr_mean_{L}"])
    for tag, pp in (("", p), ("_tied", g[f"p_tied_{L}"])):
        auc, valid = O.auc_per_list(y, pp)
        assert np.array_equal(np.nonzero(valid)[0], g[f"auc_lists{tag}_{L}"]) and not valid[3]
        assert np.abs(auc[valid] - g[f"auc{tag}_{L}"]).max() <= 1e-12
        assert abs(O.taskc_metric(y, pp) - float(g[f"taskc_mean{tag}_{L}"])) <= 1e-12
    with pytest.raises(ZeroDivisionError):
        O.taskc_metric(y[3:4], p[3:4])
    # ties keep their list order: two equal scores, the first document is ranked first
    assert O.taskr_dcg_per_list(np.array([[1., 0.]]), np.array([[.5, .5]], dtype=np.float32)) == [1 / math.log2(2) - 1 / math.log2(3)]
    assert O.taskr_dcg_per_list(np.array([[0., 1.]]), np.array([[.5, .5]], dtype=np.float32)) == [-1 / math.log2(2) + 1 / math.log2(3)]


@pytest.mark.parametrize("name", ["choopy", "mtchoopy"])
def test_oracle_positions_mode_matches_transposed_reference(name):
    """attend='positions' of the oracle against the unmodified reference with its encoder applied to the transposed
    tensor (tests/golden/model_*_positions_B5.npz)."""
    g = load_golden(f"model_{name}_positions_B5.npz")
    model = build_model(name)
    sd = {k: v.detach().double() for k, v in model.state_dict().items()}
    x = torch.from_numpy(g["x"]).double()
    out = getattr(O, name + "_forward")(sd, x, attend="positions")
    outs = out if isinstance(out, (list, tuple)) else [out]
    for i, o in enumerate(outs):
        ref = g[f"out{i}"]
        assert np.abs(o.numpy() - ref).max() <= 2e-5 * np.abs(ref).max(), (name, i)
