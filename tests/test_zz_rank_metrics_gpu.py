"""GPU parity of Metric.taskr_metric / Metric.taskc_metric (utils/metrics.py:40-76)
against goldens from the unmodified reference and against the oracle."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import rlt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [300, 40])
def test_rank_metrics_vs_reference_golden(L):
    from utils.metrics import Metric, _rank_metrics
    g = load_golden("rank_metrics.npz")
    y, p = g[f"y_{L}"], g[f"p_{L}"]
    dcg, auc, valid = _rank_metrics(y, p, True, True)
    assert np.array_equal(dcg, g[f"taskr_{L}"])                       # bit-exact: same terms, same order of additions
    assert Metric.taskr_metric(y, p) == float(g[f"taskr_mean_{L}"])
    for tag, pp in (("", p), ("_tied", g[f"p_tied_{L}"])):
        _, auc, valid = _rank_metrics(y, pp, False, True)
        assert np.array_equal(np.nonzero(valid)[0], g[f"auc_lists{tag}_{L}"])
        assert np.abs(auc[valid.astype(bool)] - g[f"auc{tag}_{L}"]).max() <= 1e-12   # sklearn: trapezoids; here: pair count
        assert abs(Metric.taskc_metric(y, pp) - float(g[f"taskc_mean{tag}_{L}"])) <= 1e-12
    with pytest.raises(ZeroDivisionError):
        Metric.taskc_metric(y[3:4], p[3:4])
    # the verify scripts pass a torch label tensor next to numpy predictions (verify_probe.py:215); CUDA tensors work too
    assert Metric.taskr_metric(torch.from_numpy(y), torch.from_numpy(p).cuda()) == float(g[f"taskr_mean_{L}"])


@pytest.mark.parametrize("B,L", [(1, 1), (3, 2), (70, 129), (9, 1000), (4, 1024)])
def test_rank_metrics_vs_oracle_with_ties(B, L):
    """Ragged sizes up to the kernel's maximum list length, heavily tied scores (stable order), one-class lists."""
    from utils.metrics import _rank_metrics
    rng = np.random.default_rng(B * 1000 + L)
    y = (rng.random((B, L)) < 0.3).astype(np.float32)
    y[0] = 0.0
    if B > 2:
        y[2] = 1.0
    p = np.round(rng.standard_normal((B, L)), 1).astype(np.float32)
    dcg, auc, valid = _rank_metrics(y, p, True, True)
    assert np.array_equal(dcg, np.array(O.taskr_dcg_per_list(y, p), dtype=np.float64))
    ref_auc, ref_valid = O.auc_per_list(y, p)
    assert np.array_equal(valid.astype(bool), ref_valid)
    assert np.array_equal(auc[ref_valid], ref_auc[ref_valid])       # same integer pair count, one division
    with pytest.raises(Exception, match="exceeds"):
        _rank_metrics(np.zeros((1, 1025), np.float32), np.zeros((1, 1025), np.float32), True, False)
