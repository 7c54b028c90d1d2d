"""Drop-in replacement of the reference utils/metrics.py (Metric, Metric_for_Loss).

Counting, cut selection and the order-exact float64 DCG summation run on the GPU (K4,
rlt_eval_cut); only the final np.mean over the per-list values is taken on the host, exactly as the
reference does (utils/metrics.py:24,38).  `taskr_metric` / `taskc_metric` (the verify scripts' full-list DCG and
per-list AUC) run in rlt_rank_metrics: ranks by counting instead of a sort, AUC as the Mann-Whitney statistic.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from rlt_b200 import ops
from rlt_b200._lib import check, ptr, stream_ptr

# same expression as the reference table (utils/metrics.py:7); extended to 1024 entries
DCG_coef_300 = [math.log(j + 2, 2) for j in range(300)]


def _eval_given_k(labels: np.ndarray, k_s):
    """GPU evaluation of per-list (count, f1, dcg) for host labels and host cut positions."""
    if not torch.cuda.is_available():
        raise RuntimeError("utils.metrics.Metric needs a CUDA device (rlt_b200 has no CPU fallback)")
    labels = np.ascontiguousarray(labels, dtype=np.float32)
    n, L = labels.shape
    if len(k_s) != n:
        raise ValueError(f"{n} lists but {len(k_s)} cut positions")
    k_arr = np.array([int(k) for k in k_s], dtype=np.int32)
    if k_arr.min() < 1:
        raise ValueError("cut positions are counts (>= 1), not indices")
    if k_arr.max() > L:
        raise ValueError(f"operands could not be broadcast together: cut position {k_arr.max()} exceeds the list length {L}")
    pyint = np.array([0 if isinstance(k, np.integer) else 1 for k in k_s], dtype=np.int32)
    dev = torch.device("cuda", torch.cuda.current_device())
    y = torch.from_numpy(labels).to(dev)
    k_d = torch.from_numpy(k_arr).to(dev)
    py_d = torch.from_numpy(pyint).to(dev)
    ops.ensure_tables()
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    f1 = torch.empty(n, dtype=torch.float64, device=dev)
    dcg = torch.empty(n, dtype=torch.float64, device=dev)
    check(ops.lib().rlt_eval_given_k(ptr(y), ptr(k_d), ptr(py_d), n, L, ptr(cnt), None, ptr(f1), ptr(dcg), stream_ptr()),
          "rlt_eval_given_k")
    return cnt.cpu().numpy(), f1.cpu().numpy(), dcg.cpu().numpy(), pyint


def _as_f32_device(a, what: str) -> torch.Tensor:
    """labels / predictions as a contiguous float32 CUDA tensor [n, L]; host arrays are copied, wider floats must be
    representable (ranks and ties are decided on the values the caller passed)."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dim() != 2:
        raise ValueError(f"{what}: expected [n_lists, seq_len], got {tuple(t.shape)}")
    if t.dtype != torch.float32:
        t32 = t.to(torch.float32)
        if not torch.equal(t32.to(t.dtype), t):
            raise NotImplementedError(f"{what}: {t.dtype} values that float32 cannot represent")
        t = t32
    return t.to(torch.device("cuda", torch.cuda.current_device())).contiguous()


_INV_LOG2 = {}


def _rank_metrics(labels, predictions, want_dcg: bool, want_auc: bool):
    if not torch.cuda.is_available():
        raise RuntimeError("utils.metrics.Metric needs a CUDA device (rlt_b200 has no CPU fallback)")
    y, s = _as_f32_device(labels, "labels"), _as_f32_device(predictions, "predictions")
    if y.shape != s.shape:
        raise ValueError(f"labels {tuple(y.shape)} and predictions {tuple(s.shape)} differ")
    n, L = y.shape
    key = (L, y.device.index)
    if key not in _INV_LOG2:   # the reference's own expression, evaluated by the host libm (utils/metrics.py:56)
        _INV_LOG2[key] = torch.tensor([1 / math.log2(i + 2) for i in range(L)], dtype=torch.float64).to(y.device)
    dcg = torch.empty(n, dtype=torch.float64, device=y.device) if want_dcg else None
    auc = torch.empty(n, dtype=torch.float64, device=y.device) if want_auc else None
    valid = torch.empty(n, dtype=torch.int32, device=y.device) if want_auc else None
    check(ops.lib().rlt_rank_metrics(ptr(s), ptr(y), ptr(_INV_LOG2[key]), n, L, ptr(dcg), ptr(auc), ptr(valid), stream_ptr()),
          "rlt_rank_metrics")
    return (dcg.cpu().numpy() if want_dcg else None, auc.cpu().numpy() if want_auc else None,
            valid.cpu().numpy() if want_auc else None)


class Metric:
    """Reference utils/metrics.py:9-38 — k is a COUNT of kept documents, not an index."""

    def __init__(self):
        pass

    @classmethod
    def f1(cls, labels: np.array, k_s: list):
        cnt, f1, _, pyint = _eval_given_k(labels, k_s)
        # the reference appends a Python int 0 when p + r == 0 (<=> no relevant document in the cut), a
        # float32 when k is a Python int and a float64 when k is a numpy integer; np.mean sees that list
        vals = [0 if cnt[i] == 0 else (np.float32(f1[i]) if pyint[i] else np.float64(f1[i])) for i in range(len(cnt))]
        return np.mean(vals)

    @classmethod
    def dcg(cls, labels: np.array, k_s: list, penalty=-1):
        if penalty != -1:
            raise NotImplementedError("only the reference's default penalty=-1 is implemented")
        _, _, dcg, _ = _eval_given_k(labels, k_s)
        return np.mean([np.float64(v) for v in dcg])

    @classmethod
    def taskr_metric(cls, labels: np.array, predictions: np.array):
        """Reference utils/metrics.py:40-58: mean over lists of the +-1/log2(i+2) DCG of the whole list re-ordered by
        descending prediction.  Tied predictions keep their list order (the reference inherits numpy's unstable
        argsort for them)."""
        dcg, _, _ = _rank_metrics(labels, predictions, True, False)
        return np.mean([float(v) for v in dcg])

    @classmethod
    def taskc_metric(cls, labels: np.array, predictions: np.array):
        """Reference utils/metrics.py:60-76: mean roc_auc_score over the lists that contain both classes
        (ZeroDivisionError when there is none, as in the reference)."""
        _, auc, valid = _rank_metrics(labels, predictions, False, True)
        tmp_auc, count_auc = 0, 0
        for a, v in zip(auc, valid):
            if not v:
                continue
            tmp_auc += np.float64(a)
            count_auc += 1
        return tmp_auc / count_auc


class Metric_for_Loss:
    """Reference utils/metrics.py:79-101: the reward of cutting one list after k documents."""

    def __init__(self) -> None:
        pass

    @staticmethod
    def _row(label: torch.Tensor, metric: str) -> torch.Tensor:
        if not torch.cuda.is_available():
            raise RuntimeError("Metric_for_Loss needs a CUDA device (rlt_b200 has no CPU fallback)")
        y = label.detach().to(device="cuda", dtype=torch.float32).reshape(1, -1).contiguous()
        r = torch.empty_like(y)
        ops.reward_matrix(y, r, metric)
        return r[0]

    @classmethod
    def f1(cls, label: torch.Tensor, k: int):
        return cls._row(label, "f1")[k - 1].to(label.device)

    @classmethod
    def dcg(cls, label: torch.Tensor, k: int, penalty: int = -1):
        if penalty != -1:
            raise NotImplementedError("only the reference's default penalty=-1 is implemented")
        return cls._row(label, "dcg")[k - 1].to(label.device)
