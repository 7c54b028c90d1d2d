"""Drop-in replacement of the reference utils/metrics.py (Metric, Metric_for_Loss).

Counting, cut selection and the order-exact float64 DCG summation run on the GPU (K4,
rlt_eval_cut); only the final np.mean over the per-list values is taken on the host, exactly as the
reference does (utils/metrics.py:24,38).  `taskr_metric` / `taskc_metric` (verify scripts only) are
outside the hot path and not provided.
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
