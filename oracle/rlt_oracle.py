"""CPU oracle: a plain restatement of the reference's hot path (TEST INFRASTRUCTURE, not product).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this module.
The product (ranked-list-truncation_b200/) never does.

Every function cites the reference code it restates (paths relative to the reference tree).  The
heavy tensor math of the reference lives in torch (nn.LSTM, nn.TransformerEncoder — not pinned by
the reference; this container has torch 2.11.0); those blocks are restated here from their
published equations with elementary torch ops so they run in float32 or float64 and are
differentiable (autograd supplies the oracle gradients).

Pinning: `tests/test_oracle_golden.py` checks this file against
  * the reference's only known-answer vector (utils/metrics.py:104-109), and
  * tests/golden/*.npz, produced by oracle/make_golden.py from the UNMODIFIED reference imported in
    the build container (oracle/refshim.py), for every model family and loss dispatched by
    run.py:59-102.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch

Tensor = torch.Tensor

# ------------------------------------------------------------------------------------------------
# Metrics (utils/metrics.py) — bit-exact restatement
# ------------------------------------------------------------------------------------------------
# utils/metrics.py:7 builds the table with math.log(j+2, 2) (NOT math.log2: they differ in the last
# bit at 112 of the first 300 positions).  The reference stops at 300 entries; the same expression
# is extended so that L up to 1000 can be evaluated (BASELINE.json config 5).
MAX_LEN = 1000
DCG_COEF = [math.log(j + 2, 2) for j in range(MAX_LEN)]
DCG_TERM64 = np.array([1.0 / c for c in DCG_COEF], dtype=np.float64)          # 1/coef in float64 (Metric.dcg)
DCG_COEF32 = np.array(DCG_COEF, dtype=np.float32)                             # float32 table (Metric_for_Loss.dcg)


def cut_positions(probs: np.ndarray) -> np.ndarray:
    """run.py:140-142 (and :137-139 on output[-1]): k = argmax over positions + 1, first max on ties."""
    p = np.asarray(probs)
    if p.ndim == 3:
        p = p[..., 0]
    return np.argmax(p, axis=1) + 1


def bicut_cut_positions(out: np.ndarray) -> list:
    """run.py:132-136: class argmax per position (tie -> class 0); k = L if every position says
    'continue', else index of the FIRST 'truncate' + 1.  Entries are Python ints (k = L) or np.int64."""
    pred = np.argmax(np.asarray(out), axis=2)
    seq_len = pred.shape[1]
    ks = []
    for row in pred:
        ks.append(seq_len if row.sum() == seq_len else np.argmin(row) + 1)
    return ks


def f1_per_list(labels: np.ndarray, k_s: Sequence) -> list:
    """utils/metrics.py:15-24, one value per list, with numpy's scalar promotion rules: precision
    is float64 when k is a numpy integer and float32 when k is a Python int; recall is float32."""
    labels = np.asarray(labels)
    n_rel = np.sum(labels, axis=1)
    vals = []
    for i in range(labels.shape[0]):
        k = k_s[i]
        hit = np.sum(labels[i, :k])
        prec = hit / k
        rec = (hit / n_rel[i]) if n_rel[i] != 0 else 0
        vals.append((2 * prec * rec / (prec + rec)) if prec + rec != 0 else 0)
    return vals


def metric_f1(labels: np.ndarray, k_s: Sequence):
    """Metric.f1 (utils/metrics.py:15-24)."""
    return np.mean(f1_per_list(labels, k_s))


def dcg_per_list(labels: np.ndarray, k_s: Sequence, penalty=-1) -> list:
    """utils/metrics.py:26-38: sum_{j<k} (+1 if label==1 else penalty) / log(j+2, 2) in float64, summed by
    ndarray.sum() (numpy pairwise order)."""
    labels = np.asarray(labels)
    vals = []
    for i in range(labels.shape[0]):
        k = int(k_s[i])
        head = labels[i, :k]
        coef = DCG_COEF[:k]
        good = (head == 1).astype(float)
        bad = (head != 1).astype(float)
        vals.append((good / coef + penalty * bad / coef).sum())
    return vals


def metric_dcg(labels: np.ndarray, k_s: Sequence, penalty=-1):
    """Metric.dcg (utils/metrics.py:26-38)."""
    return np.mean(dcg_per_list(labels, k_s, penalty))


def numpy_pairwise_sum(a: Sequence[float]) -> float:
    """Pure-Python model of numpy's float64 pairwise summation (the order ndarray.sum() uses for a
    contiguous 1-D array).  The CUDA eval kernel follows exactly this order; tests check this model
    against ndarray.sum() bit for bit."""
    n = len(a)
    if n < 8:
        s = 0.0 if n == 0 else -0.0
        # numpy starts from the first element for n<8 via the generic reduce loop (identity not added
        # to the data); summing left to right from a[0] is the same thing.
        if n == 0:
            return 0.0
        s = float(a[0])
        for v in a[1:]:
            s = s + float(v)
        return s
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])


# ------------------------------------------------------------------------------------------------
# Reward matrices (utils/metrics.py:85-101, looped over every (list, cut) cell by the losses)
# ------------------------------------------------------------------------------------------------
def reward_f1_cell(label: Tensor, k: int) -> Tensor:
    """Metric_for_Loss.f1 (utils/metrics.py:85-91) for one list and one cut position."""
    n_rel = label.sum()
    hit = label[:k].sum()
    prec = hit / k
    rec = hit / n_rel if n_rel != 0 else torch.tensor(0)
    return prec * rec * 2 / (prec + rec) if (prec + rec) != 0 else torch.tensor(0)


def reward_dcg_cell(label: Tensor, k: int, penalty: int = -1) -> Tensor:
    """Metric_for_Loss.dcg (utils/metrics.py:93-101): float32 coefficients, float32 sum."""
    head = label[:k]
    coef = torch.tensor(DCG_COEF[:k])
    return ((head == 1.).float() / coef + ((head != 1.).float() / coef) * penalty).sum()


def reward_matrix_loop(labels: Tensor, metric: str) -> Tensor:
    """The B x L Python double loop of utils/losses.py:56-65 / :80-89 / :216-225 (slow, faithful)."""
    B, L = labels.shape
    r = torch.ones(B, L, dtype=labels.dtype)
    cell = reward_f1_cell if metric == "f1" else reward_dcg_cell
    for i in range(B):
        for j in range(L):
            r[i][j] = cell(labels[i], j + 1)
    return r


def reward_matrix(labels: Tensor, metric: str) -> Tensor:
    """Vectorised equivalent of reward_matrix_loop.  F1 follows the reference operation order
    (p=c/k, r=c/N, 2pr/(p+r)); DCG is a float32 prefix sum of +-1/coef32 (the reference sums each
    prefix separately, so DCG agrees to float32 rounding, not bit for bit)."""
    B, L = labels.shape
    dt = labels.dtype
    if metric == "f1":
        c = torch.cumsum(labels, dim=1)
        k = torch.arange(1, L + 1, dtype=dt, device=labels.device).unsqueeze(0)
        n_rel = labels.sum(dim=1, keepdim=True)
        prec = c / k
        rec = torch.where(n_rel != 0, c / torch.where(n_rel != 0, n_rel, torch.ones_like(n_rel)), torch.zeros_like(c))
        den = prec + rec
        return torch.where(den != 0, prec * rec * 2 / torch.where(den != 0, den, torch.ones_like(den)),
                           torch.zeros_like(den))
    coef = torch.tensor(DCG_COEF[:L], dtype=torch.float32).to(dt).to(labels.device).unsqueeze(0)
    sign = torch.where(labels == 1., torch.ones_like(labels), -torch.ones_like(labels))
    return torch.cumsum(sign / coef, dim=1)


def reward_distribution(r: Tensor, tau: float) -> Tensor:
    """q = exp(r/tau) / sum exp(r/tau) over positions (utils/losses.py:90-92, :226-228; no max-shift)."""
    q = torch.exp(r / tau)
    return q / q.sum(dim=1, keepdim=True)


# ------------------------------------------------------------------------------------------------
# Losses (utils/losses.py).  `p` is the model output [B, L, 1] (probabilities), labels [B, L].
# `loop=True` uses the faithful Python reward loop (CPU-baseline timing), else the vectorised one.
# ------------------------------------------------------------------------------------------------
def _rewards(labels: Tensor, metric: str, loop: bool) -> Tensor:
    return reward_matrix_loop(labels, metric) if loop else reward_matrix(labels, metric)


def choopy_loss(p: Tensor, labels: Tensor, metric: str = "f1", loop: bool = False) -> Tensor:
    """ChoopyLoss.forward (utils/losses.py:55-68): -(1/B) sum p*r."""
    r = _rewards(labels, metric, loop)
    return -(p.squeeze() * r).sum() / p.shape[0]


def attncut_loss(p: Tensor, labels: Tensor, metric: str = "f1", tau: float = 0.95, loop: bool = False) -> Tensor:
    """AttnCutLoss.forward (RAML, utils/losses.py:79-96): -(1/B) sum q*log p."""
    q = reward_distribution(_rewards(labels, metric, loop), tau)
    return -(torch.log(p.squeeze()) * q).sum() / p.shape[0]


def _kl_batchmean(log_input: Tensor, target: Tensor) -> Tensor:
    """torch.nn.KLDivLoss(reduction='batchmean'): sum target*(log target - input) / batch (xlogy: 0 where target==0)."""
    return (torch.xlogy(target, target) - target * log_input).sum() / log_input.shape[0]


def div_loss(p: Tensor, labels: Tensor, metric: str = "f1", tau: float = 0.85, div_type: str = "kl",
             augmented: bool = True, loop: bool = False) -> Tensor:
    """DivLoss.forward (utils/losses.py:216-233): KL(q||p) or the Jensen-Shannon form against m=(p+q)/2."""
    tau_eff = tau if augmented else 1.0
    q = reward_distribution(_rewards(labels, metric, loop), tau_eff)
    ps = p.squeeze()
    if div_type == "kl":
        return _kl_batchmean(ps.log(), q)
    log_m = ((ps + q) / 2).log()
    return (_kl_batchmean(log_m, q) + _kl_batchmean(log_m, ps)) / 2


def rerank_loss(scores: Tensor, labels: Tensor, margin: float = 5e-4) -> Tensor:
    """RerankLoss.forward (utils/losses.py:127-141): batch-global hinge on mean(irrelevant) - mean(relevant).
    Degenerate batches make the reference build an integer tensor with requires_grad -> RuntimeError."""
    rel = labels == 1.
    irr = labels == 0.
    n_rel, n_irr = int(rel.sum()), int(irr.sum())
    if n_rel == 0 or n_irr == 0:
        raise RuntimeError("Only Tensors of floating point and complex dtype can require gradients")
    s = scores.squeeze()
    gap = (irr * s).sum() / n_irr - (rel * s).sum() / n_rel + margin
    return gap if gap > 0 else torch.zeros((), dtype=scores.dtype, requires_grad=True)


def pack_labels(labels) -> np.ndarray:
    """Bit-mask label format of row N3 (include/rlt_b200.h, rlt_pack_labels): [n, ceil(L/32)] uint32, bit i%32 of word
    i/32 set when document i is relevant (label == 1., dataloader/attncut_dataloader.py:47)."""
    y = np.asarray(labels)
    n, L = y.shape
    words = (L + 31) // 32
    padded = np.zeros((n, words * 32), dtype=np.uint8)
    padded[:, :L] = (y == 1)
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(n, words)


def unpack_labels(bits, seq_len: int) -> np.ndarray:
    b = np.ascontiguousarray(np.asarray(bits).astype("<u4")).view(np.uint8)
    return np.unpackbits(b, axis=1, bitorder="little")[:, :seq_len].astype(np.float32)


def loader_batches(X, y, batch_size: int, order):
    """What `DataLoader(TensorDataset(X, y), batch_size, shuffle=True)` (attncut_dataloader.py:86-87) collates for a given
    visiting order: consecutive slices of the order, last one partial."""
    X, y, order = np.asarray(X), np.asarray(y), np.asarray(order)
    return [(X[order[lo:lo + batch_size]], y[order[lo:lo + batch_size]]) for lo in range(0, len(order), batch_size)]


def taskr_dcg_per_list(labels, predictions) -> list:
    """The DCG_sample values of Metric.taskr_metric (utils/metrics.py:51-57): documents in descending-prediction order,
    +-1/log2(i+2) accumulated left to right as Python floats.  Ties: list order (stable sort); the reference's default
    np.argsort is unstable, so its own answer for tied predictions is implementation-defined."""
    out = []
    for pred, lab in zip(np.asarray(predictions), np.asarray(labels)):
        acc = 0
        for i, origin in enumerate(np.argsort(-pred, kind="stable")):
            acc += (1 / math.log2(i + 2)) if lab[origin] else (-1 / math.log2(i + 2))
        out.append(acc)
    return out


def taskr_metric(labels, predictions):
    """Metric.taskr_metric (utils/metrics.py:40-58)."""
    return np.mean(taskr_dcg_per_list(labels, predictions))


def auc_per_list(labels, predictions):
    """roc_auc_score per list (utils/metrics.py:73) as the Mann-Whitney statistic: the share of (relevant, irrelevant)
    pairs ranked correctly, ties counted half.  Returns (auc [n] float64, valid [n] bool); one-class lists
    (utils/metrics.py:72) are invalid."""
    labels, predictions = np.asarray(labels), np.asarray(predictions)
    auc, valid = np.zeros(len(labels)), np.zeros(len(labels), dtype=bool)
    for b, (lab, pred) in enumerate(zip(labels, predictions)):
        pos, neg = pred[lab == 1], pred[lab != 1]
        if len(pos) == 0 or len(neg) == 0:
            continue
        u2 = 2 * int((pos[:, None] > neg[None, :]).sum()) + int((pos[:, None] == neg[None, :]).sum())
        auc[b], valid[b] = u2 / (2.0 * len(pos) * len(neg)), True
    return auc, valid


def taskc_metric(labels, predictions):
    """Metric.taskc_metric (utils/metrics.py:60-76): running sum of the per-list AUCs over the two-class lists, divided
    by their number."""
    auc, valid = auc_per_list(labels, predictions)
    tmp_auc, count_auc = 0, 0
    for a, v in zip(auc, valid):
        if v:
            tmp_auc += np.float64(a)
            count_auc += 1
    return tmp_auc / count_auc


def bce_loss(p: Tensor, labels: Tensor) -> Tensor:
    """nn.BCELoss on squeezed probabilities (verify_probe.py:82, :193-196; the class term of MtCutLoss, losses.py:189):
    mean over all B*L entries of -(y log p + (1-y) log(1-p)) with torch's clamp of each log at -100."""
    q = p.squeeze()
    lp, l1p = torch.log(q).clamp_min(-100.), torch.log(1. - q).clamp_min(-100.)
    return -(labels * lp + (1. - labels) * l1p).mean()


def mtcut_loss(outputs: Sequence[Tensor], labels: Tensor, metric: str = "f1", rerank_weight: float = 0.5,
               classi_weight: float = 0.5, num_tasks: float = 3, loop: bool = False) -> Tensor:
    """MtCutLoss.forward (utils/losses.py:180-191)."""
    if num_tasks == 3:
        cls_p, rerank_s, cut_p = outputs
    elif num_tasks == 2.1:
        cls_p, cut_p = outputs
    else:
        rerank_s, cut_p = outputs
    total = div_loss(cut_p, labels, metric=metric, tau=0.85, div_type="js", augmented=True, loop=loop)
    if num_tasks == 3 or num_tasks == 2.2:
        total = total + rerank_loss(rerank_s, labels) * rerank_weight
    if num_tasks == 3 or num_tasks == 2.1:
        total = total + torch.nn.functional.binary_cross_entropy(cls_p.squeeze(), labels) * classi_weight
    return total


def bicut_loss(out: Tensor, labels: Tensor, alpha: float = 0.65, r: float = 0.0971134020,
               metric: str = "f1") -> Tensor:
    """BiCutLoss.forward (utils/losses.py:31-45) with slice_index (:21-29): the mask keeps positions up to
    and including the LAST position whose argmax class is 0 ('truncate'); all positions if there is none."""
    B, L, _ = out.shape
    pred = torch.argmax(out, dim=2)
    mask = torch.ones_like(out)
    for i in range(B):
        zeros = (pred[i] == 0).nonzero()
        last = L if zeros.numel() == 0 else int(zeros[-1])
        mask[i, last + 1:] = 0
    w = torch.zeros_like(out)
    if metric == "nci":
        j = torch.arange(L, dtype=out.dtype)
        w[..., 1] = torch.where(labels == 1, -1 / torch.log2(j + 2).unsqueeze(0), ((j + 1) / alpha).unsqueeze(0))
    else:
        w[..., 0] = torch.where(labels == 1, torch.full_like(labels, (1 - alpha) / r), torch.zeros_like(labels))
        w[..., 1] = torch.where(labels == 1, torch.zeros_like(labels), torch.full_like(labels, alpha / (1 - r)))
    return (out * mask * w).sum() / B


# ------------------------------------------------------------------------------------------------
# Model blocks (torch.nn.LSTM, torch.nn.TransformerEncoderLayer restated with elementary ops)
# ------------------------------------------------------------------------------------------------
def lstm_direction(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, reverse: bool) -> Tensor:
    """One direction of one nn.LSTM layer (batch_first): gates in row order i, f, g, o; h0 = c0 = 0."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    pre = x @ w_ih.t() + (b_ih + b_hh)
    outs = [None] * L
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        a = pre[:, t] + h @ w_hh.t()
        i, f, g, o = a.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def bilstm(x: Tensor, sd: Dict[str, Tensor], prefix: str, num_layers: int = 2) -> Tensor:
    """nn.LSTM(num_layers=2, batch_first=True, bidirectional=True) as used at models/Bicut.py:8-9,
    AttnCut.py:8, MtAttnCut.py:8, MMOECut.py:63: layer l+1 consumes [fwd || bwd] of layer l."""
    y = x
    for layer in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            outs.append(lstm_direction(y, sd[f"{prefix}weight_ih_l{layer}{suffix}"], sd[f"{prefix}weight_hh_l{layer}{suffix}"],
                                       sd[f"{prefix}bias_ih_l{layer}{suffix}"], sd[f"{prefix}bias_hh_l{layer}{suffix}"], rev))
        y = torch.cat(outs, dim=2)
    return y


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def encoder_layer(x: Tensor, sd: Dict[str, Tensor], prefix: str, n_head: int, attend: str = "lists",
                  masks: Dict[str, Tensor] = None) -> Tensor:
    """nn.TransformerEncoderLayer (post-norm, ReLU, dim_feedforward=2048, eps=1e-5).  Eval / p=0 mode by default;
    `masks` (train mode) holds the keep-and-scale factors (0 or 1/(1-p)) of torch's four dropout sites
    (torch/nn/modules/transformer.py::_sa_block/_ff_block, functional.multi_head_attention_forward):
    'attn' [L, n_head, B, B] on the attention probabilities, 'after_attn' [B, L, d] (dropout1), 'ffn' [B, L, d_ff]
    inside the feed-forward block, 'after_ffn' [B, L, d] (dropout2).
    The reference builds it WITHOUT batch_first and feeds [B, L, d] (models/Choopy.py:11,21 etc.), so the
    attention runs over dim 0 — across the B lists of the call — independently per position and head
    (attend='lists').  attend='positions' is the batch_first behaviour (attention within a list)."""
    B, L, d = x.shape
    dh = d // n_head
    w_in, b_in = sd[prefix + "self_attn.in_proj_weight"], sd[prefix + "self_attn.in_proj_bias"]
    qkv = x @ w_in.t() + b_in
    q, k, v = qkv.split(d, dim=2)

    def heads(t):  # -> [L, n_head, B, dh] (lists) or [B, n_head, L, dh] (positions)
        t = t.reshape(B, L, n_head, dh)
        return t.permute(1, 2, 0, 3) if attend == "lists" else t.permute(0, 2, 1, 3)

    qh, kh, vh = heads(q), heads(k), heads(v)
    att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    if masks is not None and "attn" in masks:
        att = att * masks["attn"]
    oh = att @ vh
    o = (oh.permute(2, 0, 1, 3) if attend == "lists" else oh.permute(0, 2, 1, 3)).reshape(B, L, d)
    o = o @ sd[prefix + "self_attn.out_proj.weight"].t() + sd[prefix + "self_attn.out_proj.bias"]
    if masks is not None and "after_attn" in masks:
        o = o * masks["after_attn"]
    y = layer_norm(x + o, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    hdn = torch.relu(y @ sd[prefix + "linear1.weight"].t() + sd[prefix + "linear1.bias"])
    if masks is not None and "ffn" in masks:
        hdn = hdn * masks["ffn"]
    f = hdn @ sd[prefix + "linear2.weight"].t() + sd[prefix + "linear2.bias"]
    if masks is not None and "after_ffn" in masks:
        f = f * masks["after_ffn"]
    return layer_norm(y + f, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])


def encoder_stack(x: Tensor, sd: Dict[str, Tensor], prefix: str, n_head: int, attend: str = "lists") -> Tensor:
    n = 0
    while f"{prefix}layers.{n}.linear1.weight" in sd:
        n += 1
    for i in range(n):
        x = encoder_layer(x, sd, f"{prefix}layers.{i}.", n_head, attend)
    return x


def _linear(x, sd, name):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


# ------------------------------------------------------------------------------------------------
# Model forwards from a reference state_dict (dropout = 0 / eval mode)
# ------------------------------------------------------------------------------------------------
def bicut_forward(sd, x):
    """models/Bicut.py:18-21: BiLSTM -> Linear(256,256) -> ReLU -> Linear(256,2) -> softmax over the 2 classes."""
    h = bilstm(x, sd, "bilstm.")
    u = _linear(torch.relu(_linear(h, sd, "fc")), sd, "softmax.1")
    return torch.softmax(u, dim=2)


def choopy_trunk(sd, x, enc_prefix, n_head=8, attend="lists"):
    """models/Choopy.py:18-21 / MtChoopy.py:24-26: concat score with the learned [L,127] table, then the encoder."""
    pe = sd["position_encoding"].unsqueeze(0).expand(x.shape[0], -1, -1)
    return encoder_stack(torch.cat((x, pe), dim=2), sd, enc_prefix, n_head, attend)


def choopy_forward(sd, x, n_head=8, attend="lists"):
    """models/Choopy.py:18-23."""
    z = _linear(choopy_trunk(sd, x, "attention_layer.", n_head, attend), sd, "decison_layer.0")
    return torch.softmax(z, dim=1)


def attncut_forward(sd, x, n_head=4, attend="lists"):
    """models/AttnCut.py:16-20."""
    h = encoder_stack(bilstm(x, sd, "encoding_layer."), sd, "attention_layer.", n_head, attend)
    return torch.softmax(_linear(h, sd, "decison_layer.0"), dim=1)


def _mt_heads(sd, h, num_tasks):
    """models/MtChoopy.py:27-32 / MtAttnCut.py:24-29: sigmoid class head, linear rerank head (no softmax), softmax cut head."""
    y0 = torch.sigmoid(_linear(h, sd, "classi.0"))
    y1 = _linear(h, sd, "rerank")
    y2 = torch.softmax(_linear(h, sd, "decison_layer.0"), dim=1)
    if num_tasks == 3:
        return [y0, y1, y2]
    if num_tasks == 2.1:
        return [y0, y2]
    return [y1, y2]


def mtchoopy_forward(sd, x, num_tasks=3, n_head=8, attend="lists"):
    """models/MtChoopy.py:23-32."""
    return _mt_heads(sd, choopy_trunk(sd, x, "encoding_layer.", n_head, attend), num_tasks)


def mtattncut_forward(sd, x, num_tasks=3, n_head=4, attend="lists"):
    """models/MtAttnCut.py:21-29."""
    h = encoder_stack(bilstm(x, sd, "pre_encoding."), sd, "encoding_layer.", n_head, attend)
    return _mt_heads(sd, h, num_tasks)


def mmoecut_forward(sd, x, num_tasks=3, n_head=4, attend="lists"):
    """models/MMOECut.py:86-110: BiLSTM -> E encoder experts on the same input -> per-task softmax gate over the
    flattened LSTM output -> gate-weighted mixture -> towers (sigmoid class | softmax rerank | softmax cut)."""
    h = bilstm(x, sd, "pre_encoding.")
    B = h.shape[0]
    n_exp = 0
    while f"experts.{n_exp}.attention_layer.layers.0.linear1.weight" in sd:
        n_exp += 1
    experts = torch.stack([encoder_stack(h, sd, f"experts.{e}.attention_layer.", n_head, attend) for e in range(n_exp)])
    outs = []
    if num_tasks == 3:
        towers = [("classification_layer", "sigmoid"), ("rerank_layer", "softmax"), ("cut_layer", "softmax")]
    elif num_tasks == 2.1:
        towers = [("classification_layer", "sigmoid"), ("cut_layer", "softmax")]
    else:
        towers = [("rerank_layer", "softmax"), ("cut_layer", "softmax")]
    for t, (name, act) in enumerate(towers):
        gate = torch.softmax(h.reshape(B, -1) @ sd[f"w_gates.{t}"], dim=1)            # [B, E]
        mix = (gate.t().reshape(n_exp, B, 1, 1) * experts).sum(dim=0)
        z = _linear(mix, sd, f"towers.{t}.{name}.0")
        outs.append(torch.sigmoid(z) if act == "sigmoid" else torch.softmax(z, dim=1))
    return outs


_TOWERS3 = [("classification_layer", "sigmoid"), ("rerank_layer", "softmax"), ("cut_layer", "softmax")]


def _experts(sd, h, n_head, attend):
    n_exp = 0
    while f"experts.{n_exp}.attention_layer.layers.0.linear1.weight" in sd:
        n_exp += 1
    return [encoder_stack(h, sd, f"experts.{e}.attention_layer.", n_head, attend) for e in range(n_exp)]


def _tower_out(sd, mix, t, name, act):
    z = _linear(mix, sd, f"towers.{t}.{name}.0")
    return torch.sigmoid(z) if act == "sigmoid" else torch.softmax(z, dim=1)


def moecut_forward(sd, x, num_tasks=3, n_head=4, attend="lists"):
    """models/MOECut.py:86-109: as MMOECut, but ONE gate (`w_gates`, a single Parameter, :68) whose softmax weights
    (:94) mix the experts once (:100-101); every tower reads the same mixture (:104)."""
    h = bilstm(x, sd, "pre_encoding.")
    B = h.shape[0]
    experts = torch.stack(_experts(sd, h, n_head, attend))
    gate = torch.softmax(h.reshape(B, -1) @ sd["w_gates"], dim=1)
    mix = (gate.t().reshape(experts.shape[0], B, 1, 1) * experts).sum(dim=0)
    towers = _TOWERS3 if num_tasks == 3 else ([_TOWERS3[0], _TOWERS3[2]] if num_tasks == 2.1 else _TOWERS3[1:])
    return [_tower_out(sd, mix, t, name, act) for t, (name, act) in enumerate(towers)]


def plecut_forward(sd, x, n_head=2, attend="lists"):
    """models/PLECut.py:77-104: gates of width 2, 2, 3 (:68-70) over the expert subsets [0:2], [1:3], [0:3]
    (:81-83, :94-96), one mixture per tower."""
    h = bilstm(x, sd, "pre_encoding.")
    B = h.shape[0]
    experts = _experts(sd, h, n_head, attend)
    outs = []
    for t, ((name, act), (lo, hi)) in enumerate(zip(_TOWERS3, ((0, 2), (1, 3), (0, 3)))):
        sub = torch.stack(experts[lo:hi])
        gate = torch.softmax(h.reshape(B, -1) @ sd[f"w_gates.{t}"], dim=1)
        mix = (gate.t().reshape(hi - lo, B, 1, 1) * sub).sum(dim=0)
        outs.append(_tower_out(sd, mix, t, name, act))
    return outs


def probebase_forward(sd, x, n_head=4, attend="lists"):
    """models/Probe.py:76-99: the MMOECut graph (two experts by default, one gate per task :68, towers class | rerank |
    cut :69-73) returning `(experts_in, experts_o, final_output)` -- the LSTM representation, the list of expert
    outputs and the list of tower outputs."""
    h = bilstm(x, sd, "pre_encoding.")
    B = h.shape[0]
    experts = _experts(sd, h, n_head, attend)
    stacked = torch.stack(experts)
    outs = []
    for t, (name, act) in enumerate(_TOWERS3):
        if f"w_gates.{t}" not in sd:
            break                                                                     # zip(towers, towers_input), :93
        gate = torch.softmax(h.reshape(B, -1) @ sd[f"w_gates.{t}"], dim=1)
        mix = (gate.t().reshape(len(experts), B, 1, 1) * stacked).sum(dim=0)
        outs.append(_tower_out(sd, mix, t, name, act))
    return h, experts, outs


def probe_forward(sd, experts_in, experts_o):
    """models/Probe.py:113-122: class (sigmoid) and rerank (softmax over positions) probes on the LSTM representation
    and on both expert outputs; returned in the reference's order c1, r1, ce1, ce2, re1, re2."""
    def cls(name, h):
        return torch.sigmoid(_linear(h, sd, f"{name}.classification_layer.0"))

    def rer(name, h):
        return torch.softmax(_linear(h, sd, f"{name}.rerank_layer.0"), dim=1)
    return (cls("probe_c1", experts_in), rer("probe_r1", experts_in), cls("probe_ce1", experts_o[0]),
            cls("probe_ce2", experts_o[1]), rer("probe_re1", experts_o[0]), rer("probe_re2", experts_o[1]))


def flat_outputs(out):
    """One flat list of a model's output tensors: ProbeBase's `(experts_in, [experts], [towers])` in that order."""
    if isinstance(out, tuple):
        return [out[0], *out[1], *out[2]]
    return list(out) if isinstance(out, list) else [out]


def loss_input(out):
    """What the criterion reads: the tower outputs (`output[-1]`, verify_probe.py:107) for ProbeBase."""
    return out[-1] if isinstance(out, tuple) else out


FORWARDS = {
    "bicut": bicut_forward, "choopy": choopy_forward, "attncut": attncut_forward,
    "mtchoopy": mtchoopy_forward, "mtattncut": mtattncut_forward, "mmoecut": mmoecut_forward,
    "moecut": moecut_forward, "plecut": plecut_forward, "probebase": probebase_forward,
}
# the two-task variants of run.py's --num_tasks (2.1 = class + cut, 2.2 = rerank + cut; MtChoopy.py:27-32,
# MtAttnCut.py:24-29, MMOECut.py:74-84): "<model>_t21" / "<model>_t22"
for _base in ("mtchoopy", "mtattncut", "mmoecut"):
    for _tag, _nt in (("t21", 2.1), ("t22", 2.2)):
        FORWARDS[f"{_base}_{_tag}"] = (lambda sd, x, _f=FORWARDS[_base], _nt=_nt, **kw: _f(sd, x, num_tasks=_nt, **kw))


def num_tasks_of(model_name: str) -> float:
    return 2.1 if model_name.endswith("_t21") else 2.2 if model_name.endswith("_t22") else 3


def criterion_for(model_name: str, metric: str = "f1", loop: bool = False, **kw):
    """The criterion run.py:59-102 pairs with each model (defaults of its CLI: div_type 'js', augmented reward)."""
    if model_name == "bicut":
        return lambda out, y: bicut_loss(out, y, metric=metric)
    if model_name == "choopy":
        return lambda out, y: choopy_loss(out, y, metric=metric, loop=loop)
    if model_name == "attncut":
        return lambda out, y: div_loss(out, y, metric=metric, div_type=kw.get("div_type", "js"),
                                       augmented=kw.get("augmented", True), loop=loop)
    return lambda out, y: mtcut_loss(out, y, metric=metric, num_tasks=kw.get("num_tasks", num_tasks_of(model_name)),
                                     rerank_weight=kw.get("rerank_weight", 0.5),
                                     classi_weight=kw.get("classi_weight", 0.5), loop=loop)


# ----------------------------------------------------------------------------------------------
# optimizer step
# ----------------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avgs, exp_avg_sqs, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              grad_scale=1.0):
    """One step of the optimizer the reference builds at run.py:104 (`optim.Adam(params, lr, weight_decay)`), restated
    from torch/optim/adam.py `_single_tensor_adam` (amsgrad=False, maximize=False) operation by operation on float32
    numpy arrays; `step` counts from 1.  Updates the arrays in place.  Python-float scalars are formed in double and
    rounded to float32 where torch hands them to a float32 kernel."""
    f = np.float32
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    step_size = f(lr / bc1)
    bc2_sqrt = f(math.sqrt(bc2))
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        g = g.astype(f) * f(grad_scale)
        if weight_decay != 0:
            g = g + f(weight_decay) * p                      # grad.add(param, alpha=weight_decay)
        m += f(1.0 - b1) * (g - m)                           # exp_avg.lerp_(grad, 1 - beta1)   (weight < 0.5 branch)
        v *= f(b2)                                           # exp_avg_sq.mul_(beta2)
        v += (f(1.0 - b2) * g) * g                           #            .addcmul_(grad, grad, value=1 - beta2)
        denom = np.sqrt(v) / bc2_sqrt + f(eps)               # (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        p += (-step_size) * (m / denom)                      # param.addcdiv_(exp_avg, denom, value=-step_size)
