"""Thin Python wrappers over the C ABI (include/rlt_b200.h): one function per entry point.

All tensors must be CUDA float32 and contiguous; there is no CPU path.  Functions enqueue work on
torch's current stream and return immediately.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

MAX_LEN = 1024
# the reference's coefficient table (utils/metrics.py:7): math.log(j+2, 2), NOT math.log2
DCG_COEF = [math.log(j + 2, 2) for j in range(MAX_LEN)]

LOSS_KINDS = {"choopy": 0, "raml": 1, "kl": 2, "js": 3}


class EncoderDesc(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("group_size", C.c_int32), ("seq_len", C.c_int32), ("d_model", C.c_int32),
                ("n_head", C.c_int32), ("d_ff", C.c_int32), ("attend_axis", C.c_int32), ("accumulate_dx", C.c_int32),
                ("ln_eps", C.c_float), ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("inference", C.c_int32),
                ("reserved", C.c_int32)]


ENCODER_PARAM_ORDER = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                       "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                       "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")


class EncoderPtrs(C.Structure):
    """rlt_encoder_weights / rlt_encoder_grads (12 device pointers in ENCODER_PARAM_ORDER)."""
    _fields_ = [(n.replace(".", "_"), C.c_void_p) for n in ENCODER_PARAM_ORDER]


class CutLossDesc(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("seq_len", C.c_int32), ("input_kind", C.c_int32), ("loss_kind", C.c_int32),
                ("metric_dcg", C.c_int32), ("accumulate_loss", C.c_int32), ("tau", C.c_float),
                ("grad_scale", C.c_float), ("loss_scale", C.c_float)]


class AuxLossDesc(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("group_size", C.c_int32), ("seq_len", C.c_int32),
                ("rerank_softmax", C.c_int32), ("class_probs", C.c_int32), ("accumulate_loss", C.c_int32),
                ("margin", C.c_float),
                ("class_weight", C.c_float), ("rerank_weight", C.c_float), ("grad_scale", C.c_float),
                ("loss_scale", C.c_float)]


class BicutLossDesc(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("seq_len", C.c_int32), ("input_kind", C.c_int32), ("metric_nci", C.c_int32),
                ("accumulate_loss", C.c_int32), ("alpha", C.c_float), ("r", C.c_float), ("grad_scale", C.c_float),
                ("loss_scale", C.c_float)]


_configured = False
_tables_on = set()


def lib():
    """The loaded library with argument types declared."""
    global _configured
    L = _lib.load()
    if not _configured:
        L.rlt_encoder_layer_saved_bytes.restype = C.c_size_t
        L.rlt_encoder_layer_workspace_bytes.restype = C.c_size_t
        for name in ("rlt_bilstm_saved_bytes", "rlt_bilstm_workspace_bytes"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_size_t
        _configured = True
    return L


def _require(t: torch.Tensor, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"{name}: expected a contiguous float32 CUDA tensor (the rlt_b200 path has no CPU fallback); "
                           f"got {type(t).__name__} {getattr(t, 'dtype', None)} on {getattr(t, 'device', None)}")


def ensure_tables():
    """Upload the DCG tables to the current device once.  The upload is a synchronous copy (complete on return, so
    ordered before any later launch on any stream); Engine.__init__ and the criterion modules call this eagerly so that
    the first call never falls inside a CUDA-graph capture."""
    dev = torch.cuda.current_device()
    if dev in _tables_on:
        return
    coef32 = np.array(DCG_COEF, dtype=np.float32)
    term64 = np.array([1.0 / c for c in DCG_COEF], dtype=np.float64)
    check(lib().rlt_set_dcg_tables(coef32.ctypes.data_as(C.c_void_p), term64.ctypes.data_as(C.c_void_p), MAX_LEN),
          "rlt_set_dcg_tables")
    _tables_on.add(dev)


# ----------------------------------------------------------------------------------------------
# encoder layer
# ----------------------------------------------------------------------------------------------
def encoder_desc(n_groups, group_size, seq_len, d_model, n_head, d_ff=2048, ln_eps=1e-5, dropout_p=0.0, seed=0,
                 accumulate_dx=False, inference=False, attend_axis=0):
    """attend_axis 0: the lists of a group attend to each other per position (what the reference computes: no batch_first);
    1: the positions of each list attend to each other (the papers' intent)."""
    return EncoderDesc(n_groups, group_size, seq_len, d_model, n_head, d_ff, int(attend_axis), int(accumulate_dx), ln_eps,
                       dropout_p, seed, int(inference), 0)


def encoder_ptrs(tensors) -> EncoderPtrs:
    """tensors: 12 tensors in ENCODER_PARAM_ORDER."""
    for t, n in zip(tensors, ENCODER_PARAM_ORDER):
        _require(t, n)
    return EncoderPtrs(*[t.data_ptr() for t in tensors])


def encoder_saved_bytes(desc) -> int:
    return int(lib().rlt_encoder_layer_saved_bytes(C.byref(desc)))


def encoder_workspace_bytes(desc) -> int:
    return int(lib().rlt_encoder_layer_workspace_bytes(C.byref(desc)))


def encoder_layer_fwd(desc, weights, x, out, saved):
    _require(x, "x"); _require(out, "out")
    check(lib().rlt_encoder_layer_fwd(C.byref(desc), C.byref(weights), ptr(x), ptr(out), ptr(saved),
                                      C.c_size_t(saved.numel() * saved.element_size()), stream_ptr()),
          "rlt_encoder_layer_fwd")


def encoder_layer_bwd(desc, weights, grads, x, saved, d_out, d_x, workspace):
    check(lib().rlt_encoder_layer_bwd(C.byref(desc), C.byref(weights), C.byref(grads), ptr(x), ptr(saved), ptr(d_out),
                                      ptr(d_x), ptr(workspace),
                                      C.c_size_t(workspace.numel() * workspace.element_size()), stream_ptr()),
          "rlt_encoder_layer_bwd")


def attention_lists_fwd(qkv, n_groups, group_size, seq_len, d_model, n_head, want_lse=True):
    """Cross-list attention core (rlt_attention_lists_fwd): qkv [G*S*L, 3d] -> (o [G*S*L, d], lse [G*S*L, n_head])."""
    T = n_groups * group_size * seq_len
    o = torch.empty(T, d_model, dtype=torch.float32, device=qkv.device)
    lse = torch.empty(T, n_head, dtype=torch.float32, device=qkv.device) if want_lse else None
    check(lib().rlt_attention_lists_fwd(ptr(qkv), ptr(o), ptr(lse), int(n_groups), int(group_size), int(seq_len), int(d_model),
                                        int(n_head), stream_ptr()), "rlt_attention_lists_fwd")
    return o, lse


def ffn_fused_fwd(y16, y, w1_h, b1, w2_h, b2, gamma, beta, out, u2=None, stats=None, h_out=None, eps=1e-5):
    """out = LayerNorm(y + relu(y W1^T + b1) W2^T + b2) in one kernel (rlt_ffn_fused_fwd).  y16 / w1_h / w2_h: fp16."""
    T, d = y.shape
    f = w1_h.shape[0]
    check(lib().rlt_ffn_fused_fwd(ptr(y16), ptr(y), ptr(w1_h), ptr(b1), ptr(w2_h), ptr(b2), ptr(gamma), ptr(beta), ptr(out),
                                  ptr(u2), ptr(stats), ptr(h_out), int(T), int(d), int(f), C.c_float(eps), stream_ptr()),
          "rlt_ffn_fused_fwd")


# ----------------------------------------------------------------------------------------------
# heads / losses / eval
# ----------------------------------------------------------------------------------------------
def choopy_embed_fwd(score, pe, x):
    n_lists, seq_len = score.shape[0], score.shape[1]
    check(lib().rlt_choopy_embed_fwd(ptr(score), ptr(pe), ptr(x), n_lists, seq_len, stream_ptr()), "rlt_choopy_embed_fwd")


def choopy_embed_bwd(dx, dpe, n_lists, seq_len):
    check(lib().rlt_choopy_embed_bwd(ptr(dx), ptr(dpe), n_lists, seq_len, stream_ptr()), "rlt_choopy_embed_bwd")


def head_dots_fwd(x, w, bias, z, n_tokens, d, n_heads):
    check(lib().rlt_head_dots_fwd(ptr(x), ptr(w), ptr(bias), ptr(z), n_tokens, d, n_heads, stream_ptr()),
          "rlt_head_dots_fwd")


def head_dots_bwd(x, w, dz, dx, dw, db, n_tokens, d, n_heads, accumulate_dx=False, relu_gate=False):
    check(lib().rlt_head_dots_bwd(ptr(x), ptr(w), ptr(dz), ptr(dx), ptr(dw), ptr(db), n_tokens, d, n_heads,
                                  int(accumulate_dx), int(relu_gate), stream_ptr()), "rlt_head_dots_bwd")


def pair_softmax_fwd(z, o, n_tokens, dropout_p=0.0, seed=0):
    check(lib().rlt_pair_softmax_fwd(ptr(z), ptr(o), C.c_size_t(n_tokens), C.c_float(dropout_p), C.c_uint64(seed),
                                     stream_ptr()), "rlt_pair_softmax_fwd")


def pair_softmax_bwd(o, d_o, dz, n_tokens, dropout_p=0.0, seed=0):
    check(lib().rlt_pair_softmax_bwd(ptr(o), ptr(d_o), ptr(dz), C.c_size_t(n_tokens), C.c_float(dropout_p), C.c_uint64(seed),
                                     stream_ptr()), "rlt_pair_softmax_bwd")


def dropout_mask(seed, site, p, n, group_size=0, device="cuda"):
    """Test hook: the keep-and-scale factors (0 or 1/(1-p)) the kernels apply at dropout site `site`."""
    import torch
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(lib().rlt_dropout_mask(C.c_uint64(seed), int(site), C.c_float(p), C.c_size_t(n), int(group_size), ptr(out),
                                 stream_ptr()), "rlt_dropout_mask")
    return out


def linear(a, w, bias, out, relu=False):
    """out[M,N] = act(a[M,K] w[N,K]^T + bias)  (tcgen05 TF32 GEMM)."""
    M, K = a.shape
    N = w.shape[0]
    check(lib().rlt_linear(ptr(a), ptr(w), ptr(bias), ptr(out), M, N, K, C.c_float(1.0), int(relu), stream_ptr()), "rlt_linear")


def linear_nn(a, b, out):
    """out[M,N] = a[M,K] b[K,N]."""
    M, K = a.shape
    N = b.shape[1]
    check(lib().rlt_linear_nn(ptr(a), ptr(b), ptr(out), M, N, K, stream_ptr()), "rlt_linear_nn")


def grad_weight(a, b, out):
    """out[M,N] += a[T,M]^T b[T,N]."""
    T, M = a.shape
    N = b.shape[1]
    check(lib().rlt_grad_weight(ptr(a), ptr(b), ptr(out), T, M, N, C.c_float(1.0), stream_ptr()), "rlt_grad_weight")


def colsum(src, out):
    """out[C] += sum over rows of src[T, C]."""
    T, Cc = src.shape
    check(lib().rlt_colsum(ptr(src), ptr(out), T, Cc, stream_ptr()), "rlt_colsum")


# ----------------------------------------------------------------------------------------------
# BiLSTM
# ----------------------------------------------------------------------------------------------
class BilstmDesc(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("seq_len", C.c_int32), ("input_size", C.c_int32), ("hidden", C.c_int32),
                ("num_layers", C.c_int32), ("training", C.c_int32)]


class BilstmPtrs(C.Structure):
    """rlt_bilstm_weights / rlt_bilstm_grads: four [layer][direction] pointer tables."""
    _fields_ = [("w_ih", C.c_void_p * 4), ("w_hh", C.c_void_p * 4), ("b_ih", C.c_void_p * 4), ("b_hh", C.c_void_p * 4)]


def bilstm_ptrs(flat):
    """flat: nn.LSTM._flat_weights order = per layer, per direction: w_ih, w_hh, b_ih, b_hh (16 tensors)."""
    if len(flat) != 16:
        raise RuntimeError(f"expected the 16 parameter tensors of a 2-layer bidirectional LSTM, got {len(flat)}")
    for t in flat:
        _require(t, "lstm parameter")
    p = BilstmPtrs()
    for i in range(4):      # i = layer * 2 + direction
        p.w_ih[i], p.w_hh[i], p.b_ih[i], p.b_hh[i] = (flat[4 * i + k].data_ptr() for k in range(4))
    return p


def bilstm_desc(n_lists, seq_len, input_size, hidden=128, num_layers=2):
    return BilstmDesc(n_lists, seq_len, input_size, hidden, num_layers, 1)


def bilstm_sizes(desc):
    sv = int(lib().rlt_bilstm_saved_bytes(C.byref(desc)))
    ws = int(lib().rlt_bilstm_workspace_bytes(C.byref(desc)))
    if sv == 0 or ws == 0:
        raise _lib.RltError("bilstm: " + _lib.load().rlt_last_error().decode())
    return sv, ws


def bilstm_fwd(desc, weights, x, y, saved, workspace):
    check(lib().rlt_bilstm_fwd(C.byref(desc), C.byref(weights), ptr(x), ptr(y), ptr(saved),
                               C.c_size_t(0 if saved is None else saved.numel() * 4), ptr(workspace),
                               C.c_size_t(workspace.numel() * 4), stream_ptr()), "rlt_bilstm_fwd")


def bilstm_bwd(desc, weights, grads, x, saved, dy, dx, workspace):
    check(lib().rlt_bilstm_bwd(C.byref(desc), C.byref(weights), C.byref(grads), ptr(x), ptr(saved), ptr(dy), ptr(dx),
                               ptr(workspace), C.c_size_t(workspace.numel() * 4), stream_ptr()), "rlt_bilstm_bwd")


# ----------------------------------------------------------------------------------------------
# MMOECut gates + towers
# ----------------------------------------------------------------------------------------------
class MoeDesc(C.Structure):
    _fields_ = [("n_lists", C.c_int32), ("seq_len", C.c_int32), ("d_lstm", C.c_int32), ("d_model", C.c_int32),
                ("n_experts", C.c_int32), ("n_tasks", C.c_int32)]


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def moe_heads_fwd(desc, h_lstm, w_gates, experts, tower_w, tower_b, gates, z):
    check(lib().rlt_moe_heads_fwd(C.byref(desc), ptr(h_lstm), ptr(w_gates), _ptr_array(experts), ptr(tower_w),
                                  ptr(tower_b), ptr(gates), ptr(z), stream_ptr()), "rlt_moe_heads_fwd")


def moe_heads_bwd(desc, h_lstm, w_gates, experts, tower_w, gates, dz, d_experts, d_tower_w, d_tower_b, d_w_gates,
                  d_h_lstm, accumulate_dh, scratch):
    check(lib().rlt_moe_heads_bwd(C.byref(desc), ptr(h_lstm), ptr(w_gates), _ptr_array(experts), ptr(tower_w),
                                  ptr(gates), ptr(dz), _ptr_array(d_experts), ptr(d_tower_w), ptr(d_tower_b),
                                  ptr(d_w_gates), ptr(d_h_lstm), int(accumulate_dh), ptr(scratch), stream_ptr()),
          "rlt_moe_heads_bwd")


def softmax_lists(z, p, n_lists, seq_len):
    check(lib().rlt_softmax_lists(ptr(z), ptr(p), n_lists, seq_len, stream_ptr()), "rlt_softmax_lists")


def softmax_lists_bwd(p, dp, dz, n_lists, seq_len):
    check(lib().rlt_softmax_lists_bwd(ptr(p), ptr(dp), ptr(dz), n_lists, seq_len, stream_ptr()), "rlt_softmax_lists_bwd")


def cut_loss(inp, labels, *, loss_kind, metric="f1", tau=1.0, input_kind=0, probs_out=None, grad=None,
             loss_per_list=None, loss_out=None, grad_scale=1.0, loss_scale=1.0, accumulate=False, label_bits=None):
    """labels: [B, L] float32 -- or None with label_bits = the [B, ceil(L/32)] int32 / uint32 words of pack_labels
    (rlt_cut_loss_bits: logits in, even L)."""
    ensure_tables()
    n_lists, seq_len = inp.shape[0], inp.shape[1]
    desc = CutLossDesc(n_lists, seq_len, input_kind, LOSS_KINDS[loss_kind] if isinstance(loss_kind, str) else loss_kind,
                       0 if metric == "f1" else 1, int(accumulate), tau, grad_scale, loss_scale)
    if label_bits is not None:
        if label_bits.shape != (n_lists, (seq_len + 31) // 32) or label_bits.element_size() != 4:
            raise ValueError(f"label_bits must be [{n_lists}, {(seq_len + 31) // 32}] 32-bit words")
        check(lib().rlt_cut_loss_bits(C.byref(desc), ptr(inp), ptr(label_bits), ptr(probs_out), ptr(grad), ptr(loss_per_list),
                                      ptr(loss_out), stream_ptr()), "rlt_cut_loss_bits")
        return
    check(lib().rlt_cut_loss(C.byref(desc), ptr(inp), ptr(labels), ptr(probs_out), ptr(grad), ptr(loss_per_list),
                             ptr(loss_out), stream_ptr()), "rlt_cut_loss")


def pack_labels(labels, check_binary=True):
    """[B, L] float32 0./1. labels -> [B, ceil(L/32)] int32 bit masks (bit j%32 of word j//32 = label j), rlt_pack_labels."""
    n_lists, seq_len = labels.shape
    bits = torch.empty(n_lists, (seq_len + 31) // 32, dtype=torch.int32, device=labels.device)
    status = torch.zeros(1, dtype=torch.int32, device=labels.device)
    check(lib().rlt_pack_labels(ptr(labels), n_lists, seq_len, ptr(bits), ptr(status), stream_ptr()), "rlt_pack_labels")
    if check_binary and int(status.item()) & 2:
        raise ValueError("labels other than 0. and 1. cannot be stored as bit masks")
    return bits


def reward_matrix(labels, rewards, metric="f1"):
    ensure_tables()
    check(lib().rlt_reward_matrix(ptr(labels), ptr(rewards), labels.shape[0], labels.shape[1],
                                  0 if metric == "f1" else 1, stream_ptr()), "rlt_reward_matrix")


def eval_cut(probs, labels, mode=0, label_bits=None):
    """Returns (k int32[B], count int32[B], n_rel int32[B], f1 float64[B], dcg float64[B]) on the device.
    labels: [B, L] float32 -- or None with label_bits = the words of pack_labels (mode 0, even L: rlt_eval_cut_bits)."""
    ensure_tables()
    n_lists, seq_len = probs.shape[0], probs.shape[1]
    dev = probs.device
    k = torch.empty(n_lists, dtype=torch.int32, device=dev)
    cnt = torch.empty_like(k)
    nrel = torch.empty_like(k)
    f1 = torch.empty(n_lists, dtype=torch.float64, device=dev)
    dcg = torch.empty_like(f1)
    if label_bits is not None:
        if mode != 0 or label_bits.shape != (n_lists, (seq_len + 31) // 32) or label_bits.element_size() != 4:
            raise ValueError("label_bits: mode 0 and [n_lists, ceil(seq_len/32)] 32-bit words")
        check(lib().rlt_eval_cut_bits(ptr(probs), ptr(label_bits), n_lists, seq_len, ptr(k), ptr(cnt), ptr(nrel), ptr(f1),
                                      ptr(dcg), stream_ptr()), "rlt_eval_cut_bits")
        return k, cnt, nrel, f1, dcg
    check(lib().rlt_eval_cut(ptr(probs), ptr(labels), n_lists, seq_len, mode, ptr(k), ptr(cnt), ptr(nrel), ptr(f1),
                             ptr(dcg), stream_ptr()), "rlt_eval_cut")
    return k, cnt, nrel, f1, dcg


def aux_heads_loss(zc, zr, labels, *, n_groups, group_size, seq_len, rerank_softmax=False, class_probs=False, margin=5e-4,
                   class_weight=0.5, rerank_weight=0.5, grad_scale=1.0, loss_scale=1.0, probs_c=None, out_r=None,
                   dzc=None, dzr=None, loss_group=None, status=None, loss_out=None, accumulate=False):
    desc = AuxLossDesc(n_groups, group_size, seq_len, int(rerank_softmax), int(class_probs), int(accumulate), margin, class_weight,
                       rerank_weight, grad_scale, loss_scale)
    check(lib().rlt_aux_heads_loss(C.byref(desc), ptr(zc), ptr(zr), ptr(labels), ptr(probs_c), ptr(out_r), ptr(dzc),
                                   ptr(dzr), ptr(loss_group), ptr(status), ptr(loss_out), stream_ptr()),
          "rlt_aux_heads_loss")


def bicut_loss(u, labels, *, input_kind=0, metric_nci=False, alpha=0.65, r=0.0971134020, probs_out=None, grad=None,
               loss_per_list=None, loss_out=None, grad_scale=1.0, loss_scale=1.0, accumulate=False):
    n_lists, seq_len = labels.shape
    desc = BicutLossDesc(n_lists, seq_len, input_kind, int(metric_nci), int(accumulate), alpha, r, grad_scale,
                         loss_scale)
    check(lib().rlt_bicut_loss(C.byref(desc), ptr(u), ptr(labels), ptr(probs_out), ptr(grad), ptr(loss_per_list),
                               ptr(loss_out), stream_ptr()), "rlt_bicut_loss")


def adam_skip_from_status(status, tensor_ids, tensor_skip):
    """tensor_skip[tensor_ids] = 1 when no group's rerank hinge is active (status bit 1 of rlt_aux_heads_loss), else 0."""
    check(lib().rlt_adam_skip_from_status(ptr(status), int(status.numel()), ptr(tensor_ids), int(tensor_ids.numel()),
                                          ptr(tensor_skip), stream_ptr()), "rlt_adam_skip_from_status")
