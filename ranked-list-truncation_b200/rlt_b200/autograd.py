"""torch.autograd.Function wrappers that put the C-ABI kernels behind the reference's nn.Module API.

Each Function's forward/backward is a handful of ctypes calls into librlt_b200.so; torch supplies
device memory, the stream and the autograd graph.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import torch

from . import ops


def _buf(nbytes: int, device) -> torch.Tensor:
    return torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)


_seed_gen = None
_seed_gen_key = None


def fresh_seed() -> int:
    """Seed of one forward call's dropout masks.  Drawn from a PRIVATE host generator seeded from
    `torch.initial_seed()` (so `torch.manual_seed()` still makes runs repeatable) plus the data-parallel rank (so
    replicas draw different masks).  torch's global CPU generator is left alone: the reference's dropout runs on the
    CUDA Philox stream and never consumes it, so a DataLoader / `DeviceLoader.shuffled_order` under the same
    `torch.manual_seed` keeps producing the reference's batches in every epoch.  The same seed is handed to the
    backward kernels."""
    global _seed_gen, _seed_gen_key
    rank = 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank()
    key = (torch.initial_seed(), rank)
    if _seed_gen is None or _seed_gen_key != key:
        _seed_gen = torch.Generator()
        _seed_gen.manual_seed((key[0] * 1000003 + 7919 * rank + 12345) % (2 ** 63 - 1))
        _seed_gen_key = key
    return int(torch.randint(0, 2 ** 62, (1,), generator=_seed_gen).item())


def reset_dropout_seed() -> None:
    """Restart the private dropout-seed stream from `torch.initial_seed()` (it restarts by itself whenever
    `torch.manual_seed` is given a NEW value; call this to replay a run under the same seed)."""
    global _seed_gen_key
    _seed_gen_key = None


def _c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError("rlt_b200 computes in float32 only (the reference's parameters and inputs are float32); "
                           f"got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError("rlt_b200 has no CPU path: move the model and its inputs to a CUDA device "
                           "(the reference run.py does this when torch.cuda.is_available())")
    return t.contiguous()


class EncoderStack(torch.autograd.Function):
    """nn.TransformerEncoder (a stack of post-norm layers) on [B, L, d]; the B lists of the call form ONE
    attention group, as in the reference (no batch_first).  `EncoderStackWithin` is the same stack with attention
    WITHIN each list (attend_axis = 1: `enc(x.transpose(0, 1)).transpose(0, 1)` on the reference module)."""
    ATTEND_AXIS = 0

    @staticmethod
    def forward(ctx, x, n_head, n_groups, ln_eps, dropout_p, *params):
        return EncoderStack._forward(ctx, 0, x, n_head, n_groups, ln_eps, dropout_p, *params)

    @staticmethod
    def _forward(ctx, attend_axis, x, n_head, n_groups, ln_eps, dropout_p, *params):
        x = _c(x)
        B, L, d = x.shape
        n_layers = len(params) // 12
        params = [_c(p.detach()) for p in params]
        d_ff = params[4].shape[0]
        if B % n_groups:
            raise RuntimeError(f"batch of {B} lists is not divisible into {n_groups} attention groups")
        dropout_p = float(dropout_p)
        need_grad = any(ctx.needs_input_grad)
        descs = [ops.encoder_desc(n_groups, B // n_groups, L, d, n_head, d_ff, ln_eps, dropout_p,
                                  fresh_seed() if dropout_p > 0 else 0, inference=not need_grad,
                                  attend_axis=attend_axis)
                 for _ in range(n_layers)]   # one mask set per layer
        desc = descs[0]
        saved_bytes = ops.encoder_saved_bytes(desc)
        if saved_bytes == 0:
            raise ops._lib.RltError("encoder: " + ops._lib.load().rlt_last_error().decode())
        cur = x
        saved, inputs = [], []
        scratch = None
        for i in range(n_layers):
            w = ops.encoder_ptrs(params[12 * i:12 * i + 12])
            out = torch.empty_like(x)
            if need_grad:
                sv = _buf(saved_bytes, x.device)
            else:  # inference: one scratch "saved" buffer reused by every layer
                if scratch is None:
                    scratch = _buf(saved_bytes, x.device)
                sv = scratch
            ops.encoder_layer_fwd(descs[i], w, cur, out, sv)
            if need_grad:
                saved.append(sv)
                inputs.append(cur)
            cur = out
        ctx.desc = desc
        ctx.descs = descs
        ctx.n_layers = n_layers
        ctx.params = params
        ctx.saved_bufs = saved
        ctx.layer_inputs = inputs
        return cur

    @staticmethod
    def backward(ctx, d_out):
        d_out = _c(d_out)
        desc, params = ctx.desc, ctx.params
        ws = _buf(ops.encoder_workspace_bytes(desc), d_out.device)
        grads = [torch.zeros_like(p) for p in params]
        cur = d_out
        for i in reversed(range(ctx.n_layers)):
            w = ops.encoder_ptrs(params[12 * i:12 * i + 12])
            g = ops.encoder_ptrs(grads[12 * i:12 * i + 12])
            d_x = torch.empty_like(d_out)
            ops.encoder_layer_bwd(ctx.descs[i], w, g, ctx.layer_inputs[i], ctx.saved_bufs[i], cur, d_x, ws)
            cur = d_x
        ctx.saved_bufs = ctx.layer_inputs = None
        return (cur, None, None, None, None, *grads)


class EncoderStackWithin(EncoderStack):
    """The encoder stack with attention within each list (the L positions of a list attend to each other)."""
    ATTEND_AXIS = 1

    @staticmethod
    def forward(ctx, x, n_head, n_groups, ln_eps, dropout_p, *params):
        return EncoderStack._forward(ctx, 1, x, n_head, n_groups, ln_eps, dropout_p, *params)


class BiLstm(torch.autograd.Function):
    """nn.LSTM(F, 128, num_layers=2, batch_first=True, bidirectional=True)(x)[0] on [B, L, F] -> [B, L, 256]."""

    @staticmethod
    def forward(ctx, x, hidden, num_layers, *flat):
        x = _c(x)
        B, L, Fin = x.shape
        flat = [_c(p.detach()) for p in flat]
        desc = ops.bilstm_desc(B, L, Fin, hidden, num_layers)
        sv_bytes, ws_bytes = ops.bilstm_sizes(desc)
        need_grad = any(ctx.needs_input_grad)
        y = torch.empty(B, L, 2 * hidden, dtype=torch.float32, device=x.device)
        saved = _buf(sv_bytes, x.device) if need_grad else None
        ws = _buf(ws_bytes, x.device)
        ops.bilstm_fwd(desc, ops.bilstm_ptrs(flat), x, y, saved, ws)
        ctx.desc, ctx.flat, ctx.saved_buf, ctx.x, ctx.ws = desc, flat, saved, x, ws
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        grads = [torch.zeros_like(p) for p in ctx.flat]
        dx = torch.empty_like(ctx.x) if ctx.needs_input_grad[0] else None
        ops.bilstm_bwd(ctx.desc, ops.bilstm_ptrs(ctx.flat), ops.bilstm_ptrs(grads), ctx.x, ctx.saved_buf, dy, dx, ctx.ws)
        ctx.saved_buf = ctx.ws = None
        return (dx, None, None, *grads)


class BicutHead(torch.autograd.Function):
    """softmax_2( dropout_p( Linear(256,2)( relu( Linear(256,256)(h) ) ) ) ) of models/Bicut.py:10-16,20-21."""

    @staticmethod
    def forward(ctx, h, w1, b1, w2, b2, dropout_p=0.0):
        h, w1, b1, w2, b2 = _c(h), _c(w1.detach()), _c(b1.detach()), _c(w2.detach()), _c(b2.detach())
        B, L, d = h.shape
        T = B * L
        a = torch.empty(T, w1.shape[0], dtype=torch.float32, device=h.device)
        ops.linear(h.view(T, d), w1, b1, a, relu=True)
        z = torch.empty(2, T, dtype=torch.float32, device=h.device)
        ops.head_dots_fwd(a, w2, b2, z, T, w1.shape[0], 2)
        o = torch.empty(B, L, 2, dtype=torch.float32, device=h.device)
        ctx.dropout_p = float(dropout_p)
        ctx.seed = fresh_seed() if ctx.dropout_p > 0 else 0
        ops.pair_softmax_fwd(z, o, T, ctx.dropout_p, ctx.seed)
        ctx.save_for_backward(h, w1, w2, a, o)
        return o

    @staticmethod
    def backward(ctx, d_o):
        h, w1, w2, a, o = ctx.saved_tensors
        d_o = _c(d_o)
        B, L, d = h.shape
        T = B * L
        f = w1.shape[0]
        dz = torch.empty(2, T, dtype=torch.float32, device=h.device)
        ops.pair_softmax_bwd(o, d_o, dz, T, ctx.dropout_p, ctx.seed)
        da = torch.empty_like(a)
        dw2 = torch.zeros_like(w2)
        db2 = torch.zeros(2, dtype=torch.float32, device=h.device)
        ops.head_dots_bwd(a, w2, dz, da, dw2, db2, T, f, 2, False, True)      # masked by the ReLU
        dw1 = torch.zeros_like(w1)
        db1 = torch.zeros(f, dtype=torch.float32, device=h.device)
        ops.grad_weight(da, h.view(T, d), dw1)
        ops.colsum(da, db1)
        dh = torch.empty_like(h)
        ops.linear_nn(da, w1, dh.view(T, d))
        return dh, dw1, db1, dw2, db2, None


class MoeGateMix(torch.autograd.Function):
    """MMOECut gates, mixtures and tower Linear(d,1) layers (models/MMOECut.py:90-105).  Called as
    `MoeGateMix.apply(h_lstm, n_towers, *gates, *tower_w, *tower_b, *experts)` (n_towers gate matrices [L*dl, E], tower
    weights [1, d] and biases [1], then the E expert outputs); returns one logit tensor [B, L] per tower.

    As for HeadDots, per-tower inputs / outputs preserve "no gradient": the gate and tower parameters of a tower whose
    logits received no gradient (rerank tower, hinge inactive) get None, as in the reference."""

    @staticmethod
    def forward(ctx, h_lstm, n_towers, *rest):
        Tk = int(n_towers)
        gates_in, tw_in, tb_in, experts = rest[:Tk], rest[Tk:2 * Tk], rest[2 * Tk:3 * Tk], rest[3 * Tk:]
        h_lstm = _c(h_lstm)
        w_gates = torch.stack([_c(g.detach()) for g in gates_in])
        tower_w = torch.cat([_c(t.detach()).reshape(1, -1) for t in tw_in], dim=0)
        tower_b = torch.cat([_c(t.detach()).reshape(1) for t in tb_in], dim=0)
        experts = [_c(e) for e in experts]
        B, L, dl = h_lstm.shape
        _, J, E = w_gates.shape
        d = experts[0].shape[2]
        if J != L * dl:
            raise RuntimeError(f"w_gates expects {J} gate inputs but the LSTM output has {L}x{dl}")
        if E != len(experts):
            raise RuntimeError(f"gates of width {E} over {len(experts)} experts")
        desc = ops.MoeDesc(B, L, dl, d, E, Tk)
        gates = torch.empty(Tk, B, E, dtype=torch.float32, device=h_lstm.device)
        z = torch.empty(Tk, B, L, dtype=torch.float32, device=h_lstm.device)
        ops.moe_heads_fwd(desc, h_lstm, w_gates, experts, tower_w, tower_b, gates, z)
        ctx.desc = desc
        ctx.shapes = ([t.shape for t in tw_in], [t.shape for t in tb_in])
        ctx.save_for_backward(h_lstm, w_gates, tower_w, gates, *experts)
        ctx.set_materialize_grads(False)
        return tuple(z[t] for t in range(Tk))

    @staticmethod
    def backward(ctx, *dzs):
        h_lstm, w_gates, tower_w, gates, *experts = ctx.saved_tensors
        Tk = w_gates.shape[0]
        if all(g is None for g in dzs):
            return (None,) * (2 + 3 * Tk + len(experts))
        B, L, _ = h_lstm.shape
        dz = torch.stack([torch.zeros(B, L, dtype=torch.float32, device=h_lstm.device) if g is None else _c(g) for g in dzs])
        d_experts = [torch.empty_like(e) for e in experts]
        d_tw = torch.zeros_like(tower_w)
        d_tb = torch.zeros(Tk, dtype=torch.float32, device=dz.device)
        d_wg = torch.zeros_like(w_gates)
        d_h = torch.empty_like(h_lstm)
        scratch = torch.empty_like(gates)
        ops.moe_heads_bwd(ctx.desc, h_lstm, w_gates, experts, tower_w, gates, dz, d_experts, d_tw, d_tb, d_wg, d_h, False,
                          scratch)
        live = [g is not None for g in dzs]
        tw_shapes, tb_shapes = ctx.shapes
        return (d_h, None,
                *[d_wg[t] if live[t] else None for t in range(Tk)],
                *[d_tw[t].reshape(tw_shapes[t]) if live[t] else None for t in range(Tk)],
                *[d_tb[t].reshape(tb_shapes[t]) if live[t] else None for t in range(Tk)],
                *d_experts)


class ChoopyEmbed(torch.autograd.Function):
    """x = cat(score, position_encoding.expand(B, L, 127)) (models/Choopy.py:19-20)."""

    @staticmethod
    def forward(ctx, score, pe):
        score, pe = _c(score), _c(pe.detach())
        B, L = score.shape[0], score.shape[1]
        if score.shape[2:] != (1,) or pe.shape != (L, 127):
            raise RuntimeError(f"Choopy expects scores [B, {pe.shape[0]}, 1]; got {tuple(score.shape)}")
        x = torch.empty(B, L, 128, dtype=torch.float32, device=score.device)
        ops.choopy_embed_fwd(score, pe, x)
        ctx.shape = (B, L)
        return x

    @staticmethod
    def backward(ctx, dx):
        dx = _c(dx)
        B, L = ctx.shape
        dpe = torch.zeros(L, 127, dtype=torch.float32, device=dx.device)
        ops.choopy_embed_bwd(dx, dpe, B, L)
        dscore = dx[:, :, 0:1].contiguous() if ctx.needs_input_grad[0] else None
        return dscore, dpe


class HeadDots(torch.autograd.Function):
    """H parallel Linear(d, 1) heads evaluated in ONE pass over h.  Called as `HeadDots.apply(h, w0, b0, w1, b1, ...)`
    with the heads' own parameters; returns H logit tensors [B, L] (one per head, views of one buffer).

    Separate inputs / outputs per head keep autograd's "no gradient" information: a head whose output received no
    gradient (the rerank head while its hinge is inactive: the reference's criterion returns a constant there,
    utils/losses.py:141) gets `None` for its weight and bias, exactly as in the reference, so that an optimizer leaves
    it alone for that step."""

    @staticmethod
    def forward(ctx, h, *wb):
        h = _c(h)
        H = len(wb) // 2
        w = torch.cat([_c(wb[2 * i].detach()).reshape(1, -1) for i in range(H)], dim=0)
        b = torch.cat([_c(wb[2 * i + 1].detach()).reshape(1) for i in range(H)], dim=0)
        B, L, d = h.shape
        z = torch.empty(H, B, L, dtype=torch.float32, device=h.device)
        ops.head_dots_fwd(h, w, b, z, B * L, d, H)
        ctx.save_for_backward(h, w)
        ctx.set_materialize_grads(False)
        ctx.wb_shapes = [t.shape for t in wb]
        return tuple(z[i] for i in range(H))

    @staticmethod
    def backward(ctx, *dzs):
        h, w = ctx.saved_tensors
        B, L, d = h.shape
        H = w.shape[0]
        if all(g is None for g in dzs):
            return (None,) * (1 + 2 * H)
        dz = torch.stack([torch.zeros(B, L, dtype=torch.float32, device=h.device) if g is None else _c(g) for g in dzs])
        dx = torch.empty_like(h)
        dw = torch.zeros_like(w)
        db = torch.zeros(H, dtype=torch.float32, device=h.device)
        ops.head_dots_bwd(h, w, dz, dx, dw, db, B * L, d, H, False)
        out = [dx]
        for i, g in enumerate(dzs):
            if g is None:
                out += [None, None]
            else:
                out += [dw[i].reshape(ctx.wb_shapes[2 * i]), db[i].reshape(ctx.wb_shapes[2 * i + 1])]
        return tuple(out)


class SoftmaxLists(torch.autograd.Function):
    """nn.Softmax(dim=1) over the L positions of logits [B, L]."""

    @staticmethod
    def forward(ctx, z):
        z = _c(z)
        p = torch.empty_like(z)
        ops.softmax_lists(z, p, z.shape[0], z.shape[1])
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        dz = torch.empty_like(p)
        ops.softmax_lists_bwd(p, _c(dp), dz, p.shape[0], p.shape[1])
        return dz


class CutLoss(torch.autograd.Function):
    """ChoopyLoss / AttnCutLoss / DivLoss on probabilities p [B, L] (the reference criteria take the model
    output, i.e. probabilities): one fused kernel produces the loss and dL/dp."""

    @staticmethod
    def forward(ctx, p, labels, loss_kind, metric, tau):
        p, labels = _c(p), _c(labels)
        B, L = labels.shape
        grad = torch.empty_like(p)
        per_list = torch.empty(B, dtype=torch.float32, device=p.device)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        ops.cut_loss(p, labels, loss_kind=loss_kind, metric=metric, tau=tau, input_kind=1, grad=grad,
                     loss_per_list=per_list, loss_out=loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


class AuxHeadsLoss(torch.autograd.Function):
    """classi_weight * BCELoss(class_p) + rerank_weight * RerankLoss(rerank_out) of MtCutLoss on the
    reference API boundary: class_p are probabilities, rerank_out the rerank head output."""

    @staticmethod
    def forward(ctx, class_p, rerank_out, labels, class_weight, rerank_weight, margin):
        labels = _c(labels)
        B, L = labels.shape
        dev = labels.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        loss_group = torch.empty(1, dtype=torch.float32, device=dev)
        pc = dpc = zr = dzr = None
        if class_p is not None:
            pc = _c(class_p).reshape(B, L)
            dpc = torch.empty_like(pc)
        if rerank_out is not None:
            zr = _c(rerank_out).reshape(B, L)
            dzr = torch.empty_like(zr)
        ops.aux_heads_loss(pc, zr, labels, n_groups=1, group_size=B, seq_len=L, rerank_softmax=False, class_probs=True,
                           margin=margin, class_weight=class_weight, rerank_weight=rerank_weight, dzc=dpc, dzr=dzr,
                           loss_group=loss_group, status=status, loss_out=loss)
        hinge_active = False
        if rerank_out is not None:
            st = int(status.item())     # the reference syncs here too (two .item() calls, utils/losses.py:136-137)
            if st & 1:
                # reference utils/losses.py:138 returns torch.tensor(0, requires_grad=True) -> RuntimeError
                raise RuntimeError("Only Tensors of floating point and complex dtype can require gradients")
            hinge_active = bool(st & 2)
        # hinge inactive: the reference's RerankLoss returns a constant zero leaf (losses.py:141) -- the rerank output
        # gets NO gradient (None), not a zero one, so the heads behind it keep `.grad is None`
        grads = [dpc.reshape(class_p.shape) if class_p is not None else None,
                 dzr.reshape(rerank_out.shape) if (rerank_out is not None and hinge_active) else None]
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, g):
        gc, gr = ctx.grads
        return (gc * g if gc is not None else None, gr * g if gr is not None else None, None, None, None, None)


class BicutLoss(torch.autograd.Function):
    """BiCutLoss on the model output [B, L, 2] (probabilities)."""

    @staticmethod
    def forward(ctx, out, labels, alpha, r, metric_nci):
        out, labels = _c(out), _c(labels)
        B, L = labels.shape
        grad = torch.empty_like(out)
        per_list = torch.empty(B, dtype=torch.float32, device=out.device)
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        ops.bicut_loss(out, labels, input_kind=1, metric_nci=metric_nci, alpha=alpha, r=r, grad=grad,
                       loss_per_list=per_list, loss_out=loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None
