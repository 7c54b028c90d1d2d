"""Step engine: drives the C-ABI kernels for whole train / inference steps without autograd.

This is the throughput path (bench.py, data-parallel training): G independent attention groups of S
lists per call, pre-allocated activations, one flat gradient bucket (what a gradient all-reduce
sends), fused logits -> loss -> dlogits (K3) and fused argmax-cut + F1/DCG (K4).  It reads the
parameters of the drop-in nn.Module (models/), so the two paths share weights and kernels; only the
orchestration differs (the nn.Module path goes through torch.autograd so that the reference run.py
can drive it unchanged).

The criterion follows run.py:59-102: Choopy -> ChoopyLoss, AttnCut -> DivLoss(js, tau .85),
Mt* / MMOECut -> MtCutLoss (JS cut loss + 0.5 rerank hinge + 0.5 BCE), BiCut -> BiCutLoss.  Losses
are averaged over groups, i.e. data-parallel training of the reference with per-replica batch S.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import ENCODER_PARAM_ORDER


def _enc_layers(enc):
    return [[dict(layer.named_parameters())[n] for n in ENCODER_PARAM_ORDER] for layer in enc.layers]


class Engine:
    def __init__(self, model, n_groups: int, group_size: int, seq_len: int = 300, metric: str = "f1",
                 rerank_weight: float = 0.5, classi_weight: float = 0.5, training: bool = True):
        self.model = model
        self.kind = type(model).__name__.lower()
        if self.kind not in ("choopy", "mtchoopy"):
            raise NotImplementedError(f"Engine: model family {type(model).__name__} is not wired yet")
        self.G, self.S, self.L = n_groups, group_size, seq_len
        self.B = n_groups * group_size
        self.T = self.B * seq_len
        self.metric = metric
        self.rerank_weight, self.classi_weight = rerank_weight, classi_weight
        self.training = training
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError("Engine: the model must live on a CUDA device (no CPU path)")
        self.dev = p0.device
        self.timers = None  # set to a dict by bench.py to time selected kernels with CUDA events

        # ---- parameters and the flat gradient bucket (one all-reduce payload)
        self.named_params = [(n, p) for n, p in model.named_parameters()]
        total = sum(p.numel() for _, p in self.named_params)
        pad = lambda n: (n + 63) // 64 * 64  # noqa: E731  keep every view 256-byte aligned
        self.grad_bucket = torch.zeros(sum(pad(p.numel()) for _, p in self.named_params), dtype=torch.float32,
                                       device=self.dev)
        self.n_param = total
        self.grads = {}
        off = 0
        for n, p in self.named_params:
            self.grads[n] = self.grad_bucket[off:off + p.numel()].view_as(p)
            off += pad(p.numel())
        by_id = {id(p): n for n, p in self.named_params}
        gof = lambda p: self.grads[by_id[id(p)]]  # noqa: E731

        # ---- encoder stack
        enc = model.attention_layer if self.kind == "choopy" else model.encoding_layer
        self.d = enc.layers[0].linear1.in_features
        self.n_head = enc.layers[0].self_attn.num_heads
        self.desc = ops.encoder_desc(self.G, self.S, self.L, self.d, self.n_head, enc.layers[0].linear1.out_features,
                                     enc.layers[0].norm1.eps)
        self.layers = _enc_layers(enc)
        self.layer_w = [ops.encoder_ptrs([p.detach() for p in lw]) for lw in self.layers]
        self.layer_g = [ops.encoder_ptrs([gof(p) for p in lw]) for lw in self.layers]
        f32 = dict(dtype=torch.float32, device=self.dev)
        nl = len(self.layers)
        saved_bytes = ops.encoder_saved_bytes(self.desc)
        self.acts = [torch.empty(self.T, self.d, **f32) for _ in range(nl + 1)]      # layer inputs / outputs
        n_saved = nl if training else 1
        self.saved = [torch.empty((saved_bytes + 3) // 4, **f32) for _ in range(n_saved)]
        if training:
            self.ws = torch.empty((ops.encoder_workspace_bytes(self.desc) + 3) // 4, **f32)
            self.dact = [torch.empty(self.T, self.d, **f32) for _ in range(2)]

        # ---- heads
        if self.kind == "choopy":
            self.head_mods = [model.decison_layer[0]]
        else:
            self.head_mods = [model.classi[0], model.rerank, model.decison_layer[0]]
        self.H = len(self.head_mods)
        self.head_w = torch.empty(self.H, self.d, **f32)
        self.head_b = torch.empty(self.H, **f32)
        self.head_dw = torch.zeros(self.H, self.d, **f32)
        self.head_db = torch.zeros(self.H, **f32)
        self.z = torch.empty(self.H, self.B, self.L, **f32)
        self.dz = torch.empty(self.H, self.B, self.L, **f32)
        self.loss_per_list = torch.empty(self.B, **f32)
        self.loss_group = torch.empty(self.G, **f32)
        self.status = torch.zeros(self.G, dtype=torch.int32, device=self.dev)
        self.loss = torch.zeros((), **f32)
        self.pe = model.position_encoding
        self.refresh_heads()

    # ------------------------------------------------------------------------------------------
    def refresh_heads(self):
        """Gather the Linear(d,1) head parameters into one [H, d] operand (call after an optimizer step)."""
        with torch.no_grad():
            for i, m in enumerate(self.head_mods):
                self.head_w[i].copy_(m.weight[0])
                self.head_b[i].copy_(m.bias[0])

    def _mark(self, name, start: bool):
        if self.timers is not None and name in self.timers:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timers[name].append(ev)

    def _forward(self, x):
        """x: [B, L, 1] scores.  Returns the final hidden states [T, d]."""
        ops.choopy_embed_fwd(x, self.pe.detach(), self.acts[0])
        for i in range(len(self.layers)):
            sv = self.saved[i if self.training else 0]
            self._mark("encoder_fwd", True)
            ops.encoder_layer_fwd(self.desc, self.layer_w[i], self.acts[i], self.acts[i + 1], sv)
            self._mark("encoder_fwd", False)
        h = self.acts[-1]
        ops.head_dots_fwd(h, self.head_w, self.head_b, self.z, self.T, self.d, self.H)
        return h

    def train_step(self, x, y):
        """Forward + criterion + backward.  Gradients land in self.grad_bucket (zeroed first), the scalar
        loss in self.loss (device).  x: [B, L, F], y: [B, L]."""
        if not self.training:
            raise RuntimeError("Engine built with training=False")
        self.grad_bucket.zero_()
        self.head_dw.zero_()
        self.head_db.zero_()
        h = self._forward(x)
        B, G = self.B, self.G
        cut = self.H - 1
        if self.kind == "choopy":
            ops.cut_loss(self.z[cut], y, loss_kind="choopy", metric=self.metric, tau=1.0, input_kind=0, grad=self.dz[cut],
                         loss_per_list=self.loss_per_list, loss_out=self.loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
        else:
            ops.cut_loss(self.z[cut], y, loss_kind="js", metric=self.metric, tau=0.85, input_kind=0, grad=self.dz[cut],
                         loss_per_list=self.loss_per_list, loss_out=self.loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
            ops.aux_heads_loss(self.z[0], self.z[1], y, n_groups=G, group_size=self.S, seq_len=self.L,
                               rerank_softmax=False, class_weight=self.classi_weight,
                               rerank_weight=self.rerank_weight, grad_scale=1.0 / G, loss_scale=1.0 / G,
                               dzc=self.dz[0], dzr=self.dz[1], loss_group=self.loss_group, status=self.status,
                               loss_out=self.loss, accumulate=True)
        d_h = self.dact[0]
        ops.head_dots_bwd(h, self.head_w, self.dz, d_h, self.head_dw, self.head_db, self.T, self.d, self.H, False)
        cur, other = self.dact[0], self.dact[1]
        for i in reversed(range(len(self.layers))):
            self._mark("encoder_bwd", True)
            ops.encoder_layer_bwd(self.desc, self.layer_w[i], self.layer_g[i], self.acts[i], self.saved[i], cur, other,
                                  self.ws)
            self._mark("encoder_bwd", False)
            cur, other = other, cur
        ops.choopy_embed_bwd(cur, self.grads["position_encoding"], self.B, self.L)
        # scatter the stacked head gradients back to the per-module views of the bucket
        for i, m in enumerate(self.head_mods):
            self.grads[self._name_of(m.weight)].copy_(self.head_dw[i:i + 1])
            self.grads[self._name_of(m.bias)].copy_(self.head_db[i:i + 1])
        return self.loss

    def _name_of(self, p):
        for n, q in self.named_params:
            if q is p:
                return n
        raise KeyError

    def infer(self, x, y):
        """Forward + fused cut selection and per-list F1 / DCG (K4).  Returns (k, f1, dcg) device tensors."""
        self._forward(x)
        k, _, _, f1, dcg = ops.eval_cut(self.z[self.H - 1], y, mode=0)   # argmax of logits == argmax of softmax
        return k, f1, dcg

    def apply_grads_to_module(self):
        """Copy the bucket into param.grad (so a torch optimizer can step the drop-in module)."""
        for n, p in self.named_params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.grads[n])
