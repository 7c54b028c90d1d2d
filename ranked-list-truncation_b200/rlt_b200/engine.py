"""Step engine: drives the C-ABI kernels for whole train / inference steps without autograd.

This is the throughput path (bench.py, data-parallel training): G independent attention groups of S
lists per call, pre-allocated activations, one flat gradient bucket (what a gradient all-reduce
sends), fused logits -> loss -> dlogits (K3) and fused argmax-cut + F1/DCG (K4).  It reads the
parameters of the drop-in nn.Module (models/), so the two paths share weights and kernels; only the
orchestration differs (the nn.Module path goes through torch.autograd so that the reference run.py
can drive it unchanged).

The criterion follows run.py:59-102: Choopy -> ChoopyLoss, AttnCut -> DivLoss(js, tau .85),
Mt* / MMOECut -> MtCutLoss (JS cut loss + rerank_weight * hinge + classi_weight * BCE), BiCut ->
BiCutLoss.  Group losses are averaged, i.e. data-parallel training of the reference with per-replica
batch S.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import ENCODER_PARAM_ORDER

KINDS = ("choopy", "mtchoopy", "bicut", "attncut", "mtattncut", "mmoecut")


def _enc_layers(enc):
    return [[dict(layer.named_parameters())[n] for n in ENCODER_PARAM_ORDER] for layer in enc.layers]


class _EncStack:
    """One nn.TransformerEncoder: per-layer pointer tables, activations and saved buffers."""

    def __init__(self, eng, enc, accumulate_dx=False):
        l0 = enc.layers[0]
        self.d = l0.linear1.in_features
        self.n_head = l0.self_attn.num_heads
        axis = 1 if getattr(eng.model, "attend", "lists") == "positions" else 0     # models._Base.attend
        self.desc = ops.encoder_desc(eng.G, eng.S, eng.L, self.d, l0.self_attn.num_heads, l0.linear1.out_features,
                                     l0.norm1.eps, attend_axis=axis)
        self.desc_infer = ops.encoder_desc(eng.G, eng.S, eng.L, self.d, l0.self_attn.num_heads, l0.linear1.out_features,
                                           l0.norm1.eps, inference=True, attend_axis=axis)
        self.desc_first = ops.encoder_desc(eng.G, eng.S, eng.L, self.d, l0.self_attn.num_heads,
                                           l0.linear1.out_features, l0.norm1.eps, accumulate_dx=accumulate_dx,
                                           attend_axis=axis)
        layers = _enc_layers(enc)
        self.n = len(layers)
        self.w = [ops.encoder_ptrs([p.detach() for p in lw]) for lw in layers]
        self.g = [ops.encoder_ptrs([eng.grad_of(p) for p in lw]) for lw in layers]
        f32 = dict(dtype=torch.float32, device=eng.dev)
        self.outs = [torch.empty(eng.T, self.d, **f32) for _ in range(self.n)]
        sb = (ops.encoder_saved_bytes(self.desc) + 3) // 4
        self.saved = [torch.empty(sb, **f32) for _ in range(self.n if eng.training else 1)]
        self.training = eng.training
        # train-mode dropout: the model's probability while stepping, one fresh mask seed per layer and step
        self.p = float(l0.dropout.p) if eng.training else 0.0
        self.seeds = [0] * self.n

    def _set_drop(self, desc, i):
        desc.dropout_p = self.p
        desc.dropout_seed = self.seeds[i]

    def forward(self, x, train=True):
        from .autograd import fresh_seed
        cur = x
        p_saved = self.p
        if not train:
            self.p = 0.0
        desc = self.desc if train else self.desc_infer      # forward only: nothing is kept, the FFN hidden stays on chip
        for i in range(self.n):
            self.seeds[i] = fresh_seed() if self.p > 0 else 0
            self._set_drop(desc, i)
            ops.encoder_layer_fwd(desc, self.w[i], cur, self.outs[i], self.saved[i if self.training else 0])
            cur = self.outs[i]
        self.p = p_saved
        return cur

    def backward(self, x, d_out, d_x, scratch, ws):
        """d_out: gradient w.r.t. the stack output (clobbered); writes (or accumulates into) d_x."""
        cur = d_out
        for i in reversed(range(self.n)):
            inp = x if i == 0 else self.outs[i - 1]
            if i == 0:
                dst, desc = d_x, self.desc_first
            else:
                dst, desc = (scratch[0] if cur is not scratch[0] else scratch[1]), self.desc
            self._set_drop(desc, i)
            ops.encoder_layer_bwd(desc, self.w[i], self.g[i], inp, self.saved[i], cur, dst, ws)
            cur = dst


class Engine:
    def __init__(self, model, n_groups: int, group_size: int, seq_len: int = 300, metric: str = "f1",
                 rerank_weight: float = 0.5, classi_weight: float = 0.5, training: bool = True):
        self.model = model
        self.kind = type(model).__name__.lower()
        if self.kind not in KINDS:
            raise NotImplementedError(f"Engine: unknown model family {type(model).__name__}")
        # run.py:327 `--num_tasks` is a float: 3 (class + rerank + cut), 2.1 (class + cut), 2.2 (rerank + cut);
        # MtChoopy.py:30-32, MtAttnCut.py:27-29, MMOECut.py:69-84, losses.py:181-191
        self.num_tasks = getattr(model, "num_tasks", 3) if self.kind in ("mtchoopy", "mtattncut", "mmoecut") else 3
        if self.kind == "mmoecut":      # the reference MMOECut keeps no num_tasks attribute: read it off the towers
            self.num_tasks = 3 if len(model.towers) == 3 else (2.1 if model.towers[0].act == "sigmoid" else 2.2)
        if self.num_tasks not in (3, 2.1, 2.2):
            raise NotImplementedError(f"Engine: num_tasks must be 3, 2.1 or 2.2 (got {self.num_tasks})")
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
        f32 = dict(dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            ops.ensure_tables()         # eager: never inside a later CUDA-graph capture

        # ---- parameters and the flat gradient bucket (one all-reduce payload)
        self.named_params = list(model.named_parameters())
        pad = lambda n: (n + 63) // 64 * 64  # noqa: E731  every view stays 256-byte aligned
        self.grad_bucket = torch.zeros(sum(pad(p.numel()) for _, p in self.named_params), **f32)
        self.n_param = sum(p.numel() for _, p in self.named_params)
        self.grads, self._by_id = {}, {}
        off = 0
        for n, p in self.named_params:
            self.grads[n] = self.grad_bucket[off:off + p.numel()].view_as(p)
            self._by_id[id(p)] = n
            off += pad(p.numel())

        k = self.kind
        # ---- front end
        self.lstm = None
        if k in ("choopy", "mtchoopy"):
            self.pe = model.position_encoding
            self.d_front = 128
        else:
            self.lstm = {"bicut": "bilstm", "attncut": "encoding_layer"}.get(k, "pre_encoding")
            mod = getattr(model, self.lstm)
            self.F = mod.input_size
            self.lstm_desc = ops.bilstm_desc(self.B, self.L, self.F, mod.hidden_size, mod.num_layers)
            flat = list(mod._flat_weights)
            self.lstm_w = ops.bilstm_ptrs([p.detach() for p in flat])
            self.lstm_g = ops.bilstm_ptrs([self.grad_of(p) for p in flat])
            sv, ws = ops.bilstm_sizes(self.lstm_desc)
            self.lstm_saved = torch.empty((sv + 3) // 4, **f32) if training else None
            self.lstm_ws = torch.empty((ws + 3) // 4, **f32)
            self.d_front = 2 * mod.hidden_size
        self.front = torch.empty(self.T, self.d_front, **f32)

        # ---- encoder stacks
        if k == "choopy":
            self.stacks = [_EncStack(self, model.attention_layer)]
        elif k == "mtchoopy":
            self.stacks = [_EncStack(self, model.encoding_layer)]
        elif k == "attncut":
            self.stacks = [_EncStack(self, model.attention_layer)]
        elif k == "mtattncut":
            self.stacks = [_EncStack(self, model.encoding_layer)]
        elif k == "mmoecut":
            self.stacks = [_EncStack(self, ex.attention_layer, accumulate_dx=True) for ex in model.experts]
        else:
            self.stacks = []
        self.d = self.stacks[0].d if self.stacks else self.d_front
        self.n_head = self.stacks[0].n_head if self.stacks else 0
        if training:
            if self.stacks:
                self.enc_ws = torch.empty((ops.encoder_workspace_bytes(self.stacks[0].desc) + 3) // 4, **f32)
            self.dact = [torch.empty(self.T, self.d, **f32) for _ in range(3)]

        # ---- heads
        if k in ("choopy", "attncut"):
            self.head_mods = [model.decison_layer[0]]
        elif k in ("mtchoopy", "mtattncut"):
            self.head_mods = {3: [model.classi[0], model.rerank, model.decison_layer[0]],
                              2.1: [model.classi[0], model.decison_layer[0]],
                              2.2: [model.rerank, model.decison_layer[0]]}[self.num_tasks]
        elif k == "mmoecut":
            self.head_mods = [t.linear for t in model.towers]
            self.w_gate_params = list(model.w_gates)
            Tk, E = len(self.w_gate_params), len(model.experts)
            self.moe_desc = ops.MoeDesc(self.B, self.L, self.d_front, self.d, E, Tk)
            self.w_gates = torch.empty(Tk, self.L * self.d_front, E, **f32)
            self.d_w_gates = torch.zeros_like(self.w_gates)
            self.gates = torch.empty(Tk, self.B, E, **f32)
            self.gate_scratch = torch.empty_like(self.gates)
            self.rerank_probs = torch.empty(self.B, self.L, **f32)
            if training:
                self.d_experts = [torch.empty(self.T, self.d, **f32) for _ in range(E)]
        else:  # bicut
            self.head_mods = [model.softmax[1]]
            self.fc_w, self.fc_b = model.fc.weight, model.fc.bias
            self.fc_out = torch.empty(self.T, model.fc.out_features, **f32)
            self.probs2 = torch.empty(self.B, self.L, 2, **f32)
            self.dprobs2 = torch.empty_like(self.probs2)
            if training:
                self.d_fc = torch.empty_like(self.fc_out)
        self.H = 2 if k == "bicut" else len(self.head_mods)
        # which logit rows feed the auxiliary criteria (None = task absent)
        self.i_class = 0 if self.num_tasks in (3, 2.1) else None
        self.i_rerank = {3: 1, 2.1: None, 2.2: 0}[self.num_tasks]
        self.d_head = self.fc_out.shape[1] if k == "bicut" else self.d
        self.head_w = torch.empty(self.H, self.d_head, **f32)
        self.head_b = torch.empty(self.H, **f32)
        self.head_dw = torch.zeros(self.H, self.d_head, **f32)
        self.head_db = torch.zeros(self.H, **f32)
        self.z = torch.empty(self.H, self.B, self.L, **f32)
        self.dz = torch.empty(self.H, self.B, self.L, **f32)
        self.loss_per_list = torch.empty(self.B, **f32)
        self.loss_group = torch.empty(self.G, **f32)
        self.status = torch.zeros(self.G, dtype=torch.int32, device=self.dev)
        self.loss = torch.zeros((), **f32)
        # torch.optim.Adam leaves a parameter without gradient alone; the rerank head has none while its hinge is
        # inactive (losses.py:141).  adam_skip[t] (device, one int32 per parameter in named_params order) is raised for
        # the rerank head's parameters by the criterion when no group's hinge is active; FusedAdam.for_engine reads it.
        self.adam_skip = torch.zeros(len(self.named_params), dtype=torch.int32, device=self.dev)
        self._rerank_param_ids = None
        if self.i_rerank is not None and k in ("mtchoopy", "mtattncut", "mmoecut"):
            index = {id(p): i for i, (_, p) in enumerate(self.named_params)}
            owned = [self.head_mods[self.i_rerank].weight, self.head_mods[self.i_rerank].bias]
            if k == "mmoecut":
                owned.append(self.w_gate_params[self.i_rerank])
            self._rerank_param_ids = torch.tensor([index[id(p)] for p in owned], dtype=torch.int32, device=self.dev)
        self.refresh()

    # ------------------------------------------------------------------------------------------
    def grad_of(self, p):
        return self.grads[self._by_id[id(p)]]

    def refresh(self):
        """Gather the small stacked operands (head weights, MMOECut gates) from the module parameters.  Called at the
        start of every forward, so whoever steps the parameters (FusedAdam, a torch optimizer, load_state_dict) is
        seen by the next step; the copies are a few launches over < 3 MB and are part of the timed step."""
        with torch.no_grad():
            if self.kind == "bicut":
                self.head_w.copy_(self.head_mods[0].weight)
                self.head_b.copy_(self.head_mods[0].bias)
            else:
                for i, m in enumerate(self.head_mods):
                    self.head_w[i].copy_(m.weight[0])
                    self.head_b[i].copy_(m.bias[0])
            if self.kind == "mmoecut":
                for t, p in enumerate(self.w_gate_params):
                    self.w_gates[t].copy_(p)

    # ------------------------------------------------------------------------------------------
    def _forward(self, x, train=False):
        from .autograd import fresh_seed
        k = self.kind
        self.refresh()      # the stacked head / gate operands follow the parameters (an optimizer may have stepped them)
        if self.lstm is None:
            ops.choopy_embed_fwd(x, self.pe.detach(), self.front)
        else:
            ops.bilstm_fwd(self.lstm_desc, self.lstm_w, x, self.front, self.lstm_saved, self.lstm_ws)
        if k == "bicut":
            ops.linear(self.front, self.fc_w.detach(), self.fc_b.detach(), self.fc_out, relu=True)
            ops.head_dots_fwd(self.fc_out, self.head_w, self.head_b, self.z, self.T, self.d_head, 2)
            self.bicut_p = float(self.model._dropout_p) if train else 0.0
            self.bicut_seed = fresh_seed() if self.bicut_p > 0 else 0
            ops.pair_softmax_fwd(self.z, self.probs2, self.T, self.bicut_p, self.bicut_seed)
            return self.fc_out
        tops = [st.forward(self.front, train) for st in self.stacks]
        if k == "mmoecut":
            ops.moe_heads_fwd(self.moe_desc, self.front, self.w_gates, tops, self.head_w, self.head_b, self.gates, self.z)
        else:
            ops.head_dots_fwd(tops[0], self.head_w, self.head_b, self.z, self.T, self.d, self.H)
        self._tops = tops
        return tops[0]

    def _criterion(self, y):
        """Fills self.loss and self.dz (gradient w.r.t. the head logits)."""
        k, B, G = self.kind, self.B, self.G
        cut = self.H - 1
        if k == "bicut":
            ops.bicut_loss(self.probs2, y, input_kind=1, metric_nci=False, grad=self.dprobs2,
                           loss_per_list=self.loss_per_list, loss_out=self.loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
            ops.pair_softmax_bwd(self.probs2, self.dprobs2, self.dz, self.T, self.bicut_p, self.bicut_seed)
            return
        if k == "choopy":
            ops.cut_loss(self.z[cut], y, loss_kind="choopy", metric=self.metric, tau=1.0, input_kind=0, grad=self.dz[cut],
                         loss_per_list=self.loss_per_list, loss_out=self.loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
            return
        ops.cut_loss(self.z[cut], y, loss_kind="js", metric=self.metric, tau=0.85, input_kind=0, grad=self.dz[cut],
                     loss_per_list=self.loss_per_list, loss_out=self.loss, grad_scale=1.0 / B, loss_scale=1.0 / B)
        if k == "attncut":
            return
        ic, ir = self.i_class, self.i_rerank
        ops.aux_heads_loss(None if ic is None else self.z[ic], None if ir is None else self.z[ir], y, n_groups=G,
                           group_size=self.S, seq_len=self.L, rerank_softmax=(k == "mmoecut"),
                           class_weight=self.classi_weight, rerank_weight=self.rerank_weight, grad_scale=1.0 / G,
                           loss_scale=1.0 / G, out_r=self.rerank_probs if (k == "mmoecut" and ir is not None) else None,
                           dzc=None if ic is None else self.dz[ic], dzr=None if ir is None else self.dz[ir],
                           loss_group=self.loss_group, status=self.status, loss_out=self.loss, accumulate=True)
        if self._rerank_param_ids is not None:
            ops.adam_skip_from_status(self.status, self._rerank_param_ids, self.adam_skip)

    def train_step(self, x, y):
        """Forward + criterion + backward.  Gradients land in self.grad_bucket (zeroed first), the scalar
        loss in self.loss (device tensor).  x: [B, L, F], y: [B, L]."""
        if not self.training:
            raise RuntimeError("Engine built with training=False")
        k = self.kind
        self.grad_bucket.zero_()
        self.head_dw.zero_()
        self.head_db.zero_()
        top = self._forward(x, train=True)
        self._criterion(y)
        d_front = self.dact[2]
        if k == "bicut":
            ops.head_dots_bwd(self.fc_out, self.head_w, self.dz, self.d_fc, self.head_dw, self.head_db, self.T,
                              self.d_head, 2, False, True)
            ops.grad_weight(self.d_fc, self.front, self.grad_of(self.fc_w))
            ops.colsum(self.d_fc, self.grad_of(self.fc_b))
            ops.linear_nn(self.d_fc, self.fc_w.detach(), d_front)
            self.grad_of(self.head_mods[0].weight).copy_(self.head_dw)
            self.grad_of(self.head_mods[0].bias).copy_(self.head_db)
        else:
            if k == "mmoecut":
                self.d_w_gates.zero_()
                ops.moe_heads_bwd(self.moe_desc, self.front, self.w_gates, self._tops, self.head_w, self.gates, self.dz,
                                  self.d_experts, self.head_dw, self.head_db, self.d_w_gates, d_front, False,
                                  self.gate_scratch)
                for st, dtop in zip(self.stacks, self.d_experts):
                    st.backward(self.front, dtop, d_front, self.dact[:2], self.enc_ws)   # accumulates into d_front
                for t, p in enumerate(self.w_gate_params):
                    self.grad_of(p).copy_(self.d_w_gates[t])
            else:
                d_top = self.dact[0]
                ops.head_dots_bwd(top, self.head_w, self.dz, d_top, self.head_dw, self.head_db, self.T, self.d, self.H,
                                  False, False)
                self.stacks[0].backward(self.front, d_top, d_front, self.dact[:2], self.enc_ws)
            for i, m in enumerate(self.head_mods):
                self.grad_of(m.weight).copy_(self.head_dw[i:i + 1])
                self.grad_of(m.bias).copy_(self.head_db[i:i + 1])
        if self.lstm is None:
            ops.choopy_embed_bwd(d_front, self.grads["position_encoding"], self.B, self.L)
        else:
            ops.bilstm_bwd(self.lstm_desc, self.lstm_w, self.lstm_g, x, self.lstm_saved, d_front, None, self.lstm_ws)
        return self.loss

    def capture_train_step(self, x, y, warmup: int = 3, optimizer=None, metrics: bool = False):
        """CUDA-graph capture of `train_step` on the static input buffers x, y (the
        reference's batch of 63 lists is launch-bound -- ~250 launches for ~1 ms of device work).  Returns a
        zero-argument callable that replays the whole forward + criterion + backward as ONE graph launch; new batches
        are copied into x / y before the replay, gradients land in `grad_bucket`, the loss in `self.loss`.  Everything
        the step touches is pre-allocated by the Engine and every kernel is launched on the current stream without host
        synchronisation, so the capture needs no special path.  Dropout draws its seeds on the host once per step, which
        a replay would freeze: capture is refused for p > 0.

        optimizer (a `FusedAdam.for_engine(self)`): its step joins the graph -- the per-parameter step counts that
        drive the bias corrections live on the device (rlt_adam_step_masked), nothing of the update is a host scalar
        that changes from step to step.  metrics: the cut positions and per-list F1 / DCG of the batch (run.py:131-145,
        rlt_eval_cut) join it too and land in `self.graph_metrics` = (k, count, n_rel, f1, dcg).  With both, one
        replay is run.py's whole inner loop body for a batch: one launch from the host, no synchronisation."""
        if not self.training:
            raise RuntimeError("Engine built with training=False")
        if float(getattr(self.model, "_dropout_p", 0.0)) > 0.0 or any(float(getattr(st, "p", 0.0)) > 0.0 for st in self.stacks):
            raise RuntimeError("capture_train_step: dropout seeds are drawn on the host every step; capture with dropout = 0")
        if optimizer is not None and getattr(optimizer, "_fixed_grads", None) is None:
            raise RuntimeError("capture_train_step: the optimizer must read the Engine's bucket (FusedAdam.for_engine)")

        def body():
            self.train_step(x, y)
            if metrics:
                if self.kind == "bicut":
                    self.graph_metrics = ops.eval_cut(self.probs2, y, mode=1)
                else:
                    self.graph_metrics = ops.eval_cut(self.z[self.H - 1], y, mode=0)
            if optimizer is not None:
                optimizer.step()

        saved = None
        if optimizer is not None:       # the warm-up steps below must not train: keep the state and put it back
            saved = ([p.detach().clone() for p in optimizer.params], optimizer.state_dict())
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):           # warm-up off the capture: attribute settings, occupancy queries, tables
            for _ in range(max(1, warmup)):
                body()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        if saved is not None:                   # (a capture records, it does not execute)
            with torch.no_grad():
                for p, v in zip(optimizer.params, saved[0]):
                    p.copy_(v)
            optimizer.load_state_dict(saved[1])
        self._graph = graph                     # keeps the captured pool alive
        if optimizer is None:
            return graph.replay

        def replay():
            graph.replay()
            optimizer.step_count += 1           # host mirror of the device-side step counts
        return replay

    def infer(self, x, y):
        """Forward + fused cut selection and per-list F1 / DCG (K4).  Returns (k, f1, dcg) device tensors."""
        self._forward(x)
        if self.kind == "bicut":
            k, _, _, f1, dcg = ops.eval_cut(self.probs2, y, mode=1)
        else:  # argmax of the logits == argmax of their softmax
            k, _, _, f1, dcg = ops.eval_cut(self.z[self.H - 1], y, mode=0)
        return k, f1, dcg

    def apply_grads_to_module(self):
        """Copy the bucket into param.grad (so a torch optimizer can step the drop-in module)."""
        for n, p in self.named_params:
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.grads[n])
