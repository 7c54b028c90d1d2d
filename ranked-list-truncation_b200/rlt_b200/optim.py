"""Fused Adam for the truncation models: the reference's optimizer step (`run.py:104`,
`optim.Adam(self.model.parameters(), lr=args.lr, weight_decay=self.weight_decay)`, stepped once per batch at
`run.py:129`) as ONE kernel launch over every parameter tensor.

`FusedAdam` keeps torch.optim.Adam's constructor arguments and `step()` / `zero_grad()` / `state_dict()` surface for
the options the reference uses (L2 weight decay folded into the gradient, `amsgrad=False`, `maximize=False`).  The
parameters stay where torch allocated them; gradients are read from `param.grad` (drop-in nn.Module path) or from an
`Engine`'s flat, all-reduced gradient bucket; the two moment buffers are flat.  There is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, ptr, stream_ptr
from .ops import lib


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, *,
                 maximize=False, grads=None, grad_scale=1.0, skip=None):
        """params: iterable of CUDA fp32 parameters (e.g. `model.parameters()`).  grads (optional): one gradient
        tensor per parameter (views of an Engine's bucket: `FusedAdam.for_engine`); default: `param.grad` at step
        time.  grad_scale: factor applied to the gradient first (1/world after a SUM all-reduce)."""
        if amsgrad or maximize:
            raise NotImplementedError("FusedAdam: amsgrad / maximize are not used by the reference (run.py:104)")
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedAdam: invalid hyper-parameter")      # torch.optim.Adam raises ValueError too
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("FusedAdam: parameters must be contiguous float32 CUDA tensors (no CPU path)")
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.grad_scale = float(grad_scale)
        self.step_count = 0
        self._fixed_grads = list(grads) if grads is not None else None
        dev = self.params[0].device
        pad = lambda n: (n + 63) // 64 * 64  # noqa: E731
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += pad(p.numel())
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._state_offset = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._offsets = offs
        chunk = int(lib().rlt_adam_chunk_elems())
        ct, cf, cl = [], [], []
        for t, p in enumerate(self.params):
            for first in range(0, p.numel(), chunk):
                ct.append(t)
                cf.append(first)
                cl.append(min(chunk, p.numel() - first))
        i32 = dict(dtype=torch.int32, device=dev)
        self._chunk_tensor, self._chunk_first, self._chunk_len = (torch.tensor(a, **i32) for a in (ct, cf, cl))
        self._n_chunks = len(ct)
        self._param_ptrs = torch.tensor([p.data_ptr() for p in self.params], dtype=torch.int64, device=dev)
        self._grad_key, self._grad_ptrs = None, None
        self.tensor_steps = torch.zeros(len(self.params), **i32)       # torch: state[p]['step'], per parameter
        self._ext_skip = skip
        self._skip_key, self._skip = None, None

    @classmethod
    def for_engine(cls, engine, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        """Step the Engine's parameters straight from its flat gradient bucket (after the all-reduce)."""
        params = [p for _, p in engine.named_params]
        grads = [engine.grads[n] for n, _ in engine.named_params]
        return cls(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grads=grads, grad_scale=grad_scale,
                   skip=engine.adam_skip)

    def _grads(self):
        """(device table of gradient addresses, device skip flags or None)."""
        gs = self._fixed_grads if self._fixed_grads is not None else [p.grad for p in self.params]
        for g, p in zip(gs, self.params):
            if g is not None and not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == p.numel()):
                raise RuntimeError("FusedAdam.step: gradients must be contiguous float32 CUDA tensors")
        dev = self.params[0].device
        key = tuple(0 if g is None else g.data_ptr() for g in gs)
        if key != self._grad_key:       # autograd may re-allocate .grad after zero_grad(set_to_none=True)
            self._grad_ptrs = torch.tensor(key, dtype=torch.int64, device=dev)
            self._grad_key = key
        if self._ext_skip is not None:
            return self._grad_ptrs, self._ext_skip
        skip_key = tuple(g is None for g in gs)
        if not any(skip_key):
            return self._grad_ptrs, None
        if skip_key != self._skip_key:
            self._skip = torch.tensor([int(v) for v in skip_key], dtype=torch.int32, device=dev)
            self._skip_key = skip_key
        return self._grad_ptrs, self._skip

    @torch.no_grad()
    def step(self):
        gp, skip = self._grads()
        self.step_count += 1
        d = self.defaults
        check(lib().rlt_adam_step_masked(ptr(self._param_ptrs), ptr(gp), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                         ptr(self._state_offset), ptr(self._chunk_tensor), ptr(self._chunk_first),
                                         ptr(self._chunk_len), C.c_int(self._n_chunks), C.c_int(len(self.params)),
                                         ptr(self.tensor_steps), ptr(skip), C.c_double(d["lr"]),
                                         C.c_double(d["betas"][0]), C.c_double(d["betas"][1]), C.c_double(d["eps"]),
                                         C.c_double(d["weight_decay"]), C.c_double(self.grad_scale), stream_ptr()),
              "rlt_adam_step_masked")

    def zero_grad(self, set_to_none=True):
        if self._fixed_grads is not None:
            return      # the Engine zeroes its bucket at the start of train_step
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def moments(self, i):
        """(exp_avg, exp_avg_sq) views of parameter i (same shapes as torch.optim.Adam's state tensors)."""
        p, o = self.params[i], self._offsets[i]
        return self.exp_avg[o:o + p.numel()].view_as(p), self.exp_avg_sq[o:o + p.numel()].view_as(p)

    def state_dict(self):
        return {"step": self.step_count, "tensor_steps": self.tensor_steps.clone(), "defaults": dict(self.defaults),
                "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        if "tensor_steps" in sd:
            self.tensor_steps.copy_(sd["tensor_steps"])
        else:
            self.tensor_steps.fill_(self.step_count)
        self.defaults.update(sd["defaults"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
