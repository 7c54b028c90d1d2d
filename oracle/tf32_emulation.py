"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): emulate tensor-core operand rounding inside the oracle.

SURVEY.md section 8(c) derives the 1e-3 parity tolerance from an emulation of TF32 operands with wide accumulation
("operands rounded, fp32 accumulate") inside the restated reference forward/backward.  This module reproduces that
experiment for any oracle forward so that a tolerance written in a GPU parity test can be justified by a number
computed HERE, on the CPU, from the same fixture: inside `with tf32_matmuls():` every `torch.matmul` / `@` of the
oracle rounds BOTH operands to TF32 (10 explicit mantissa bits, round to nearest even) before a float64 product, in
the forward AND in the two gradient products of the backward (dA = r(dC) r(B)^T, dB = r(A)^T r(dC)).  Everything
else (softmax, LayerNorm, gates, losses) stays in float64 -- i.e. the error measured is the floor any TF32 tensor-core
implementation with exact accumulation has; the CUDA path adds fp32 accumulation order on top.
"""
from __future__ import annotations

import torch
from torch.overrides import TorchFunctionMode


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """float -> nearest TF32 value (ties to even), returned in float64."""
    b = x.detach().to(torch.float32).contiguous().view(torch.int32)
    b = (b + 0x0FFF + ((b >> 13) & 1)) & ~0x1FFF
    return b.view(torch.float32).to(torch.float64)


class _RoundedMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ra, rb = round_tf32(a), round_tf32(b)
        ctx.save_for_backward(ra, rb)
        ctx.shapes = (a.shape, b.shape, a.dtype, b.dtype)
        return torch.matmul(ra, rb).to(a.dtype)

    @staticmethod
    def backward(ctx, g):
        ra, rb = ctx.saved_tensors
        sa, sb, da, db = ctx.shapes
        rg = round_tf32(g)
        ga = torch.matmul(rg, rb.transpose(-1, -2))
        gb = torch.matmul(ra.transpose(-1, -2), rg)
        while ga.dim() > len(sa):
            ga = ga.sum(0)
        while gb.dim() > len(sb):
            gb = gb.sum(0)
        for i, n in enumerate(sa):      # broadcast batch dims
            if ga.shape[i] != n:
                ga = ga.sum(i, keepdim=True)
        for i, n in enumerate(sb):
            if gb.shape[i] != n:
                gb = gb.sum(i, keepdim=True)
        return ga.to(da), gb.to(db)


class tf32_matmuls(TorchFunctionMode):
    """Context manager: route every matmul issued inside through `_RoundedMatmul`.  `skip(a, b)` may exempt products
    the CUDA path evaluates in exact fp32 (e.g. the MMOECut gate GEMV)."""

    def __init__(self, skip=None):
        super().__init__()
        self.skip = skip

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__) and len(args) == 2 and not kwargs:
            a, b = args
            if a.dim() >= 2 and b.dim() >= 2 and not (self.skip and self.skip(a, b)):
                with torch._C.DisableTorchFunctionSubclass():
                    return _RoundedMatmul.apply(a, b)
        return func(*args, **kwargs)
