"""Seeded synthetic robust04-shaped ranked lists.

score_j : sorted-descending( s0 * (1 - 0.67 j/L) + 0.3 N(0,1) ), s0 = 9   (DRMM-TKS-like range)
extra features : U(0,1)   (the two neighbour-similarity features of the AttnCut data)
label_j ~ Bernoulli(a * exp(-j/tau)), a ~ U(0.2, 0.9), tau ~ U(10, 80) per list, >= 1 relevant forced
Returns X [n, L, F] float32 and y [n, L] float32 in {0,1}, the shapes the reference loaders yield
(dataloader/attncut_dataloader.py:21-59, choopy_dataloader.py:21-45).
"""
from __future__ import annotations

import torch


def synthetic_lists(n_lists: int, seq_len: int = 300, n_features: int = 3, seed: int = 20240229,
                    device: str | torch.device = "cpu"):
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    j = torch.arange(seq_len, device=dev, dtype=torch.float32)
    base = 9.0 * (1.0 - 0.67 * j / seq_len)
    score = base.unsqueeze(0) + 0.3 * torch.randn(n_lists, seq_len, generator=g, device=dev)
    score, _ = torch.sort(score, dim=1, descending=True)
    feats = [score.unsqueeze(2)]
    if n_features > 1:
        feats.append(torch.rand(n_lists, seq_len, n_features - 1, generator=g, device=dev))
    x = torch.cat(feats, dim=2).contiguous()
    a = 0.2 + 0.7 * torch.rand(n_lists, 1, generator=g, device=dev)
    tau = 10.0 + 70.0 * torch.rand(n_lists, 1, generator=g, device=dev)
    prob = a * torch.exp(-j.unsqueeze(0) / tau)
    y = (torch.rand(n_lists, seq_len, generator=g, device=dev) < prob).float()
    empty = y.sum(dim=1) == 0
    y[empty, 0] = 1.0
    return x, y.contiguous()
