"""CPU baseline: the reference's training / inference step re-stated with the same third-party calls it
makes (torch.nn.LSTM, torch.nn.TransformerEncoder WITHOUT batch_first, Linear heads) and its Python
B x L reward loop (oracle/rlt_oracle.reward_matrix_loop).  TEST / BENCH INFRASTRUCTURE ONLY
(bench.py `cpu_baseline` and `--impl reference`); `kind: "port"` because /root/reference cannot travel
to the GPU box.  Cost structure matches the reference: eager ATen ops on all host cores plus ~46 tiny
ATen dispatches per (list, position) cell in the criterion.
"""
from __future__ import annotations

import time
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import rlt_oracle as O


def _enc(d, h, n):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=d, nhead=h, dropout=0.0), num_layers=n)


class PortModel(nn.Module):
    """Forward graphs of reference models/{Bicut,Choopy,AttnCut,MtChoopy,MtAttnCut,MMOECut}.py (dropout 0)."""

    def __init__(self, kind: str, seq_len: int = 300, n_features: int = 3):
        super().__init__()
        self.kind = kind
        self.seq_len = seq_len
        if kind in ("choopy", "mtchoopy"):
            self.pe = nn.Parameter(torch.randn(seq_len, 127))
            self.enc = _enc(128, 8, 3)
            d = 128
        else:
            self.lstm = nn.LSTM(n_features, 128, num_layers=2, batch_first=True, bidirectional=True)
            d = 256
            if kind in ("attncut", "mtattncut"):
                self.enc = _enc(256, 4, 1)
            elif kind == "mmoecut":
                self.experts = nn.ModuleList([_enc(256, 4, 1) for _ in range(3)])
                self.gates = nn.ParameterList([nn.Parameter(torch.randn(256 * seq_len, 3)) for _ in range(3)])
            elif kind == "bicut":
                self.fc = nn.Linear(256, 256)
                self.cls = nn.Linear(256, 2)
        n_heads = {"choopy": 1, "attncut": 1, "bicut": 0}.get(kind, 3)
        self.heads = nn.ModuleList([nn.Linear(d, 1) for _ in range(n_heads)])

    def forward(self, x):
        k = self.kind
        if k in ("choopy", "mtchoopy"):
            h = self.enc(torch.cat((x, self.pe.expand(x.shape[0], -1, -1)), dim=2))
        else:
            h = self.lstm(x)[0]
            if k == "bicut":
                return torch.softmax(self.cls(torch.relu(self.fc(h))), dim=2)
            if k in ("attncut", "mtattncut"):
                h = self.enc(h)
        if k in ("choopy", "attncut"):
            return torch.softmax(self.heads[0](h), dim=1)
        if k in ("mtchoopy", "mtattncut"):
            return [torch.sigmoid(self.heads[0](h)), self.heads[1](h), torch.softmax(self.heads[2](h), dim=1)]
        ex = torch.stack([e(h) for e in self.experts])
        outs = []
        for t in range(3):
            gate = torch.softmax(h.reshape(h.shape[0], -1) @ self.gates[t], dim=1)
            z = self.heads[t]((gate.t().reshape(3, -1, 1, 1) * ex).sum(0))
            outs.append(torch.sigmoid(z) if t == 0 else torch.softmax(z, dim=1))
        return outs


def train_step(model: PortModel, x, y, metric="f1", optimizer=None):
    """run.py:121-145: forward, criterion (Python reward loop), backward, optimizer step (run.py:129, when an optimizer
    is given), host cut + metrics."""
    model.zero_grad(set_to_none=True)
    out = model(x)
    loss = O.criterion_for(model.kind, metric=metric, loop=True)(out, y)
    loss.backward()
    if optimizer is not None:
        optimizer.step()
    last = out[-1] if isinstance(out, list) else out
    if model.kind == "bicut":
        ks = O.bicut_cut_positions(last.detach().numpy())
    else:
        ks = O.cut_positions(last.detach().numpy())
    yn = y.numpy()
    return loss.item(), O.metric_f1(yn, ks), O.metric_dcg(yn, ks)


@torch.no_grad()
def infer_step(model: PortModel, x, y):
    """run.py:166-186 without the criterion: forward, argmax cut, Metric.f1 / Metric.dcg."""
    out = model(x)
    last = out[-1] if isinstance(out, list) else out
    ks = O.bicut_cut_positions(last.numpy()) if model.kind == "bicut" else O.cut_positions(last.numpy())
    yn = y.numpy()
    return O.metric_f1(yn, ks), O.metric_dcg(yn, ks)


def time_lists_per_s(kind: str, x, y, group_size: int = 64, mode: str = "train", steps: int = 2, warmup: int = 1):
    """Median lists/s of `steps` reference-style steps on batches of `group_size` lists (host CPU, all cores)."""
    torch.manual_seed(1234)
    model = PortModel(kind, seq_len=x.shape[1], n_features=x.shape[2])
    model.train() if mode == "train" else model.eval()
    optimizer = torch.optim.Adam(model.parameters(), lr=3e-5, weight_decay=1e-3) if mode == "train" else None   # run.py:104
    n_batches = x.shape[0] // group_size
    times = []
    for i in range(warmup + steps):
        b = i % n_batches
        xb, yb = x[b * group_size:(b + 1) * group_size], y[b * group_size:(b + 1) * group_size]
        t0 = time.perf_counter()
        if mode == "train":
            train_step(model, xb, yb, optimizer=optimizer)
        else:
            infer_step(model, xb, yb)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return group_size / float(np.median(times)), times
