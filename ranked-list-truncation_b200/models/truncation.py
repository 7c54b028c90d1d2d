"""The six truncation model families of the reference behind their original nn.Module API.

Each class builds the SAME torch submodules, in the same order, as the reference constructor it
mirrors (cited per class).  Those submodules are only parameter holders: this guarantees identical
parameter names (checkpoints interchange both ways, run.py:208-220), identical shapes and — under
the same torch seed — identical initial values.  `forward` never calls them; it hands their
parameters to the CUDA kernels through rlt_b200.autograd.

Train-mode dropout follows torch's placement (four sites per encoder layer, the two logits of BiCut) with masks from
a counter hash instead of torch's Philox stream (csrc/dropout.cuh): same distribution, different bits.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn

from rlt_b200 import autograd as F
from rlt_b200.ops import ENCODER_PARAM_ORDER


def _encoder(d_model: int, n_head: int, num_layers: int, dropout: float) -> nn.TransformerEncoder:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch warns that nested tensors need batch_first; irrelevant here
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=n_head, dropout=dropout)
        return nn.TransformerEncoder(layer, num_layers=num_layers)


def _bilstm(input_size: int, hidden: int = 128, layers: int = 2) -> nn.LSTM:
    return nn.LSTM(input_size=input_size, hidden_size=hidden, num_layers=layers, batch_first=True, bidirectional=True)


def _lin_softmax(d_model: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_features=d_model, out_features=1), nn.Softmax(dim=1))


def _lin_sigmoid(d_model: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_features=d_model, out_features=1), nn.Sigmoid())


def _encoder_params(enc: nn.TransformerEncoder):
    flat = []
    for layer in enc.layers:
        named = dict(layer.named_parameters())
        flat.extend(named[n] for n in ENCODER_PARAM_ORDER)
    return flat


class _Base(nn.Module):
    _dropout_p = 0.0
    # Which axis the encoder layers attend over.  "lists" (default) is what the reference computes: its
    # nn.TransformerEncoderLayer is built without batch_first and fed [B, L, d], so the B lists of a forward call attend to
    # each other at every position.  "positions" is the papers' intent -- the L positions of each list
    # attend to each other -- and equals `enc(x.transpose(0, 1)).transpose(0, 1)` on the reference module.  Set it on an
    # instance (`model.attend = "positions"`); parameters, state_dict and everything else are unchanged.
    attend = "lists"

    def _p(self) -> float:
        """Dropout probability in effect: the constructor's value in train(), 0 in eval()."""
        return float(self._dropout_p) if self.training else 0.0

    def _encode(self, x, enc: nn.TransformerEncoder):
        layer0 = enc.layers[0]
        if self.attend not in ("lists", "positions"):
            raise ValueError(f"attend must be 'lists' or 'positions', got {self.attend!r}")
        fn = F.EncoderStack if self.attend == "lists" else F.EncoderStackWithin
        return fn.apply(x, layer0.self_attn.num_heads, 1, layer0.norm1.eps, self._p(), *_encoder_params(enc))

    @staticmethod
    def _heads(h, linears):
        """logits (one [B, L] tensor per head) of H Linear(d, 1) heads evaluated in one pass over h."""
        return F.HeadDots.apply(h, *[t for m in linears for t in (m.weight, m.bias)])


class Choopy(_Base):
    """Reference models/Choopy.py:6-23."""

    def __init__(self, seq_len: int = 300, d_model: int = 128, n_head: int = 8, num_layers: int = 3, dropout=0.2):
        super().__init__()
        self.seq_len = seq_len
        self._dropout_p = float(dropout)
        self.position_encoding = nn.Parameter(torch.randn(self.seq_len, 127), requires_grad=True)
        self.attention_layer = _encoder(d_model, n_head, num_layers, dropout)
        self.decison_layer = _lin_softmax(d_model)

    def forward(self, x):
        h = self._encode(F.ChoopyEmbed.apply(x, self.position_encoding), self.attention_layer)
        z = self._heads(h, [self.decison_layer[0]])
        return F.SoftmaxLists.apply(z[0]).unsqueeze(2)


class MtChoopy(_Base):
    """Reference models/MtChoopy.py:5-32."""

    def __init__(self, seq_len: int = 300, d_model: int = 128, n_head: int = 8, num_layers: int = 3,
                 num_tasks: float = 3, dropout: float = 0.4):
        super().__init__()
        self.seq_len = seq_len
        self.num_tasks = num_tasks
        self._dropout_p = float(dropout)
        self.position_encoding = nn.Parameter(torch.randn(self.seq_len, 127), requires_grad=True)
        self.encoding_layer = _encoder(d_model, n_head, num_layers, dropout)
        self.classi = _lin_sigmoid(d_model)
        self.rerank = nn.Linear(in_features=d_model, out_features=1)
        self.decison_layer = _lin_softmax(d_model)

    def forward(self, x):
        h = self._encode(F.ChoopyEmbed.apply(x, self.position_encoding), self.encoding_layer)
        return _mt_outputs(self, h)


def _mt_outputs(self, h):
    """The three heads of MtChoopy / MtAttnCut (MtChoopy.py:27-32, MtAttnCut.py:24-29)."""
    z = self._heads(h, [self.classi[0], self.rerank, self.decison_layer[0]])
    y0 = torch.sigmoid(z[0]).unsqueeze(2)
    y1 = z[1].unsqueeze(2)
    y2 = F.SoftmaxLists.apply(z[2]).unsqueeze(2)
    if self.num_tasks == 3:
        return [y0, y1, y2]
    if self.num_tasks == 2.1:
        return [y0, y2]
    return [y1, y2]


class BiCut(_Base):
    """Reference models/Bicut.py:5-21."""

    def __init__(self, input_size=231449, lstm_hiden_size=128, lstm_layers=2, fc_dimensions=256, dropout=0.4):
        super().__init__()
        self._dropout_p = float(dropout)
        self.bilstm = _bilstm(input_size, lstm_hiden_size, lstm_layers)
        self.fc = nn.Linear(in_features=lstm_hiden_size * 2, out_features=fc_dimensions)
        self.softmax = nn.Sequential(nn.ReLU(), nn.Linear(in_features=fc_dimensions, out_features=2),
                                     nn.Dropout(dropout), nn.Softmax(dim=2))

    def forward(self, x):
        h = F.BiLstm.apply(x, self.bilstm.hidden_size, self.bilstm.num_layers, *self.bilstm._flat_weights)
        return F.BicutHead.apply(h, self.fc.weight, self.fc.bias, self.softmax[1].weight, self.softmax[1].bias, self._p())


class AttnCut(_Base):
    """Reference models/AttnCut.py:5-20."""

    def __init__(self, input_size: int = 3, d_model: int = 256, n_head: int = 4, num_layers: int = 1,
                 dropout: float = 0.4):
        super().__init__()
        self._dropout_p = float(dropout)
        self.encoding_layer = _bilstm(input_size)
        self.attention_layer = _encoder(d_model, n_head, num_layers, dropout)
        self.decison_layer = _lin_softmax(d_model)

    def forward(self, x):
        e = self.encoding_layer
        h = F.BiLstm.apply(x, e.hidden_size, e.num_layers, *e._flat_weights)
        h = self._encode(h, self.attention_layer)
        z = self._heads(h, [self.decison_layer[0]])
        return F.SoftmaxLists.apply(z[0]).unsqueeze(2)


class MtAttnCut(_Base):
    """Reference models/MtAttnCut.py:4-29."""

    def __init__(self, input_size: int = 3, d_model: int = 256, n_head: int = 4, num_layers: int = 1,
                 num_tasks: float = 3, dropout: float = 0.4):
        super().__init__()
        self.num_tasks = num_tasks
        self._dropout_p = float(dropout)
        self.pre_encoding = _bilstm(input_size)
        self.encoding_layer = _encoder(d_model, n_head, num_layers, dropout)
        self.classi = _lin_sigmoid(d_model)
        self.rerank = nn.Linear(in_features=d_model, out_features=1)
        self.decison_layer = _lin_softmax(d_model)

    def forward(self, x):
        e = self.pre_encoding
        h = F.BiLstm.apply(x, e.hidden_size, e.num_layers, *e._flat_weights)
        return _mt_outputs(self, self._encode(h, self.encoding_layer))


def _moe_towers(h, gates, towers, experts):
    """Gate softmax, expert mixture and tower Linear(d, 1) of every tower in one kernel pass (MMOECut.py:90-105); the
    parameters go in one by one so that a tower without gradient keeps `.grad is None` (rlt_b200.autograd.MoeGateMix)."""
    return F.MoeGateMix.apply(h, len(towers), *gates, *[t.linear.weight for t in towers],
                              *[t.linear.bias for t in towers], *experts)


class _Expert(nn.Module):
    """Parameter holder with the reference's attribute name (MMOECut.py:6-14)."""

    def __init__(self, d_model, n_head, num_layers, dropout):
        super().__init__()
        self.attention_layer = _encoder(d_model, n_head, num_layers, dropout)


class _Tower(nn.Module):
    """Parameter holder for TowerCut / TowerClass / TowerRerank (MMOECut.py:17-53)."""

    def __init__(self, d_model, attr: str, act: str):
        super().__init__()
        setattr(self, attr, _lin_sigmoid(d_model) if act == "sigmoid" else _lin_softmax(d_model))
        self.attr, self.act = attr, act

    @property
    def linear(self) -> nn.Linear:
        return getattr(self, self.attr)[0]


class MMOECut(_Base):
    """Reference models/MMOECut.py:56-110."""

    def __init__(self, seq_len: int = 300, num_experts=3, num_tasks=3, input_size=3, encoding_size=128, d_model=256,
                 n_head=4, num_layers=1, dropout=0.2):
        super().__init__()
        self.seq_len = seq_len
        self.expert_hidden = d_model
        self._dropout_p = float(dropout)
        self.pre_encoding = _bilstm(input_size, encoding_size)
        self.softmax = nn.Softmax(dim=1)
        self.experts = nn.ModuleList([_Expert(self.expert_hidden, n_head, num_layers, dropout)
                                      for _ in range(num_experts)])
        self.w_gates = nn.ParameterList([nn.Parameter(torch.randn(encoding_size * self.seq_len * 2, num_experts),
                                                      requires_grad=True) for _ in range(int(num_tasks))])
        cls = lambda: _Tower(self.expert_hidden, "classification_layer", "sigmoid")  # noqa: E731
        rer = lambda: _Tower(self.expert_hidden, "rerank_layer", "softmax")          # noqa: E731
        cut = lambda: _Tower(self.expert_hidden, "cut_layer", "softmax")             # noqa: E731
        if num_tasks == 3:
            self.towers = nn.ModuleList([cls(), rer(), cut()])
        elif num_tasks == 2.1:
            self.towers = nn.ModuleList([cls(), cut()])
        elif num_tasks == 2.2:
            self.towers = nn.ModuleList([rer(), cut()])

    def forward(self, x):
        e = self.pre_encoding
        h = F.BiLstm.apply(x, e.hidden_size, e.num_layers, *e._flat_weights)
        experts = [self._encode(h, ex.attention_layer) for ex in self.experts]
        z = _moe_towers(h, list(self.w_gates), list(self.towers), experts)   # one [B, L] logit tensor per tower
        outs = []
        for t, tower in enumerate(self.towers):
            outs.append((torch.sigmoid(z[t]) if tower.act == "sigmoid" else F.SoftmaxLists.apply(z[t])).unsqueeze(2))
        return outs


class MOECut(_Base):
    """Reference models/MOECut.py:56-109: MMOECut with ONE gate Parameter shared by all
    towers (`w_gates` [2*encoding_size*seq_len, num_experts], :68; every tower sees the same mixture, :94-104)."""

    def __init__(self, seq_len: int = 300, num_experts=3, num_tasks=3, input_size=3, encoding_size=128, d_model=256,
                 n_head=4, num_layers=1, dropout=0.2):
        super().__init__()
        self.seq_len = seq_len
        self.expert_hidden = d_model
        self._dropout_p = float(dropout)
        self.pre_encoding = _bilstm(input_size, encoding_size)
        self.softmax = nn.Softmax(dim=1)
        self.experts = nn.ModuleList([_Expert(self.expert_hidden, n_head, num_layers, dropout)
                                      for _ in range(num_experts)])
        self.w_gates = nn.Parameter(torch.randn(encoding_size * self.seq_len * 2, num_experts), requires_grad=True)
        cls = lambda: _Tower(self.expert_hidden, "classification_layer", "sigmoid")  # noqa: E731
        rer = lambda: _Tower(self.expert_hidden, "rerank_layer", "softmax")          # noqa: E731
        cut = lambda: _Tower(self.expert_hidden, "cut_layer", "softmax")             # noqa: E731
        if num_tasks == 3:
            self.towers = nn.ModuleList([cls(), rer(), cut()])
        elif num_tasks == 2.1:
            self.towers = nn.ModuleList([cls(), cut()])
        elif num_tasks == 2.2:
            self.towers = nn.ModuleList([rer(), cut()])

    def forward(self, x):
        e = self.pre_encoding
        h = F.BiLstm.apply(x, e.hidden_size, e.num_layers, *e._flat_weights)
        experts = [self._encode(h, ex.attention_layer) for ex in self.experts]
        # the shared gate is presented once per tower; autograd sums the three gate gradients into the one Parameter
        z = _moe_towers(h, [self.w_gates] * len(self.towers), list(self.towers), experts)
        outs = []
        for t, tower in enumerate(self.towers):
            outs.append((torch.sigmoid(z[t]) if tower.act == "sigmoid" else F.SoftmaxLists.apply(z[t])).unsqueeze(2))
        return outs


class PLECut(_Base):
    """Reference models/PLECut.py:56-104 (row N4): three towers, each with its own gate over a SUBSET of the experts --
    class tower: experts [0:2], rerank tower: experts [1:3], cut tower: experts [0:3] (gate widths 2, 2, 3; :68-70,
    :81-83, :94-96).  Each tower is one gate/mix/tower call on its expert subset."""

    _SUBSETS = ((0, 2), (1, 3), (0, 3))

    def __init__(self, seq_len: int = 300, num_experts=3, input_size=3, encoding_size=128, d_model=256, n_head=2,
                 num_layers=1, dropout=0.1):
        super().__init__()
        self.seq_len = seq_len
        self.expert_hidden = d_model
        self._dropout_p = float(dropout)
        self.pre_encoding = _bilstm(input_size, encoding_size)
        self.softmax = nn.Softmax(dim=1)
        self.experts = nn.ModuleList([_Expert(self.expert_hidden, n_head, num_layers, dropout)
                                      for _ in range(num_experts)])
        n_in = encoding_size * self.seq_len * 2
        self.w_gates = nn.ParameterList([nn.Parameter(torch.randn(n_in, 2), requires_grad=True),
                                         nn.Parameter(torch.randn(n_in, 2), requires_grad=True),
                                         nn.Parameter(torch.randn(n_in, 3), requires_grad=True)])
        self.towers = nn.ModuleList([_Tower(self.expert_hidden, "classification_layer", "sigmoid"),
                                     _Tower(self.expert_hidden, "rerank_layer", "softmax"),
                                     _Tower(self.expert_hidden, "cut_layer", "softmax")])

    def forward(self, x):
        e = self.pre_encoding
        h = F.BiLstm.apply(x, e.hidden_size, e.num_layers, *e._flat_weights)
        experts = [self._encode(h, ex.attention_layer) for ex in self.experts]
        outs = []
        for tower, gate, (lo, hi) in zip(self.towers, self.w_gates, self._SUBSETS):
            z = _moe_towers(h, [gate], [tower], experts[lo:hi])[0]
            outs.append((torch.sigmoid(z) if tower.act == "sigmoid" else F.SoftmaxLists.apply(z)).unsqueeze(2))
        return outs


class _StandaloneTower(_Tower):
    """TowerClass / TowerRerank / TowerCut used on their own (verify_probe.py:66-71 builds them directly as probes):
    Linear(d_model, 1) evaluated by the head kernel, then sigmoid or the softmax over the list positions."""

    def forward(self, x):
        z = F.HeadDots.apply(x, self.linear.weight, self.linear.bias)[0]
        return (torch.sigmoid(z) if self.act == "sigmoid" else F.SoftmaxLists.apply(z)).unsqueeze(2)


class TowerCut(_StandaloneTower):
    """Reference models/Probe.py:17-27."""

    def __init__(self, d_model):
        super().__init__(d_model, "cut_layer", "softmax")


class TowerClass(_StandaloneTower):
    """Reference models/Probe.py:30-40."""

    def __init__(self, d_model):
        super().__init__(d_model, "classification_layer", "sigmoid")


class TowerRerank(_StandaloneTower):
    """Reference models/Probe.py:43-53."""

    def __init__(self, d_model):
        super().__init__(d_model, "rerank_layer", "softmax")


class TaskC(_StandaloneTower):
    """Reference models/Classification.py:3-13 (the single-task classifier of verify_classification.py)."""

    def __init__(self, d_model: int = 128) -> None:
        super().__init__(d_model, "classification_layer", "sigmoid")


class TaskR(_StandaloneTower):
    """Reference models/Rerank.py:3-13."""

    def __init__(self, d_model: int = 128) -> None:
        super().__init__(d_model, "rerank_layer", "softmax")


class ProbeBase(_Base):
    """Reference models/Probe.py:56-99 (row N4): the MMOECut graph with two experts by default and always three towers
    (class, rerank, cut; :67-71 -- `num_tasks` only sets the number of gates, and the towers zip against them, :93),
    returning the intermediate representations next to the tower outputs: `(experts_in, experts_o, final_output)`."""

    def __init__(self, seq_len: int = 300, num_experts=2, num_tasks=3, input_size=3, encoding_size=128, d_model=256,
                 n_head=4, num_layers=1, dropout=0.2):
        super().__init__()
        self.seq_len = seq_len
        self.expert_hidden = d_model
        self._dropout_p = float(dropout)
        self.pre_encoding = _bilstm(input_size, encoding_size)
        self.softmax = nn.Softmax(dim=1)
        self.experts = nn.ModuleList([_Expert(self.expert_hidden, n_head, num_layers, dropout)
                                      for _ in range(num_experts)])
        self.w_gates = nn.ParameterList([nn.Parameter(torch.randn(encoding_size * self.seq_len * 2, num_experts),
                                                      requires_grad=True) for _ in range(int(num_tasks))])
        self.towers = nn.ModuleList([TowerClass(self.expert_hidden), TowerRerank(self.expert_hidden),
                                     TowerCut(self.expert_hidden)])

    def forward(self, x):
        e = self.pre_encoding
        h = F.BiLstm.apply(x, e.hidden_size, e.num_layers, *e._flat_weights)
        experts = [self._encode(h, ex.attention_layer) for ex in self.experts]
        towers = list(self.towers)[:len(self.w_gates)]          # zip(towers, towers_input) of Probe.py:93
        z = _moe_towers(h, list(self.w_gates), towers, experts)
        outs = []
        for t, tower in enumerate(towers):
            outs.append((torch.sigmoid(z[t]) if tower.act == "sigmoid" else F.SoftmaxLists.apply(z[t])).unsqueeze(2))
        return h, experts, outs


class Probe(_Base):
    """Reference models/Probe.py:102-122: six Linear(d, 1) probes -- class / rerank on the LSTM representation and on
    each of the two expert outputs.  The two probes that read the same tensor share one pass of the head kernel."""

    def __init__(self, encoding_size=128, d_model=256) -> None:
        super().__init__()
        self.probe_c1 = TowerClass(d_model=encoding_size * 2)
        self.probe_r1 = TowerRerank(d_model=encoding_size * 2)
        self.probe_ce1 = TowerClass(d_model=d_model)
        self.probe_ce2 = TowerClass(d_model=d_model)
        self.probe_re1 = TowerRerank(d_model=d_model)
        self.probe_re2 = TowerRerank(d_model=d_model)

    def _pair(self, h, cls: TowerClass, rer: TowerRerank):
        z = self._heads(h, [cls.linear, rer.linear])
        return torch.sigmoid(z[0]).unsqueeze(2), F.SoftmaxLists.apply(z[1]).unsqueeze(2)

    def forward(self, experts_in, experts_o):
        c1, r1 = self._pair(experts_in, self.probe_c1, self.probe_r1)
        ce1, re1 = self._pair(experts_o[0], self.probe_ce1, self.probe_re1)
        ce2, re2 = self._pair(experts_o[1], self.probe_ce2, self.probe_re2)
        return c1, r1, ce1, ce2, re1, re2
