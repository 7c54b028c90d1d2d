"""Drop-in replacement of the reference utils/losses.py criteria (same classes, ctor args, call signature).

The reference computes its reward matrices with a Python double loop over every (list, cut) cell
(utils/losses.py:56-65, 80-89, 216-225); here each criterion is ONE fused kernel launch that
produces the loss and the gradient w.r.t. the model output (rlt_cut_loss / rlt_aux_heads_loss /
rlt_bicut_loss).  WassDistLoss (never instantiated by run.py) is outside the hot path.
"""
from __future__ import annotations

import torch as t
from torch import nn

from rlt_b200 import autograd as F


def _probs(output: t.Tensor) -> t.Tensor:
    """[B, L, 1] -> [B, L] (the reference squeezes; keep the batch dim for B == 1)."""
    return output.reshape(output.shape[0], output.shape[1])


class BiCutLoss(nn.Module):
    """Reference utils/losses.py:11-45."""

    def __init__(self, alpha: float = 0.65, r: float = 0.0971134020, metric: str = 'nci'):
        super().__init__()
        self.metric = metric
        self.alpha = alpha
        self.r = r

    def forward(self, output: t.Tensor, labels: t.Tensor):
        return F.BicutLoss.apply(output, labels, self.alpha, self.r, self.metric == 'nci')


class ChoopyLoss(nn.Module):
    """Reference utils/losses.py:48-68."""

    def __init__(self, metric: str = 'f1'):
        super().__init__()
        self.metric = metric

    def forward(self, output: t.Tensor, labels: t.Tensor):
        return F.CutLoss.apply(_probs(output), labels, "choopy", 'f1' if self.metric == 'f1' else 'dcg', 1.0)


class AttnCutLoss(nn.Module):
    """Reference utils/losses.py:71-96 (RAML)."""

    def __init__(self, metric: str = 'f1', tau: float = 0.95):
        super().__init__()
        self.metric = metric
        self.tau = tau

    def forward(self, output: t.Tensor, labels: t.Tensor):
        return F.CutLoss.apply(_probs(output), labels, "raml", 'f1' if self.metric == 'f1' else 'dcg', self.tau)


class RerankLoss(nn.Module):
    """Reference utils/losses.py:99-141: hinge on mean(irrelevant) - mean(relevant) + margin over the batch."""

    def __init__(self, margin: float = 5e-4, reduction: str = 'mean'):
        super().__init__()
        self.margin = margin
        self.reduction = reduction

    def forward(self, output: t.Tensor, labels: t.Tensor):
        return F.AuxHeadsLoss.apply(None, output, labels, 0.0, 1.0, self.margin)


class DivLoss(nn.Module):
    """Reference utils/losses.py:194-233."""

    def __init__(self, metric: str = 'f1', tau: float = 0.85, div_type: str = 'kl', augmented: bool = True):
        super().__init__()
        self.metric = metric
        self.div_type = div_type
        self.augmented = augmented
        self.tau = tau if self.augmented else 1.

    def forward(self, output: t.Tensor, labels: t.Tensor):
        kind = "kl" if self.div_type == 'kl' else "js"
        return F.CutLoss.apply(_probs(output), labels, kind, 'f1' if self.metric == 'f1' else 'dcg', self.tau)


class MtCutLoss(nn.Module):
    """Reference utils/losses.py:164-191."""

    def __init__(self, metric: str = 'f1', rerank_weight: float = 0.5, classi_weight: float = 0.5,
                 num_tasks: float = 3):
        super().__init__()
        self.rerank_weight, self.classi_weight = rerank_weight, classi_weight
        # unused by the reference too, but part of its state_dict (utils/losses.py:173)
        self.weights = nn.Parameter(t.randn(int(num_tasks)), requires_grad=True)
        self.cutloss = DivLoss(metric=metric, div_type='js', augmented=True)
        self.rerankloss = RerankLoss()
        self.num_tasks = num_tasks

    def forward(self, output, labels: t.Tensor):
        pred_y = rerank_y = None
        if self.num_tasks == 3:
            pred_y, rerank_y, cut_y = output
        elif self.num_tasks == 2.1:
            pred_y, cut_y = output
        else:
            rerank_y, cut_y = output
        loss = self.cutloss(cut_y, labels)
        aux = F.AuxHeadsLoss.apply(pred_y, rerank_y, labels, self.classi_weight, self.rerank_weight,
                                   self.rerankloss.margin)
        return loss.add(aux)
