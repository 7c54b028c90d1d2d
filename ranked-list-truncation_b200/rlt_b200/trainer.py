"""Device-resident host loop: what `Trainer.run` of the reference does per epoch
(run.py:113-206) -- shuffled batches, forward, criterion, backward, Adam, arg-max / BiCut cut, Metric.f1 / Metric.dcg,
then the same over the test split under `no_grad` -- with the split held in HBM, the batch collated by one gather launch
(`DeviceLoader`), the cut positions and per-list F1 / DCG computed on the device (`rlt_eval_cut`) and the optimizer step
fused (`FusedAdam`).  run.py synchronises three times per step (`output.cpu()`, `y.cpu()`, `loss.item()`); this loop
reads the step losses and per-list metrics back ONCE per epoch.

The numbers it reports are run.py's: per-step `train/loss_step`, per-epoch `{train,test}/{loss,F1,DCG}_epoch` (epoch value
= mean over the batches of the batch means, run.py:150,190).  Under the same `torch.manual_seed` it builds the model,
the criterion and the batches in run.py's order, so it lands on the trajectory of the unmodified run.py
(tests/golden/run_py_traj.json; tests/test_zzzz_trainer_gpu.py needs no reference tree).

`run.py` itself keeps working unchanged on top of the drop-in packages (tools/run_reference.py); this module is the
B200-side alternative to its loop, not a replacement of its CLI."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .data import DeviceLoader, rank_tensors
from .optim import FusedAdam

# hyper_parameter_drmm_tks.conf of the reference (the values run.py:338-346 reads for --dataset-name drmm_tks)
HYPER = {
    "bicut": dict(batch_size=63, lr=1e-4, weight_decay=0.0024756345581373493, dropout=0.01),
    "choopy": dict(batch_size=63, lr=1e-3, weight_decay=0.0024756345581373493, dropout=0.1),
    "attncut": dict(batch_size=63, lr=3e-5, weight_decay=0.0014756345581373493, dropout=0.1),
    "mtchoopy": dict(batch_size=63, lr=1e-3, weight_decay=0.0024756345581373493, dropout=0.1, rerank_weight=0.5,
                     class_weight=0.5),
    "mtattncut": dict(batch_size=63, lr=3e-5, weight_decay=0.0024756345581373493, dropout=0.1, rerank_weight=0.5,
                      class_weight=0.5),
    "mmoecut": dict(batch_size=63, lr=3e-5, weight_decay=0.0, dropout=0.1, rerank_weight=0.4, class_weight=0.6),
}
ONE_FEATURE = ("choopy", "mtchoopy")          # cp_dataloader: the retrieval score alone (run.py:67,80)


def build(model_name: str, *, criterion: str = "dcg", num_tasks: float = 3, seq_len: int = 300, dropout=None,
          div_type: str = "js", augmented_reward: int = 1, **hyper):
    """(model, criterion) exactly as Trainer.__init__ builds them (run.py:59-100), in that order (the initial weights
    and MtCutLoss's unused random Parameter consume torch's generator)."""
    import models
    from utils import losses
    h = dict(HYPER[model_name])
    h.update(hyper)
    p = h["dropout"] if dropout is None else dropout
    if model_name == "bicut":
        return models.BiCut(input_size=3, dropout=p), losses.BiCutLoss(metric=criterion)
    if model_name == "choopy":
        return models.Choopy(seq_len=seq_len, dropout=p), losses.ChoopyLoss(metric=criterion)
    if model_name == "attncut":
        return (models.AttnCut(input_size=3, dropout=p),
                losses.DivLoss(metric=criterion, div_type=div_type, augmented=augmented_reward))
    if model_name == "mtchoopy":
        return (models.MtChoopy(seq_len=seq_len, num_tasks=num_tasks, dropout=p),
                losses.MtCutLoss(metric=criterion, rerank_weight=h["rerank_weight"], classi_weight=h["class_weight"],
                                 num_tasks=num_tasks))
    if model_name == "mtattncut":
        return (models.MtAttnCut(input_size=3, num_tasks=num_tasks, dropout=p),
                losses.MtCutLoss(metric=criterion, rerank_weight=h["rerank_weight"], classi_weight=h["class_weight"],
                                 num_tasks=num_tasks))
    if model_name == "mmoecut":
        return (models.MMOECut(seq_len=seq_len, num_tasks=num_tasks, input_size=3, dropout=p, num_experts=3),
                losses.MtCutLoss(metric=criterion, num_tasks=num_tasks))
    raise ValueError(f"unknown model {model_name!r}")


class DeviceTrainer:
    """trainer = DeviceTrainer("choopy", X_train, X_test, y_train, y_test, criterion="f1"); trainer.run(epochs)"""

    def __init__(self, model_name: str, X_train, X_test, y_train, y_test, *, criterion: str = "dcg", num_tasks: float = 3,
                 dropout=None, lr=None, weight_decay=None, batch_size=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("rlt_b200 has no CPU path: DeviceTrainer needs a CUDA device")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model_name = model_name
        h = dict(HYPER[model_name])
        for k, v in (("lr", lr), ("weight_decay", weight_decay), ("batch_size", batch_size)):
            if v is not None:
                h[k] = v
        self.hyper = h
        self.seq_len = int(X_train.shape[1])
        # run.py:59-104: loaders (no random numbers yet), model, criterion, optimizer
        self.train_loader = DeviceLoader(X_train, y_train, h["batch_size"], True, device=self.device)
        self.test_loader = DeviceLoader(X_test, y_test, h["batch_size"], True, device=self.device)
        self.model, self.criterion = build(model_name, criterion=criterion, num_tasks=num_tasks, seq_len=self.seq_len,
                                           dropout=dropout)
        self.model = self.model.to(self.device)
        self.criterion = self.criterion.to(self.device)
        self.optimizer = FusedAdam(self.model.parameters(), lr=h["lr"], weight_decay=h["weight_decay"])
        self.scalars = {k: [] for k in ("train/loss_step", "train/loss_epoch", "train/F1_epoch", "train/DCG_epoch",
                                        "test/loss_epoch", "test/F1_epoch", "test/DCG_epoch")}
        self.best_test_f1 = self.best_test_dcg = -float("inf")

    @classmethod
    def from_pickles(cls, model_name: str, database, dataset_name: str = "drmm_tks", **kw):
        """The reference's on-disk formats (dataloader/attncut_dataloader.py:21-59, choopy_dataloader.py:21-45)."""
        x_tr, x_te, y_tr, y_te = rank_tensors(database, dataset_name, stats=model_name not in ONE_FEATURE)
        return cls(model_name, x_tr, x_te, y_tr, y_te, **kw)

    def _cut_metrics(self, output, y):
        """run.py:131-145 on the device: (f1 [B], dcg [B]) float64."""
        if self.model_name == "bicut":
            _, _, _, f1, dcg = ops.eval_cut(output.detach().contiguous(), y, mode=1)
        else:
            last = output[-1] if isinstance(output, (list, tuple)) else output
            _, _, _, f1, dcg = ops.eval_cut(last.detach().reshape(y.shape).contiguous(), y, mode=0)
        return f1, dcg

    def _epoch(self, loader, train: bool):
        losses, f1s, dcgs = [], [], []
        self.model.train(train)
        for xb, yb in loader:
            if train:
                self.optimizer.zero_grad()
                output = self.model(xb)
                loss = self.criterion(output, yb)
                loss.backward()
                self.optimizer.step()
            else:
                with torch.no_grad():
                    output = self.model(xb)
                    loss = self.criterion(output, yb)
            f1, dcg = self._cut_metrics(output, yb)
            losses.append(loss.detach().reshape(()))
            f1s.append(f1)
            dcgs.append(dcg)
        # the one host synchronisation of the epoch
        step_losses = torch.stack(losses).cpu().tolist()
        f1_means = [float(np.mean(v.cpu().numpy())) for v in f1s]          # Metric.f1 / Metric.dcg: np.mean per batch
        dcg_means = [float(np.mean(v.cpu().numpy())) for v in dcgs]
        n = len(step_losses)
        return step_losses, sum(step_losses) / n, sum(f1_means) / n, sum(dcg_means) / n

    def train_epoch(self, epoch: int):
        step_losses, loss, f1, dcg = self._epoch(self.train_loader, True)
        self.scalars["train/loss_step"].extend(step_losses)
        for tag, v in (("train/loss_epoch", loss), ("train/F1_epoch", f1), ("train/DCG_epoch", dcg)):
            self.scalars[tag].append(v)
        return loss, f1, dcg

    def test(self, epoch: int):
        _, loss, f1, dcg = self._epoch(self.test_loader, False)
        for tag, v in (("test/loss_epoch", loss), ("test/F1_epoch", f1), ("test/DCG_epoch", dcg)):
            self.scalars[tag].append(v)
        self.best_test_f1 = max(self.best_test_f1, f1)
        self.best_test_dcg = max(self.best_test_dcg, dcg)
        return loss, f1, dcg

    def run(self, epochs: int):
        for epoch in range(epochs):
            self.train_epoch(epoch)
            self.test(epoch)
        return self.scalars
