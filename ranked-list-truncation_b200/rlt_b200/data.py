"""Seeded synthetic robust04-shaped ranked lists.

score_j : sorted-descending( s0 * (1 - 0.67 j/L) + 0.3 N(0,1) ), s0 = 9   (DRMM-TKS-like range)
extra features : U(0,1)   (the two neighbour-similarity features of the AttnCut data)
label_j ~ Bernoulli(a * exp(-j/tau)), a ~ U(0.2, 0.9), tau ~ U(10, 80) per list, >= 1 relevant forced
Returns X [n, L, F] float32 and y [n, L] float32 in {0,1}, the shapes the reference loaders yield
(dataloader/attncut_dataloader.py:21-59, choopy_dataloader.py:21-45).

Row N3: `DeviceLoader` keeps a whole split resident in HBM and replaces the torch DataLoader
of dataloader/attncut_dataloader.py:82-87 -- same iteration protocol, same batches under the same torch seed.
"""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np
import torch

from . import _lib


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


def shuffled_order(n: int) -> torch.Tensor:
    """The order in which `DataLoader(TensorDataset(..), shuffle=True)` visits n items in its next epoch, consuming the
    global CPU generator exactly as torch does: one int64 draw for the iterator's base seed
    (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__), one for the RandomSampler's private generator
    (sampler.py, RandomSampler.__iter__), then `randperm(n)` from that generator.  With the same `torch.manual_seed` the
    batches therefore contain the same lists as the reference's loader (attncut_dataloader.py:86-87)."""
    torch.empty((), dtype=torch.int64).random_()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def batch_slices(n_lists: int, batch_size: int, rank: int = 0, world: int = 1, drop_uneven: bool = True):
    """[lo, hi) slices of the epoch order that `rank` collates.  The batch is the unit of sharding (the reference attends
    across the lists of a batch, so a batch must stay whole -- parallel.py): rank r takes batches r, r+W, r+2W, ... of the
    order every rank derives from the same seed, without communication.  `drop_uneven` (world > 1) ends the epoch after
    the last complete round of W batches so that every rank runs the same number of steps (and of all-reduces)."""
    from .parallel import shard_groups
    n_batches = (n_lists + batch_size - 1) // batch_size
    if world > 1 and drop_uneven:
        n_batches -= n_batches % world
    return [(b * batch_size, min(n_lists, (b + 1) * batch_size)) for b in shard_groups(n_batches, rank, world)]


class DeviceLoader:
    """A split held in HBM once -- X [N, L, F] float32, labels [N, L] as float32 or as bit masks (`pack_labels=True`:
    one uint32 per 32 documents) -- iterated like the reference's `data.DataLoader(TensorDataset(X, y), batch_size,
    shuffle=True)`: `for X_b, y_b in loader` yields `[b, L, F]` / `[b, L]` float32 CUDA tensors, the last batch partial,
    `len(loader)` batches per epoch, a new permutation every epoch.  A batch is one rlt_gather_lists launch; the only
    host->device traffic per epoch is the permutation (8 N bytes).  With `world > 1` (one process per GPU) every rank
    holds the split and collates its own batches of the common order (`batch_slices`)."""

    def __init__(self, X: torch.Tensor, y: torch.Tensor, batch_size: int = 20, shuffle: bool = True,
                 pack_labels: bool = False, device=None, rank: int = 0, world: int = 1, drop_uneven: bool = True):
        if X.dim() != 3 or y.dim() != 2 or X.shape[:2] != y.shape:
            raise ValueError(f"expected X [N, L, F] and y [N, L]; got {tuple(X.shape)} and {tuple(y.shape)}")
        if batch_size < 1:
            raise ValueError("batch_size should be a positive integer value")
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceLoader needs a CUDA device (rlt_b200 has no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.batch_size, self.shuffle = int(batch_size), bool(shuffle)
        self.rank, self.world, self.drop_uneven = int(rank), int(world), bool(drop_uneven)
        self.n_lists, self.seq_len, self.n_features = (int(v) for v in X.shape)
        self.X = X.detach().to(self.device, torch.float32).contiguous()
        labels = y.detach().to(self.device, torch.float32).contiguous()
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.y = self.y_bits = None
        if pack_labels:
            words = (self.seq_len + 31) // 32
            self.y_bits = torch.empty(self.n_lists, words, dtype=torch.int32, device=self.device)
            _lib.check(_lib.load().rlt_pack_labels(_lib.ptr(labels), self.n_lists, self.seq_len, _lib.ptr(self.y_bits),
                                                   _lib.ptr(self._status), _lib.stream_ptr()), "rlt_pack_labels")
            if int(self._status.item()) & 2:
                raise ValueError("labels other than 0. and 1. cannot be stored as bit masks")
        else:
            self.y = labels

    def __len__(self) -> int:
        return len(batch_slices(self.n_lists, self.batch_size, self.rank, self.world, self.drop_uneven))

    def gather(self, index: torch.Tensor | None, n_out: int | None = None):
        """(X[index], y[index]) as fresh contiguous CUDA tensors; `index` int64 on the device (None: the first n_out)."""
        n_out = int(index.numel()) if index is not None else int(n_out)
        if index is not None and (index.dtype != torch.int64 or index.device != self.device or not index.is_contiguous()):
            raise ValueError("index must be a contiguous int64 tensor on the loader's device")
        xb = torch.empty(n_out, self.seq_len, self.n_features, dtype=torch.float32, device=self.device)
        yb = torch.empty(n_out, self.seq_len, dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().rlt_gather_lists(_lib.ptr(self.X), _lib.ptr(self.y), _lib.ptr(self.y_bits), _lib.ptr(index),
                                                self.n_lists, n_out, self.seq_len, self.n_features, _lib.ptr(xb), _lib.ptr(yb),
                                                _lib.ptr(self._status), _lib.stream_ptr()), "rlt_gather_lists")
        return xb, yb

    def check(self) -> None:
        """Raises IndexError if any gather since the last check saw an index outside [0, N) (one host sync)."""
        if int(self._status.item()) & 1:
            self._status.zero_()
            raise IndexError("DeviceLoader: list index out of range")

    def __iter__(self):
        order = (shuffled_order(self.n_lists) if self.shuffle else torch.arange(self.n_lists)).to(self.device)
        for lo, hi in batch_slices(self.n_lists, self.batch_size, self.rank, self.world, self.drop_uneven):
            yield self.gather(order[lo:hi])
        self.check()                          # once per epoch, after the last batch: no per-step host sync


def device_loaders(X_train, X_test, y_train, y_test, batch_size: int = 20, pack_labels: bool = False):
    """The (train_loader, test_loader) pair of dataloader/attncut_dataloader.py:72-89 (both shuffled, as there) on
    tensors the caller has already built (the reference reads them from pickles, :29-59)."""
    return (DeviceLoader(X_train, y_train, batch_size, True, pack_labels),
            DeviceLoader(X_test, y_test, batch_size, True, pack_labels))


def rank_tensors(database, dataset_name: str = "bm25", stats: bool = True):
    """(X_train, X_test, y_train, y_test) from the reference's pickle files, as `Rank_Dataset.data_prepare` builds them
    (dataloader/attncut_dataloader.py:21-59 with `stats=True`: X [N, L, 3] = retrieval score + the two neighbour
    statistics of `attncut/{name}_{split}.pkl`; dataloader/choopy_dataloader.py:21-45 with `stats=False`: X [N, L, 1]);
    y [N, L] float32, 1. where the document id is in `gt.pkl[qid]`.  `database` is the reference's
    `DATASET_BASE + '/' + retrieve_data` directory.  Same values as the reference (python floats rounded once to
    float32), built through numpy instead of nested Python lists; every query of a split must have the same length."""
    database = Path(database)

    def load(name):
        with open(database / name, "rb") as f:
            return pickle.load(f)

    gt = {key: set(docs) for key, docs in load("gt.pkl").items()}

    def split(which):
        raw = load(f"{dataset_name}_{which}.pkl")
        extra = load(f"attncut/{dataset_name}_{which}.pkl") if stats else None
        xs, ys = [], []
        for key, ranked in raw.items():
            feats = np.fromiter(ranked.values(), dtype=np.float64, count=len(ranked)).reshape(-1, 1)
            if stats:
                feats = np.column_stack((feats, np.asarray(extra[key], dtype=np.float64)))
            rel = gt[key]                                            # KeyError for an unjudged query, as in the reference
            xs.append(feats)
            ys.append(np.fromiter((1.0 if doc in rel else 0.0 for doc in ranked), dtype=np.float32, count=len(ranked)))
        return torch.from_numpy(np.stack(xs).astype(np.float32)), torch.from_numpy(np.stack(ys))

    (x_tr, y_tr), (x_te, y_te) = split("train"), split("test")
    return x_tr, x_te, y_tr, y_te


def write_synthetic_pickles(database, dataset_name: str = "bm25", n_train: int = 199, n_test: int = 50, seq_len: int = 300,
                            seed: int = 20240229) -> None:
    """Synthetic robust04-shaped data in the reference's on-disk formats, so that the reference's
    own loaders -- and `rank_tensors` -- can be driven without the original corpus: `{name}_{train,test}.pkl`
    (dict qid -> dict doc_id -> score, rank order), `attncut/{name}_{train,test}.pkl` (dict qid -> list[L][2]) and `gt.pkl`
    (dict qid -> list of relevant doc ids)."""
    database = Path(database)
    (database / "attncut").mkdir(parents=True, exist_ok=True)
    x, y = synthetic_lists(n_train + n_test, seq_len, 3, seed=seed)
    gt, q = {}, 0
    for which, n in (("train", n_train), ("test", n_test)):
        raw, extra = {}, {}
        for _ in range(n):
            qid = str(301 + q)
            docs = [f"D{q:04d}-{j:03d}" for j in range(seq_len)]
            raw[qid] = {d: float(v) for d, v in zip(docs, x[q, :, 0].tolist())}
            extra[qid] = x[q, :, 1:].double().tolist()
            gt[qid] = [d for d, r in zip(docs, y[q].tolist()) if r == 1.0] + [f"D{q:04d}-unretrieved"]
            q += 1
        with open(database / f"{dataset_name}_{which}.pkl", "wb") as f:
            pickle.dump(raw, f)
        with open(database / "attncut" / f"{dataset_name}_{which}.pkl", "wb") as f:
            pickle.dump(extra, f)
    with open(database / "gt.pkl", "wb") as f:
        pickle.dump(gt, f)
