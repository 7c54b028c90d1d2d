"""CPU: host logic of the device-resident data path against torch's own DataLoader --
the loader the reference builds in dataloader/attncut_dataloader.py:82-87 -- and the oracle's label bit-mask format."""
import numpy as np
import pytest
import torch
from torch.utils import data

from oracle import rlt_oracle as O
from rlt_b200.data import DeviceLoader, rank_tensors, shuffled_order, synthetic_lists, write_synthetic_pickles


@pytest.mark.parametrize("N,bs", [(23, 5), (249, 64), (64, 64), (7, 20)])
def test_shuffled_order_reproduces_torch_dataloader(N, bs):
    X, y = synthetic_lists(N, 12, 3, seed=N)
    torch.manual_seed(1000 + N)
    loader = data.DataLoader(data.TensorDataset(X, y), batch_size=bs, shuffle=True)
    ref = [[(xb.numpy(), yb.numpy()) for xb, yb in loader] for _ in range(3)]       # three epochs
    torch.manual_seed(1000 + N)
    for epoch in ref:
        got = O.loader_batches(X.numpy(), y.numpy(), bs, shuffled_order(N).numpy())
        assert len(got) == len(epoch) == (N + bs - 1) // bs
        for (gx, gy), (rx, ry) in zip(got, epoch):
            assert np.array_equal(gx, rx) and np.array_equal(gy, ry)
    # both loaders leave the global generator in the same state
    a = torch.rand(1)
    torch.manual_seed(1000 + N)
    for _ in range(3):
        for _ in data.DataLoader(data.TensorDataset(X, y), batch_size=bs, shuffle=True):
            pass
    assert torch.equal(a, torch.rand(1))


@pytest.mark.parametrize("L", [1, 31, 32, 33, 300, 1000])
def test_label_bit_masks_round_trip(L):
    _, y = synthetic_lists(9, L, 1, seed=L)
    bits = O.pack_labels(y.numpy())
    assert bits.shape == (9, (L + 31) // 32) and bits.dtype == np.dtype("<u4")
    assert np.array_equal(O.unpack_labels(bits, L), y.numpy())
    for b, i in ((0, 0), (4, L - 1), (8, L // 2)):
        assert ((int(bits[b, i // 32]) >> (i % 32)) & 1) == int(y[b, i])


def test_device_loader_refuses_cpu_only_host():
    if torch.cuda.is_available():
        pytest.skip("needs a host without CUDA")
    X, y = synthetic_lists(4, 8, 3)
    with pytest.raises(RuntimeError, match="no CPU"):
        DeviceLoader(X, y, batch_size=2)
    with pytest.raises(ValueError):
        DeviceLoader(X, y[:, :4], batch_size=2)


def test_rank_tensors_from_reference_pickle_formats(tmp_path):
    """The reference's on-disk formats -> the tensors its Rank_Dataset builds (attncut_dataloader.py:21-59,
    choopy_dataloader.py:21-45): against the synthetic source, and against the unmodified reference loaders when the
    reference tree is present (this container; not the GPU box)."""
    db = tmp_path / "robust04"
    write_synthetic_pickles(db, "bm25", n_train=7, n_test=3, seq_len=40, seed=11)
    x, y = synthetic_lists(10, 40, 3, seed=11)
    x_tr, x_te, y_tr, y_te = rank_tensors(db, "bm25", stats=True)
    assert x_tr.shape == (7, 40, 3) and x_te.shape == (3, 40, 3) and y_tr.shape == (7, 40) and y_te.dtype == torch.float32
    assert torch.equal(torch.cat([x_tr, x_te]), x) and torch.equal(torch.cat([y_tr, y_te]), y)
    c_tr, c_te, cy_tr, cy_te = rank_tensors(db, "bm25", stats=False)
    assert c_tr.shape == (7, 40, 1) and torch.equal(c_tr, x_tr[:, :, :1]) and torch.equal(cy_te, y_te)

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference tree not present: checked against the synthetic source only")
    import sys
    sys.dont_write_bytecode = True                              # the reference tree is read-only
    ref_dl = refshim._load_package("_rlt_reference_dataloader", refshim.REFERENCE_ROOT / "dataloader")
    import importlib
    for mod_name, stats in (("attncut_dataloader", True), ("choopy_dataloader", False)):
        mod = importlib.import_module(f"_rlt_reference_dataloader.{mod_name}")
        old = mod.DATASET_BASE
        mod.DATASET_BASE = str(tmp_path)                       # a module global of the loaded copy; no file is touched
        try:
            ref = mod.Rank_Dataset("robust04", "bm25")
        finally:
            mod.DATASET_BASE = old
        ours = rank_tensors(db, "bm25", stats=stats)
        for a, b in zip(ours, (ref.getX_train(), ref.getX_test(), ref.gety_train(), ref.gety_test())):
            assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    assert ref_dl.at_dataloader is not None
