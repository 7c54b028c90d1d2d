"""CPU: host logic of the device-resident data path against torch's own DataLoader --
the loader the reference builds in dataloader/attncut_dataloader.py:82-87 -- and the oracle's label bit-mask format."""
import numpy as np
import pytest
import torch
from torch.utils import data

from oracle import rlt_oracle as O
from rlt_b200.data import DeviceLoader, shuffled_order, synthetic_lists


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
