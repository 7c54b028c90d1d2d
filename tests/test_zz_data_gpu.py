"""GPU parity of the device-resident data path: DeviceLoader yields, bit for bit, the
batches of the torch DataLoader the reference builds (dataloader/attncut_dataloader.py:82-87) under the same seed."""
import numpy as np
import pytest
import torch
from torch.utils import data

from oracle import rlt_oracle as O
from rlt_b200.data import DeviceLoader, device_loaders, synthetic_lists

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pack", [False, True])
@pytest.mark.parametrize("N,L,F,bs", [(249, 300, 3, 64), (100, 300, 1, 63), (37, 41, 3, 8), (5, 1000, 47, 20), (2000, 300, 3, 512)])
def test_device_loader_matches_torch_dataloader(N, L, F, bs, pack):
    X, y = synthetic_lists(N, L, F, seed=N + L)
    torch.manual_seed(N)
    ref = [[(xb, yb) for xb, yb in data.DataLoader(data.TensorDataset(X, y), batch_size=bs, shuffle=True)] for _ in range(2)]
    loader = DeviceLoader(X, y, batch_size=bs, shuffle=True, pack_labels=pack)
    assert len(loader) == len(ref[0])
    if pack:
        assert loader.y is None and np.array_equal(loader.y_bits.cpu().numpy().view(np.uint32), O.pack_labels(y.numpy()))
    torch.manual_seed(N)
    for epoch in ref:
        got = list(loader)
        assert len(got) == len(epoch)
        for (gx, gy), (rx, ry) in zip(got, epoch):
            assert gx.is_cuda and gx.dtype == torch.float32 and gx.is_contiguous() and gy.is_contiguous()
            assert torch.equal(gx.cpu(), rx) and torch.equal(gy.cpu(), ry)


def test_device_loader_edges():
    X, y = synthetic_lists(10, 300, 3, seed=3)
    loader = DeviceLoader(X, y, batch_size=4, shuffle=False)
    got = list(loader)
    assert [int(b[0].shape[0]) for b in got] == [4, 4, 2]
    assert torch.equal(torch.cat([b[0] for b in got]).cpu(), X) and torch.equal(torch.cat([b[1] for b in got]).cpu(), y)
    xb, yb = loader.gather(None, 3)
    assert torch.equal(xb.cpu(), X[:3]) and torch.equal(yb.cpu(), y[:3])
    idx = torch.tensor([9, 0, 9, 3], device="cuda")
    xb, yb = loader.gather(idx)
    assert torch.equal(xb.cpu(), X[idx.cpu()]) and torch.equal(yb.cpu(), y[idx.cpu()])
    loader.gather(torch.tensor([1, 10], device="cuda"))
    with pytest.raises(IndexError):
        loader.check()
    loader.check()                                        # the flag was cleared
    with pytest.raises(ValueError):
        loader.gather(torch.tensor([1], device="cuda", dtype=torch.int32))
    with pytest.raises(ValueError, match="bit masks"):
        DeviceLoader(X, y * 0.5, batch_size=4, pack_labels=True)
    tr, te = device_loaders(X[:6], X[6:], y[:6], y[6:], batch_size=4, pack_labels=True)
    assert len(tr) == 2 and len(te) == 1 and next(iter(te))[1].shape == (4, 300)


def test_device_loader_feeds_a_model_step():
    """The batches are what the kernels read: one Choopy train step straight from the loader."""
    import models
    from utils import losses
    X, y = synthetic_lists(12, 300, 1, seed=5)
    torch.manual_seed(0)
    model = models.Choopy(seq_len=300, dropout=0.0).cuda()
    loader = DeviceLoader(X, y, batch_size=8, pack_labels=True)
    xb, yb = next(iter(loader))
    loss = losses.ChoopyLoss(metric="f1")(model(xb), yb)
    loss.backward()
    assert torch.isfinite(loss).item() and model.decison_layer[0].weight.grad is not None
