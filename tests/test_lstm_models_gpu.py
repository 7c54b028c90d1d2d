"""GPU parity of the BiLSTM op and of the LSTM-based families (BiCut, AttnCut, MtAttnCut, MMOECut) against the
oracle and the reference goldens.  Tolerances: SURVEY.md section 8(c) (see test_encoder_gpu.py)."""
import numpy as np
import pytest
import torch

from helpers import build_model, check_weights, grad_errors, load_golden
from oracle import rlt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,F", [(3, 17, 3), (5, 40, 25), (2, 300, 3), (4, 33, 47)])
def test_bilstm_fwd_bwd_vs_oracle(B, L, F):
    from rlt_b200.autograd import BiLstm
    torch.manual_seed(B * 100 + L)
    lstm = torch.nn.LSTM(input_size=F, hidden_size=128, num_layers=2, batch_first=True, bidirectional=True)
    x = torch.randn(B, L, F)
    dy = torch.randn(B, L, 256) * 0.1
    sd64 = {k: v.detach().double().requires_grad_(True) for k, v in lstm.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    y64 = O.bilstm(x64, sd64, "")
    (y64 * dy.double()).sum().backward()
    lstm = lstm.cuda()
    xc = x.cuda().requires_grad_(True)
    y = BiLstm.apply(xc, 128, 2, *lstm._flat_weights)
    (y * dy.cuda()).sum().backward()
    err = (y.detach().cpu().double() - y64.detach()).abs().max().item()
    assert err <= 1e-3, ("y", err)
    gmax = max(v.grad.abs().max().item() for v in sd64.values())
    for n, p in lstm.named_parameters():
        e = (p.grad.cpu().double() - sd64[n].grad).abs().max().item()
        assert e <= 2e-3 * gmax, (n, e, gmax)
    e = (xc.grad.cpu().double() - x64.grad).abs().max().item()
    assert e <= 2e-3 * x64.grad.abs().max().item(), ("dx", e)


def _criterion(name):
    from utils import losses
    torch.manual_seed(0)
    if name == "bicut":
        return losses.BiCutLoss(metric="f1")
    if name == "attncut":
        return losses.DivLoss(metric="f1", div_type="js", augmented=True)
    if name == "mtattncut":
        return losses.MtCutLoss(metric="f1", rerank_weight=0.5, classi_weight=0.5, num_tasks=3)
    return losses.MtCutLoss(metric="f1", num_tasks=3)


@pytest.mark.parametrize("B", [5, 16])
@pytest.mark.parametrize("name", ["bicut", "attncut", "mtattncut", "mmoecut", "moecut", "plecut"])
def test_lstm_family_vs_reference_golden(name, B):
    g = load_golden(f"model_{name}_B{B}.npz")
    model = build_model(name)
    check_weights(model, g)
    model = model.cuda().train()
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    out = model(x)
    outs = out if isinstance(out, list) else [out]
    assert len(outs) == int(g["n_out"])
    for i, o in enumerate(outs):
        ref = g[f"out{i}"]
        assert tuple(o.shape) == ref.shape
        err = np.abs(o.detach().cpu().numpy() - ref).max()
        assert err <= 1e-3 * np.abs(ref).max(), (name, i, err, np.abs(ref).max())
    last = outs[-1].detach().cpu().numpy()
    if name == "bicut":
        assert [int(k) for k in O.bicut_cut_positions(last)] == [int(k) for k in O.bicut_cut_positions(g["out0"])]
    else:
        assert np.array_equal(np.argmax(last[..., 0], 1), np.argmax(g[f"out{len(outs) - 1}"][..., 0], 1))
    loss = _criterion(name).cuda()(out, y)
    ref_loss = float(g["loss"])
    assert abs(loss.item() - ref_loss) <= 1e-3 * max(abs(ref_loss), 1e-2), (loss.item(), ref_loss)
    loss.backward()
    named = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    rel_l2, rel_max, rel_norm = grad_errors(named, g)
    assert rel_l2 <= 2e-3 and rel_max <= 1e-3, (name, rel_l2, rel_max, rel_norm)
    model.eval()
    with torch.no_grad():
        out2 = model(x)
    o2 = out2[-1] if isinstance(out2, list) else out2
    assert torch.allclose(o2, outs[-1].detach(), rtol=0, atol=0)


@pytest.mark.parametrize("B,L,F", [(3, 17, 3), (130, 40, 3), (257, 23, 25)])
def test_bilstm_tcgen05_recurrence_matches_plain_kernels(B, L, F):
    """The persistent tcgen05 recurrence (fp16 operands, fp32 accumulate) against the plain fp32 kernels, forward
    and backward, including ragged tiles (B not a multiple of 128)."""
    from rlt_b200 import _lib
    from rlt_b200.autograd import BiLstm
    torch.manual_seed(B + L)
    lstm = torch.nn.LSTM(input_size=F, hidden_size=128, num_layers=2, batch_first=True, bidirectional=True).cuda()
    x = torch.randn(B, L, F, device="cuda")
    dy = torch.randn(B, L, 256, device="cuda") * 1e-4
    res = {}
    for be in (1, 0):
        _lib.set_option("lstm_backend", be)
        try:
            lstm.zero_grad(set_to_none=True)
            xc = x.clone().requires_grad_(True)
            y = BiLstm.apply(xc, 128, 2, *lstm._flat_weights)
            (y * dy).sum().backward()
            res[be] = (y.detach().clone(), {n: p.grad.clone() for n, p in lstm.named_parameters()}, xc.grad.clone())
        finally:
            _lib.set_option("lstm_backend", 0)
    y1, g1, dx1 = res[1]
    y0, g0, dx0 = res[0]
    assert (y0 - y1).abs().max().item() <= 1e-3
    gmax = max(v.abs().max().item() for v in g1.values())
    for n in g1:
        assert (g0[n] - g1[n]).abs().max().item() <= 2e-3 * gmax, n
    assert (dx0 - dx1).abs().max().item() <= 2e-3 * dx1.abs().max().item()
