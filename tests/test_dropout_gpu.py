"""Train-mode dropout (GPU).  torch's Philox stream cannot be matched bit for bit, so parity is established with the
SAME masks: rlt_dropout_mask returns exactly the keep-and-scale factors the kernels apply (counter hash of
csrc/dropout.cuh); the oracle's encoder layer / torch ops are evaluated in float64 with those masks and compared with
the kernels' outputs and gradients at the tolerances of the p = 0 tests.  Plus the statistics of the masks."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mask_statistics_and_determinism():
    from rlt_b200 import ops
    n = 1 << 20
    for p in (0.2, 0.4):
        thr = round(p * 65536)
        scale = 65536.0 / (65536 - thr)
        masks = []
        for site in (2, 3, 4, 5):
            m = ops.dropout_mask(1234567, site, p, n)
            vals = torch.unique(m)
            assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - scale) < 1e-6
            keep = (m > 0).float().mean().item()
            assert abs(keep - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5 + 1e-4, (site, p, keep)
            assert abs(m.mean().item() - 1.0) < 5e-3                         # E[mask] = 1
            masks.append(m)
        assert torch.equal(masks[0], ops.dropout_mask(1234567, 2, p, n))     # pure function of (seed, site, index)
        assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[0], ops.dropout_mask(1234568, 2, p, n))
        a = ops.dropout_mask(99, 1, p, 24 * 63 * 63, group_size=63)          # attention site, odd group size
        assert abs((a > 0).float().mean().item() - (1 - p)) < 0.01
    # neighbouring elements are independent (no correlation inside a hash group of 4)
    m = (ops.dropout_mask(7, 3, 0.4, n) > 0).float().view(-1, 4)
    c = torch.corrcoef(m.t())
    assert (c - torch.eye(4, device=c.device)).abs().max().item() < 0.01


def _encoder_layer_errors(B, L, d, n_head, p):
    """Kernel vs float64 oracle (with the kernel's own masks when p > 0): output error and gradient errors."""
    from oracle import rlt_oracle as O
    from rlt_b200 import ops
    from rlt_b200.autograd import EncoderStack, fresh_seed, reset_dropout_seed
    from models.truncation import _encoder_params
    torch.manual_seed(5)
    enc = torch.nn.TransformerEncoder(torch.nn.TransformerEncoderLayer(d_model=d, nhead=n_head, dropout=p), 1,
                                      enable_nested_tensor=False)
    f = enc.layers[0].linear1.out_features
    x = torch.randn(B, L, d)
    gout = torch.randn(B, L, d)
    # ---- kernels, with a known seed (the Function draws it from torch's host generator)
    torch.manual_seed(77)
    reset_dropout_seed()
    seed = fresh_seed()
    reset_dropout_seed()      # the forward below draws the same seed again
    enc_c = enc.cuda()
    xc = x.cuda().requires_grad_(True)
    out = EncoderStack.apply(xc, n_head, 1, 1e-5, p, *_encoder_params(enc_c))
    out.backward(gout.cuda())
    # ---- the masks the kernels used
    T = B * L
    masks = None
    if p > 0:
        masks = {"attn": ops.dropout_mask(seed, 1, p, L * n_head * B * B, group_size=B).view(L, n_head, B, B),
                 "after_attn": ops.dropout_mask(seed, 2, p, T * d).view(B, L, d),
                 "ffn": ops.dropout_mask(seed, 3, p, T * f).view(B, L, f),
                 "after_ffn": ops.dropout_mask(seed, 4, p, T * d).view(B, L, d)}
        masks = {k: v.double().cpu() for k, v in masks.items()}
    # ---- oracle in float64 with the same masks
    sd = {"layers.0." + k: v.detach().double().cpu().requires_grad_(True) for k, v in enc.layers[0].state_dict().items()}
    x64 = x.double().requires_grad_(True)
    ref = O.encoder_layer(x64, sd, "layers.0.", n_head, masks=masks)
    ref.backward(gout.double())
    err = (out.detach().cpu().double() - ref.detach()).abs().max().item() / ref.abs().max().item()
    pairs = {"x": (xc.grad.cpu().double(), x64.grad)}
    for name, prm in enc_c.layers[0].named_parameters():
        pairs[name] = (prm.grad.cpu().double(), sd["layers.0." + name].grad)
    num = sum(((a - b) ** 2).sum() for a, b in pairs.values()).sqrt().item()
    den = sum((b ** 2).sum() for _, b in pairs.values()).sqrt().item()
    per = {n: ((a - b).norm() / b.norm()).item() for n, (a, b) in pairs.items()}
    return err, num / den, per


@pytest.mark.parametrize("B,L,d,n_head,p", [(16, 40, 128, 8, 0.2), (7, 33, 256, 4, 0.4), (63, 12, 128, 8, 0.2), (100, 8, 128, 8, 0.4),
                                            # the generic attention kernels: PLECut's head dim 128 (PLECut.py:57, n_head = 2)
                                            # and a group of more than 128 lists
                                            (9, 20, 256, 2, 0.3), (130, 6, 128, 8, 0.2)])
def test_encoder_layer_dropout_matches_oracle_with_the_same_masks(B, L, d, n_head, p):
    """With a random upstream gradient the gradient error of ANY reduced-precision forward is dominated by ReLU gates
    whose pre-activation rounds across zero (relative L2 ~ sqrt(flip fraction), ~1.5e-2 on linear1.*, for TF32 and
    fp16 alike), so the reference point is the SAME kernel at p = 0: dropout must not add error beyond it, and the
    output itself must meet the usual 1e-3."""
    err0, g0, per0 = _encoder_layer_errors(B, L, d, n_head, 0.0)
    err, g, per = _encoder_layer_errors(B, L, d, n_head, p)
    assert err0 <= 1e-3 and err <= 1e-3, (err0, err)
    assert g <= 1.5 * g0 + 1e-3, (g, g0)
    for n in per:
        assert per[n] <= 2.0 * per0[n] + 3e-3, (n, per[n], per0[n])     # a wrong mask would be an O(p) error
    # biases that sit inside a dropout see the MASKED upstream gradient (an unmasked sum would be off by ~p)
    assert per["linear2.bias"] <= 1e-3 and per["self_attn.out_proj.bias"] <= 1e-2, per


def test_bicut_logit_dropout_matches_torch_with_the_same_mask():
    from rlt_b200 import ops
    from rlt_b200.autograd import BicutHead, fresh_seed, reset_dropout_seed
    torch.manual_seed(3)
    B, L, p = 9, 50, 0.4
    h = torch.randn(B, L, 256)
    fc = torch.nn.Linear(256, 256)
    cls = torch.nn.Linear(256, 2)
    gout = torch.randn(B, L, 2)
    torch.manual_seed(11)
    reset_dropout_seed()
    seed = fresh_seed()
    reset_dropout_seed()
    hc = h.cuda().requires_grad_(True)
    fcc, clsc = fc.cuda(), cls.cuda()
    o = BicutHead.apply(hc, fcc.weight, fcc.bias, clsc.weight, clsc.bias, p)
    o.backward(gout.cuda())
    T = B * L
    m = ops.dropout_mask(seed, 5, p, 2 * T).view(2, T).t().reshape(B, L, 2).double().cpu()   # planes [2, T] -> [B, L, 2]
    h64 = h.double().requires_grad_(True)
    w1, b1 = fc.weight.detach().double().cpu().requires_grad_(True), fc.bias.detach().double().cpu().requires_grad_(True)
    w2, b2 = cls.weight.detach().double().cpu().requires_grad_(True), cls.bias.detach().double().cpu().requires_grad_(True)
    u = torch.relu(h64 @ w1.t() + b1) @ w2.t() + b2
    ref = torch.softmax(u * m, dim=2)
    ref.backward(gout.double())
    assert (o.detach().cpu().double() - ref.detach()).abs().max().item() <= 1e-3
    # downstream of the dropout the mask must be applied exactly; upstream of the fc ReLU the TF32 forward flips a few
    # gates whose pre-activation rounds across zero (relative L2 ~ 1e-2, the same as without dropout)
    for a, b, tol in ((clsc.weight.grad, w2.grad, 2e-3), (clsc.bias.grad, b2.grad, 2e-3), (hc.grad, h64.grad, 2e-2),
                      (fcc.weight.grad, w1.grad, 2e-2)):
        assert ((a.cpu().double() - b).norm() / b.norm()).item() <= tol


@pytest.mark.parametrize("name", ["choopy", "bicut", "attncut", "mmoecut", "plecut"])
def test_modules_train_with_reference_default_dropout(name):
    """run.py trains with the constructors' default dropout (0.2 / 0.4): forward + backward work in train(), masks
    change from call to call, eval() is deterministic and dropout-free."""
    import models
    from utils import losses
    from rlt_b200.data import synthetic_lists
    torch.manual_seed(0)
    if name == "choopy":
        model, crit, F = models.Choopy(seq_len=60), losses.ChoopyLoss(), 1
    elif name == "bicut":
        model, crit, F = models.BiCut(input_size=3), losses.BiCutLoss(metric="f1"), 3
    elif name == "attncut":
        model, crit, F = models.AttnCut(input_size=3), losses.DivLoss(metric="f1", div_type="js"), 3
    elif name == "plecut":      # run.py 'mtple': n_head = 2 -> head dim 128 -> the generic attention kernels
        model, crit, F = models.PLECut(seq_len=60, input_size=3), losses.MtCutLoss(metric="f1"), 3
    else:
        model, crit, F = models.MMOECut(seq_len=60, input_size=3), losses.MtCutLoss(metric="f1"), 3
    model = model.cuda().train()
    x, y = synthetic_lists(12, 60, F, seed=3, device="cuda")
    out1 = model(x)
    loss = crit(out1, y)
    loss.backward()
    assert torch.isfinite(loss)
    for n_, p_ in model.named_parameters():
        if p_.grad is None:     # only the rerank head / tower / gate may lack a gradient (inactive hinge, losses.py:141)
            assert "rerank" in n_ or n_ == "w_gates.1", n_
            continue
        assert torch.isfinite(p_.grad).all(), n_
    first = (out1[-1] if isinstance(out1, list) else out1).detach()
    second = model(x)
    second = (second[-1] if isinstance(second, list) else second).detach()
    assert not torch.equal(first, second)                      # fresh masks per call
    model.eval()
    with torch.no_grad():
        a = model(x)
        b = model(x)
    a = a[-1] if isinstance(a, list) else a
    b = b[-1] if isinstance(b, list) else b
    assert torch.equal(a, b)


def test_engine_and_module_agree_under_dropout():
    """The throughput engine and the nn.Module path draw their mask seeds from the same host generator."""
    import models
    from utils import losses
    from rlt_b200.data import synthetic_lists
    from rlt_b200.engine import Engine
    torch.manual_seed(1)
    model = models.Choopy(seq_len=300, dropout=0.2).cuda().train()
    x, y = synthetic_lists(16, 300, 1, seed=5, device="cuda")
    from rlt_b200.autograd import reset_dropout_seed
    torch.manual_seed(42)
    reset_dropout_seed()
    loss = losses.ChoopyLoss()(model(x), y)
    loss.backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters()}
    eng = Engine(model, n_groups=1, group_size=16, seq_len=300)
    torch.manual_seed(42)
    reset_dropout_seed()
    eng_loss = eng.train_step(x, y)
    assert abs(eng_loss.item() - loss.item()) <= 1e-5 * max(1.0, abs(loss.item()))
    for n in ("decison_layer.0.weight", "attention_layer.layers.0.linear1.weight", "attention_layer.layers.2.self_attn.in_proj_weight"):
        assert torch.allclose(eng.grads[n], ref[n], rtol=2e-3, atol=1e-7), n
