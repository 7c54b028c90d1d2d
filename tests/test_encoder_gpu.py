"""GPU parity of the encoder layer (fwd + bwd) and of the Choopy-family models against the oracle and
the reference goldens.  Tolerances follow SURVEY.md section 8(c): TF32 tensor-core contractions with fp32
accumulate -> outputs max|d| <= 1e-3 * max|ref|, loss rel <= 1e-3, gradients global rel-L2 <= 2e-3 and
max|d| <= 1e-3 * max over all tensors |g_ref|."""
import numpy as np
import pytest
import torch

from helpers import build_model, check_weights, grad_errors, load_golden
from oracle import rlt_oracle as O

pytestmark = pytest.mark.gpu


def _layer_sd(d, n_head, seed):
    import warnings
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        layer = torch.nn.TransformerEncoderLayer(d_model=d, nhead=n_head, dropout=0.0)
    with torch.no_grad():   # make the affine LN parameters non-trivial
        for n, p in layer.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    return {k: v.detach().clone() for k, v in layer.state_dict().items()}


@pytest.mark.parametrize("backend", [1, 0])
@pytest.mark.parametrize("d,n_head,S,L,G", [(128, 8, 5, 40, 1), (128, 8, 16, 300, 1), (256, 4, 7, 33, 2),
                                            (256, 4, 64, 40, 1), (128, 8, 64, 20, 2),
                                            # the reference's default batch (63), its test batch (49), a group > 64 lists
                                            (128, 8, 63, 20, 2), (256, 4, 49, 24, 1), (128, 8, 100, 12, 1)])
def test_encoder_layer_fwd_bwd_vs_oracle(backend, d, n_head, S, L, G):
    from rlt_b200 import _lib, ops
    from rlt_b200.autograd import EncoderStack
    _lib.set_option("gemm_backend", backend)
    try:
        sd = _layer_sd(d, n_head, seed=d + S)
        torch.manual_seed(7)
        x = torch.randn(G * S, L, d)
        dy = torch.randn(G * S, L, d) * 0.01
        # oracle in float64, group by group
        sd64 = {("layers.0." + k): v.double().requires_grad_(True) for k, v in sd.items()}
        x64 = x.double().requires_grad_(True)
        out64 = torch.cat([O.encoder_stack(x64[g * S:(g + 1) * S], sd64, "", n_head) for g in range(G)])
        (out64 * dy.double()).sum().backward()
        params = [sd[n].cuda().requires_grad_(True) for n in ops.ENCODER_PARAM_ORDER]
        xc = x.cuda().requires_grad_(True)
        out = EncoderStack.apply(xc, n_head, G, 1e-5, 0.0, *params)
        (out * dy.cuda()).sum().backward()
        # backend 1 (exact fp32 FMA) pins the kernel logic at 2e-5.  backend 0 (TF32 tensor cores): outputs at
        # 1e-3; for gradients this synthetic test (random upstream gradient, a few hundred tokens) only asserts a
        # smoke-level bound, because TF32 flips the ReLU mask of ~3e-4 of the hidden units whose pre-activation is
        # ~0, which alone is a ~1e-2 relative error on linear1.weight / linear1.bias (SURVEY.md section 7 item 4
        # measured the same on emulated TF32).  The 1e-3 contract of section 8(c) is asserted on the real models.
        tol = 1e-3 if backend == 0 else 2e-5
        err = (out.detach().cpu().double() - out64.detach()).abs().max().item()
        assert err <= tol * out64.abs().max().item(), ("out", err)
        gtol = 4e-2 if backend == 0 else 4e-5
        gmax = max(sd64["layers.0." + n].grad.abs().max().item() for n in ops.ENCODER_PARAM_ORDER)
        for n, p in zip(ops.ENCODER_PARAM_ORDER, params):
            ref = sd64["layers.0." + n].grad
            e = (p.grad.cpu().double() - ref).abs().max().item()
            assert e <= gtol * gmax, (n, e, gmax)
        e = (xc.grad.cpu().double() - x64.grad).abs().max().item()
        assert e <= gtol * x64.grad.abs().max().item(), ("dx", e)
    finally:
        _lib.set_option("gemm_backend", 0)


@pytest.mark.parametrize("B", [5, 16])
@pytest.mark.parametrize("name", ["choopy", "mtchoopy"])
def test_choopy_family_vs_reference_golden(name, B):
    from utils import losses
    g = load_golden(f"model_{name}_B{B}.npz")
    model = build_model(name)
    check_weights(model, g)
    model = model.cuda().train()
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    out = model(x)
    outs = out if isinstance(out, list) else [out]
    for i, o in enumerate(outs):
        ref = g[f"out{i}"]
        assert tuple(o.shape) == ref.shape
        err = np.abs(o.detach().cpu().numpy() - ref).max()
        assert err <= 1e-3 * np.abs(ref).max(), (name, i, err, np.abs(ref).max())
    # cut positions identical (reference margin is far above the tolerance on these fixtures)
    assert np.array_equal(np.argmax(outs[-1].detach().cpu().numpy()[..., 0], 1), np.argmax(g[f"out{len(outs)-1}"][..., 0], 1))
    torch.manual_seed(0)
    crit = (losses.ChoopyLoss(metric="f1") if name == "choopy" else
            losses.MtCutLoss(metric="f1", rerank_weight=0.5, classi_weight=0.5, num_tasks=3)).cuda()
    loss = crit(out, y)
    ref_loss = float(g["loss"])
    assert abs(loss.item() - ref_loss) <= 1e-3 * max(abs(ref_loss), 1e-2), (loss.item(), ref_loss)
    loss.backward()
    named = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    rel_l2, rel_max, rel_norm = grad_errors(named, g)
    assert rel_l2 <= 2e-3 and rel_max <= 1e-3, (name, rel_l2, rel_max, rel_norm)
    # eval mode under no_grad (run.py:166-167) gives the same outputs
    model.eval()
    with torch.no_grad():
        out2 = model(x)
    o2 = out2[-1] if isinstance(out2, list) else out2
    assert torch.equal(o2, outs[-1].detach())


@pytest.mark.parametrize("S,L,G", [(64, 20, 2), (63, 9, 1), (16, 300, 1)])
def test_fused_backward_paths_match_unfused(S, L, G):
    """The one-pass FFN backward (ffn_bwd_kernel: dH + db1 + dW2) and the bias-gradient column sums folded into
    gemm_dw must reproduce the separate kernels they replace: same fp16 operands and fp32 accumulation, only the
    summation order over tokens differs (TF32 rounding of dQKV in the column-sum product: 2^-11 per element)."""
    from rlt_b200 import _lib, ops
    from rlt_b200.autograd import EncoderStack
    d, n_head = 128, 8
    sd = _layer_sd(d, n_head, seed=3)
    torch.manual_seed(11)
    x = torch.randn(G * S, L, d)
    dy = torch.randn(G * S, L, d) * 0.01

    def run(fused, dw_colsum):
        _lib.set_option("ffn_bwd_fused", fused)
        _lib.set_option("dw_colsum", dw_colsum)
        try:
            params = [sd[n].cuda().requires_grad_(True) for n in ops.ENCODER_PARAM_ORDER]
            xc = x.cuda().requires_grad_(True)
            out = EncoderStack.apply(xc, n_head, G, 1e-5, 0.0, *params)
            (out * dy.cuda()).sum().backward()
            return [p.grad.double().cpu() for p in params] + [xc.grad.double().cpu()]
        finally:
            _lib.set_option("ffn_bwd_fused", 1)
            _lib.set_option("dw_colsum", 1)

    ref = run(0, 0)
    new = run(1, 1)
    names = list(ops.ENCODER_PARAM_ORDER) + ["dx"]
    for n, a, b in zip(names, ref, new):
        scale = a.abs().max().item() + 1e-30
        e = (a - b).abs().max().item()
        tol = 5e-4 if n == "self_attn.in_proj_bias" else 2e-5      # in_proj_bias: exact fp32 sums vs a TF32-operand product
        assert e <= tol * scale, (n, e, scale)


@pytest.mark.parametrize("d,n_head", [(128, 8), (256, 4)])
@pytest.mark.parametrize("S,L,G", [(5, 40, 1), (64, 20, 2), (63, 9, 1), (16, 300, 1), (64, 300, 3), (1, 1, 1), (3, 171, 1)])
def test_fused_ffn_forward_matches_the_three_kernel_path(S, L, G, d, n_head):
    """ffn_fwd_kernel (FFN1 + ReLU + FFN2 + residual + LayerNorm2 in one cta_group::2 kernel, hidden in tensor memory)
    against the path it replaces (FFN1 GEMM -> fp16 hidden in HBM -> FFN2 GEMM -> LayerNorm kernel): same fp16
    operands, fp32 accumulation in tensor memory, only the summation order of the LayerNorm statistics differs.
    Token counts include partial 256-token tiles, a single token, and several tiles per CTA pair; the saved hidden /
    pre-norm sum / statistics are checked through the (unchanged) backward."""
    from rlt_b200 import _lib, ops
    from rlt_b200.autograd import EncoderStack
    sd = _layer_sd(d, n_head, seed=3)
    torch.manual_seed(11 + S)
    x = torch.randn(G * S, L, d)
    dy = torch.randn(G * S, L, d) * 0.01

    def run(fused, grad=True):
        _lib.set_option("ffn_fwd_fused", fused)
        try:
            params = [sd[n].cuda().requires_grad_(grad) for n in ops.ENCODER_PARAM_ORDER]
            xc = x.cuda().requires_grad_(grad)
            out = EncoderStack.apply(xc, n_head, G, 1e-5, 0.0, *params)
            if not grad:
                return [out.detach().double().cpu()]
            (out * dy.cuda()).sum().backward()
            return [out.detach().double().cpu()] + [p.grad.double().cpu() for p in params] + [xc.grad.double().cpu()]
        finally:
            _lib.set_option("ffn_fwd_fused", 1)

    ref = run(0)
    new = run(1)
    names = ["out"] + list(ops.ENCODER_PARAM_ORDER) + ["dx"]
    for n, a, b in zip(names, ref, new):
        scale = a.abs().max().item() + 1e-30
        e = (a - b).abs().max().item()
        # the output itself at fp32 rounding; the gradients pass through TF32 / fp16 operand rounding downstream, where a
        # last-bit difference of an activation can move an operand by one TF32 ulp (2^-11)
        assert e <= (2e-5 if n == "out" else 5e-4) * scale, (n, e, scale)
    # forward only (torch.no_grad / eval): the hidden is not written at all, the output is the same
    inf = run(1, grad=False)[0]
    assert (inf - new[0]).abs().max().item() <= 1e-6 * new[0].abs().max().item()


@pytest.mark.parametrize("S,L,G", [(64, 20, 2), (63, 9, 1), (5, 41, 1), (16, 300, 1), (64, 301, 2), (1, 1, 1), (33, 7, 3)])
def test_tcgen05_attention_forward_matches_the_mma_sync_path(S, L, G):
    """attn_lists_fwd_tc_kernel (tcgen05: two positions per 128-row tile, fp16 hi / lo score operands, P through tensor
    memory) against the mma.sync kernel it replaces for head dim 16 and groups of <= 64 lists: layer output and -- through
    the unchanged backward, which consumes the saved attention output and log-sum-exp -- every gradient.  Shapes cover
    partial groups (keys masked), odd sequence lengths (half-empty position pair) and several groups."""
    from rlt_b200 import _lib, ops
    from rlt_b200.autograd import EncoderStack
    d, n_head = 128, 8
    sd = _layer_sd(d, n_head, seed=5)
    torch.manual_seed(17 + S)
    x = torch.randn(G * S, L, d)
    dy = torch.randn(G * S, L, d) * 0.01

    def run(tc):
        _lib.set_option("attn_tc", tc)
        try:
            params = [sd[n].cuda().requires_grad_(True) for n in ops.ENCODER_PARAM_ORDER]
            xc = x.cuda().requires_grad_(True)
            out = EncoderStack.apply(xc, n_head, G, 1e-5, 0.0, *params)
            (out * dy.cuda()).sum().backward()
            return [out.detach().double().cpu()] + [p.grad.double().cpu() for p in params] + [xc.grad.double().cpu()]
        finally:
            _lib.set_option("attn_tc", 1)

    ref = run(0)
    new = run(1)
    names = ["out"] + list(ops.ENCODER_PARAM_ORDER) + ["dx"]
    gmax = max(a.abs().max().item() for a in ref[1:-1])
    for n, a, b in zip(names, ref, new):
        e = (a - b).abs().max().item()
        # P and V enter the second product with 11 significant bits on both paths (fp16 here, TF32 there): the outputs
        # agree to that rounding, far inside the 1e-3 contract checked against the oracle elsewhere in this file.  The
        # gradients get the smoke-level bound of test_encoder_layer_fwd_bwd_vs_oracle (4e-2 of the largest gradient
        # entry): with a random upstream gradient a last-bit change of y flips ReLU gates of the FFN (relative 1e-2 on
        # linear1.*); the 1e-3 contract is asserted on the model goldens
        scale = a.abs().max().item() + 1e-30 if n in ("out", "dx") else gmax
        assert e <= (3e-4 if n == "out" else 4e-2) * scale, (n, e, scale)


@pytest.mark.parametrize("d", [128, 256])
@pytest.mark.parametrize("f", [128, 256, 384, 640, 2048])
@pytest.mark.parametrize("T", [1, 300, 5000, 70001])
def test_fused_ffn_forward_hidden_widths_and_token_counts(d, f, T):
    """rlt_ffn_fused_fwd on its own against a float64 evaluation with the same fp16-rounded operands, for hidden widths
    of 1, 2, 3, 5 and 16 (d 128) / 2 ... 32 (d 256) chunks -- fewer chunks than S buffers, ring wrap-arounds at every
    phase -- and token counts from one row to several tiles per CTA pair with a ragged last tile.  Both variants: forward
    only, and with the hidden / pre-norm sums / statistics saved."""
    from rlt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(d + f + T)
    y = torch.randn(T, d, device="cuda", generator=g)
    w1 = torch.randn(f, d, device="cuda", generator=g) * d ** -0.5
    w2 = torch.randn(d, f, device="cuda", generator=g) * f ** -0.5
    b1 = torch.randn(f, device="cuda", generator=g) * 0.1
    b2 = torch.randn(d, device="cuda", generator=g) * 0.1
    gamma = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    beta = 0.1 * torch.randn(d, device="cuda", generator=g)
    y16, w1h, w2h = y.half(), w1.half().contiguous(), w2.half().contiguous()
    n = min(T, 2048)
    rows = torch.cat([torch.arange(n // 2), torch.arange(T - (n - n // 2), T)]).unique().cuda()      # first and last rows
    hd = torch.relu(y16[rows].double() @ w1h.double().t() + b1.double()).half().double()
    u = y[rows].double() + hd @ w2h.double().t() + b2.double()
    ref = torch.nn.functional.layer_norm(u, (d,), gamma.double(), beta.double(), 1e-5)
    out = torch.full((T, d), float("nan"), device="cuda")
    ops.ffn_fused_fwd(y16, y, w1h, b1, w2h, b2, gamma, beta, out)
    assert (out[rows].double() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    assert torch.isfinite(out).all()
    out2 = torch.full((T, d), float("nan"), device="cuda")
    u2 = torch.empty_like(y)
    st = torch.empty(T, 2, device="cuda")
    h = torch.empty(T, f, device="cuda", dtype=torch.float16)
    ops.ffn_fused_fwd(y16, y, w1h, b1, w2h, b2, gamma, beta, out2, u2=u2, stats=st, h_out=h)
    assert torch.equal(out, out2)                                             # the two variants run the same arithmetic
    assert (h[rows].double() - hd).abs().max().item() <= 2e-3 * max(hd.abs().max().item(), 1.0)
    assert (u2[rows].double() - u).abs().max().item() <= 2e-4 * u.abs().max().item()
    mu, var = u.mean(1), u.var(1, unbiased=False)
    assert (st[rows, 0].double() - mu).abs().max().item() <= 1e-4 * max(mu.abs().max().item(), 1.0)
    assert (st[rows, 1].double() - (var + 1e-5).rsqrt()).abs().max().item() <= 1e-3 * (var + 1e-5).rsqrt().max().item()
