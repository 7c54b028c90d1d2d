"""GPU diagnostic: rlt_attention_lists_fwd (tcgen05 vs mma.sync path) against a float64 torch evaluation."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import _lib, ops  # noqa: E402


def ref(qkv, G, S, L, d, nh):
    dh = d // nh
    x = qkv.double().view(G, S, L, 3, nh, dh)
    q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]          # [G, S, L, nh, dh]
    q, k, v = (t.permute(0, 2, 3, 1, 4) for t in (q, k, v))         # [G, L, nh, S, dh]
    s = q @ k.transpose(-1, -2) / dh ** 0.5
    o = torch.softmax(s, -1) @ v                                    # [G, L, nh, S, dh]
    lse = torch.logsumexp(s, -1)                                    # [G, L, nh, S]
    return o.permute(0, 3, 1, 2, 4).reshape(G * S * L, d), lse.permute(0, 3, 1, 2).reshape(G * S * L, nh)


for (G, S, L) in ((1, 64, 2), (1, 64, 1), (2, 64, 20), (1, 5, 41), (3, 33, 7)):
    d, nh = 128, 8
    torch.manual_seed(G + S + L)
    qkv = torch.randn(G * S * L, 3 * d, device="cuda")
    ro, rl = ref(qkv, G, S, L, d, nh)
    for tc in (0, 1):
        _lib.set_option("attn_tc", tc)
        o, lse = ops.attention_lists_fwd(qkv, G, S, L, d, nh)
        torch.cuda.synchronize()
        eo = (o.double() - ro).abs()
        el = (lse.double() - rl).abs()
        print(f"G={G} S={S} L={L} tc={tc}: max|do| {eo.max().item():.3e} (max|o| {ro.abs().max().item():.2f})  max|dlse| {el.max().item():.3e}")
        if tc == 1 and eo.max().item() > 1e-2:
            bad = (eo > 1e-2).nonzero()
            toks = bad[:, 0].unique()
            print("   bad tokens:", toks[:16].tolist(), "... n =", toks.numel(), " bad cols:", bad[:, 1].unique()[:20].tolist())
    _lib.set_option("attn_tc", 1)
