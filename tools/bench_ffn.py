"""GPU: time the fused feed-forward kernel (rlt_ffn_fused_fwd) alone, forward-only and with the hidden saved, against
its tensor-pipe and HBM rooflines; check it against a float64 torch evaluation of the same block."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import ops  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 300
    d, f = (int(sys.argv[2]) if len(sys.argv) > 2 else 128), 2048
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm, tfs = pk.get("hbm_gbs", 6650.0), pk.get("bf16_tflops_sustained", 1400.0)
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn(T, d, device="cuda", generator=g)
    w1 = torch.randn(f, d, device="cuda", generator=g) * d ** -0.5
    w2 = torch.randn(d, f, device="cuda", generator=g) * f ** -0.5
    b1 = torch.randn(f, device="cuda", generator=g) * 0.1
    b2 = torch.randn(d, device="cuda", generator=g) * 0.1
    gamma = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    beta = 0.1 * torch.randn(d, device="cuda", generator=g)
    y16, w1h, w2h = y.half(), w1.half().contiguous(), w2.half().contiguous()
    out = torch.empty_like(y)
    u2 = torch.empty_like(y)
    st = torch.empty(T, 2, device="cuda")
    h = torch.empty(T, f, device="cuda", dtype=torch.float16)
    # parity on a slice: float64 evaluation with the same fp16-rounded operands
    n = min(T, 4096)
    ops.ffn_fused_fwd(y16, y, w1h, b1, w2h, b2, gamma, beta, out, u2, st, h)
    hd = torch.relu(y16[:n].double() @ w1h.double().t() + b1.double()).half().double()
    u = y[:n].double() + hd @ w2h.double().t() + b2.double()
    ref = torch.nn.functional.layer_norm(u, (d,), gamma.double(), beta.double(), 1e-5)
    print("max |out - ref| / max|ref| =", ((out[:n].double() - ref).abs().max() / ref.abs().max()).item(),
          " hidden max err", (h[:n].double() - hd).abs().max().item(), " u2 err", (u2[:n].double() - u).abs().max().item())
    flops = 4.0 * T * d * f
    for label, kw in (("forward only", {}), ("hidden saved", dict(u2=u2, stats=st, h_out=h))):
        fn = lambda: ops.ffn_fused_fwd(y16, y, w1h, b1, w2h, b2, gamma, beta, out, **kw)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        nbytes = T * (d * 2 + 2 * d * 4) + 2 * d * f * 2 + (T * (f * 2 + d * 4 + 8) if kw else 0)
        print(f"{label:13s}: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s ({flops / ms / 1e9 / tfs:.2f} of sustained bf16)  "
              f"{nbytes / ms / 1e6:7.0f} GB/s ({nbytes / ms / 1e6 / hbm:.2f} of copy peak)")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "--timeline"):
    main()


def timeline(train=False):
    """Per-chunk clock64 stamps of pair 0 (leader CTA): where the MMA thread and the first epilogue warp spend a chunk."""
    import ctypes as C
    T, d, f = 256 * 74 * 4, 128, 2048
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn(T, d, device="cuda", generator=g)
    w1h = (torch.randn(f, d, device="cuda", generator=g) * d ** -0.5).half()
    w2h = (torch.randn(d, f, device="cuda", generator=g) * f ** -0.5).half()
    z = torch.zeros(f, device="cuda")
    zd = torch.zeros(d, device="cuda")
    out = torch.empty_like(y)
    buf = torch.zeros(64 * 24, dtype=torch.int64, device="cuda")
    lib = ops.lib()
    kw = dict(u2=torch.empty_like(y), stats=torch.empty(T, 2, device="cuda"),
              h_out=torch.empty(T, f, device="cuda", dtype=torch.float16)) if train else {}
    for _ in range(2):
        ops.ffn_fused_fwd(y.half(), y, w1h, z, w2h, zd, zd + 1, zd, out, **kw)
    lib.rlt_ffn_fused_set_timeline(C.c_void_p(buf.data_ptr()))
    ops.ffn_fused_fwd(y.half(), y, w1h, z, w2h, zd, zd + 1, zd, out, **kw)
    torch.cuda.synchronize()
    lib.rlt_ffn_fused_set_timeline(C.c_void_p(0))
    t = buf.view(64, 24).cpu().numpy()
    t0 = t[0, 0]
    names = ["mma1:top", "waits", "issued", "Z:full", "mma2:top", "waits", "issued", "Z:u done", "epi:top", "s_full",
             "S in regs", "math done", "h_empty", "H written", "Z:u2 out", "Z:stats", "m2:w2_full", "m2:h_full", "m1:w1_full"]
    print("chunk " + " ".join(f"{n:>11s}" for n in names))
    for gidx in range(40):
        print(f"{gidx:5d} " + " ".join(f"{int(t[gidx, k] - t0) if t[gidx, k] else 0:11d}" for k in range(19)))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "--timeline":
    timeline(train=len(sys.argv) > 2 and sys.argv[2] == "train")
