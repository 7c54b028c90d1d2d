"""HBM roofline of the streaming kernels K3 (cut head + reward loss + gradient) and K4 (argmax cut + F1/DCG) on
synthetic probabilities / labels (BASELINE.json configs[4]: batch inference sweep, L = 300 -> 1000).  GPU box only."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import ops  # noqa: E402


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e-3


def main():
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    for L, n in ((300, 2_000_000), (500, 1_200_000), (1000, 600_000)):     # inputs of 4.8 GB: far larger than L2
        g = torch.Generator(device="cuda").manual_seed(L)
        z = torch.randn(n, L, device="cuda", generator=g)
        y = (torch.rand(n, L, device="cuda", generator=g) < 0.1).float()
        p = torch.softmax(z, dim=1)
        t = timeit(lambda: ops.eval_cut(p, y))
        b = n * (8 * L + 28)
        print(f"K4 eval_cut        L={L:4d} n={n}: {t*1e3:7.3f} ms  {n/t/1e6:8.1f} M lists/s  {b/t/1e9:7.0f} GB/s  {b/t/1e9/peak:.2f} of copy peak")
        yb = ops.pack_labels(y)
        t = timeit(lambda: ops.eval_cut(p, None, label_bits=yb))
        b = n * (4 * L + 4 * yb.shape[1] + 28)
        print(f"K4 eval_cut_bits   L={L:4d} n={n}: {t*1e3:7.3f} ms  {n/t/1e6:8.1f} M lists/s  {b/t/1e9:7.0f} GB/s  {b/t/1e9/peak:.2f} of copy peak (4 L + L/8 bytes per list)")
        del yb
        if L <= 320:
            grad = torch.empty_like(z)
            lpl = torch.empty(n, device="cuda")
            bits = ops.pack_labels(y)
            for kind, metric in (("choopy", "f1"), ("js", "f1"), ("js", "dcg"), ("raml", "f1"), ("raml", "dcg"), ("kl", "f1")):
                t = timeit(lambda: ops.cut_loss(z, y, loss_kind=kind, metric=metric, tau=0.85, grad=grad, loss_per_list=lpl))
                b = n * (12 * L + 4)
                print(f"K3 cut_loss {kind:6s}/{metric:3s} L={L:4d} n={n}: {t*1e3:7.3f} ms  {n/t/1e6:8.1f} M lists/s  {b/t/1e9:7.0f} GB/s  "
                      f"{b/t/1e9/peak:.2f} of copy peak")
                t = timeit(lambda: ops.cut_loss(z, None, label_bits=bits, loss_kind=kind, metric=metric, tau=0.85, grad=grad,
                                                loss_per_list=lpl))
                b = n * (8 * L + 4 * bits.shape[1] + 4)
                print(f"K3 cut_loss_bits {kind:6s}/{metric:3s} L={L:4d}: {t*1e3:7.3f} ms  {n/t/1e6:8.1f} M lists/s  {b/t/1e9:7.0f} GB/s  "
                      f"{b/t/1e9/peak:.2f} of copy peak (8 L + L/8 bytes per list)")
        del z, y, p


if __name__ == "__main__":
    main()
