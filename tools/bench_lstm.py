"""GPU: BiLSTM forward / backward time by batch size and tile shape (rlt_set_option("lstm_tile")), with a parity check of
each shape against the plain fp32 validation kernels."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200 import _lib  # noqa: E402
from rlt_b200.autograd import BiLstm  # noqa: E402


def run(B, L, F, tile, reps=3):
    _lib.set_option("lstm_tile", tile)
    torch.manual_seed(1)
    lstm = torch.nn.LSTM(input_size=F, hidden_size=128, num_layers=2, batch_first=True, bidirectional=True).cuda()
    x = torch.randn(B, L, F, device="cuda")
    dy = torch.randn(B, L, 256, device="cuda") * 1e-3
    ps = [p.detach().clone().requires_grad_(True) for p in lstm._flat_weights]
    for _ in range(2):
        y = BiLstm.apply(x, 128, 2, *ps)
        (y * dy).sum().backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        e[0].record()
        y = BiLstm.apply(x, 128, 2, *ps)
        e[1].record()
        (y * dy).sum().backward()
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]) / reps
        tb += e[1].elapsed_time(e[2]) / reps
    return tf, tb, y.detach(), [p.grad.clone() for p in ps]


def main():
    L, F = 300, 3
    for B in (63, 64, 256, 2048, 4096):
        ref = None
        if B <= 256:
            _lib.set_option("lstm_backend", 1)
            _, _, yr, gr = run(B, L, F, 0, reps=1)
            _lib.set_option("lstm_backend", 0)
            ref = (yr, gr)
        for tile in (64, 32, 0):
            tf, tb, y, g = run(B, L, F, tile)
            msg = f"B={B:5d} tile={tile:2d}: fwd {tf:7.3f} ms  bwd {tb:7.3f} ms  ({B / (tf + tb) * 1e3:9.0f} lists/s fwd+bwd)"
            if ref is not None:
                ey = (y - ref[0]).abs().max().item()
                eg = max(((a - b).abs().max() / (b.abs().max() + 1e-12)).item() for a, b in zip(g, ref[1]))
                msg += f"   vs fp32 kernels: max|dy| {ey:.2e}  max rel grad err {eg:.2e}"
            print(msg)
    _lib.set_option("lstm_tile", 0)


if __name__ == "__main__":
    main()
