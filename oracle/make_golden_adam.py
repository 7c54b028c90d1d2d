"""Golden vectors for the optimizer row (N1): torch.optim.Adam exactly as the reference constructs it
(run.py:104: `optim.Adam(self.model.parameters(), lr=args.lr, weight_decay=self.weight_decay)`), stepped on CPU in the
build container.  Writes tests/golden/adam.npz.   usage: python oracle/make_golden_adam.py"""
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
SHAPES = [(40, 16), (48, 16), (48,), (256, 16), (256,), (1,), (7, 3), (4097,), (65, 3)]


def main():
    out = {}
    for case, (lr, wd, steps) in {"wd": (3e-5, 1e-3, 4), "nowd": (1e-3, 0.0, 3)}.items():
        g = torch.Generator().manual_seed(20240229)
        params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.1) for s in SHAPES]
        opt = torch.optim.Adam(params, lr=lr, weight_decay=wd)
        out[f"{case}_hyper"] = np.array([lr, wd, steps], dtype=np.float64)
        for i, p in enumerate(params):
            out[f"{case}_p0_{i}"] = p.detach().numpy().copy()
        for t in range(steps):
            for i, p in enumerate(params):
                scale = 10.0 ** (-(i % 4) - 1)        # gradients of very different magnitudes
                p.grad = torch.randn(p.shape, generator=g) * scale
                out[f"{case}_g{t}_{i}"] = p.grad.numpy().copy()
            opt.step()
            for i, p in enumerate(params):
                out[f"{case}_p{t + 1}_{i}"] = p.detach().numpy().copy()
        for i, p in enumerate(params):
            st = opt.state[p]
            out[f"{case}_m_{i}"] = st["exp_avg"].numpy().copy()
            out[f"{case}_v_{i}"] = st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(ROOT / "tests" / "golden" / "adam.npz", **out)
    print("wrote tests/golden/adam.npz with", len(out), "arrays; torch", torch.__version__)


if __name__ == "__main__":
    main()
