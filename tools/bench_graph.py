"""Small-batch regime (the reference's run.py batch: one group of 63 lists): eager train step (~250 launches) against
the CUDA-graph replay of Engine.capture_train_step, both followed by the fused Adam step.  GPU box only.
usage: python tools/bench_graph.py [model] [group_size]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import models  # noqa: E402
from rlt_b200.data import synthetic_lists  # noqa: E402
from rlt_b200.engine import Engine  # noqa: E402
from rlt_b200.optim import FusedAdam  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "choopy"
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 63
    torch.manual_seed(1234)
    model = bench.build_model(models, name).cuda()
    eng = Engine(model, n_groups=1, group_size=S, seq_len=300, training=True)
    opt = FusedAdam.for_engine(eng, lr=3e-5, weight_decay=1e-3)
    x, y = synthetic_lists(S, 300, bench.N_FEATURES[name], seed=1, device="cuda")

    def timed(fn, n=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    def eager():
        eng.train_step(x, y)
        opt.step()

    ms_e = timed(eager)
    replay = eng.capture_train_step(x, y)

    def graphed():
        replay()
        opt.step()

    ms_g = timed(graphed)
    print(f"{name}: one group of {S} lists x 300 per step: eager {ms_e:.3f} ms/step ({S / ms_e * 1e3:.0f} lists/s), "
          f"CUDA graph {ms_g:.3f} ms/step ({S / ms_g * 1e3:.0f} lists/s), x{ms_e / ms_g:.2f}")


if __name__ == "__main__":
    main()
