"""Per-call-site device time of one train step (in-situ CUDA-event brackets, see common.h KernelTag).
usage: python tools/bench_tags.py [model] [groups] — GPU box only."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
sys.path.insert(0, str(ROOT))
from rlt_b200 import _lib, ops  # noqa: E402
from rlt_b200.data import synthetic_lists  # noqa: E402
from rlt_b200.engine import Engine  # noqa: E402
import models  # noqa: E402
import bench  # noqa: E402

TAGS = {1: "QKV", 2: "OUT_PROJ", 3: "FFN1", 4: "FFN2", 5: "D_FFN2", 6: "D_FFN1", 7: "DW_FFN2", 8: "DW_FFN1",
        9: "ATTN_FWD", 10: "ATTN_BWD", 11: "FFN_FUSED", 12: "LSTM"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "choopy"
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    lib = ops.lib()
    B = G * 64
    x, y = synthetic_lists(B, 300, bench.N_FEATURES[name], seed=1, device="cuda")
    torch.manual_seed(1234)
    model = bench.build_model(models, name).cuda()
    eng = Engine(model, n_groups=G, group_size=64, seq_len=300, training=True)
    for _ in range(2):
        eng.train_step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.train_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {B} lists/step, {step_ms:.2f} ms/step, {B / step_ms * 1e3:.0f} lists/s")
    acc = 0.0
    for tag, label in TAGS.items():
        _lib.set_option("time_tag", tag)
        lib.rlt_timing_reset()
        eng.train_step(x, y)
        torch.cuda.synchronize()
        tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
        _lib.check(lib.rlt_timing_read(ctypes.byref(tot), ctypes.byref(cnt)), "timing")
        if cnt.value:
            acc += tot.value
            print(f"  {label:10s} n={cnt.value:3d}  total {tot.value:8.3f} ms  avg {tot.value / cnt.value:7.3f} ms  "
                  f"{100 * tot.value / step_ms:5.1f}% of step")
    _lib.set_option("time_tag", 0)
    lib.rlt_timing_reset()
    print(f"  tagged total {acc:.2f} ms of {step_ms:.2f}")


if __name__ == "__main__":
    main()
