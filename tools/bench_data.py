"""HBM roofline of the row N3 / N4 kernels: the batch gather of the device-resident data path
(float and bit-mask labels) and the full-list rank metrics.  GPU box only; not yet run (round 1 spent its GPU budget on
parity) -- first item of the next GPU visit."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "ranked-list-truncation_b200"))
from rlt_b200.data import DeviceLoader, synthetic_lists  # noqa: E402
from utils.metrics import _rank_metrics  # noqa: E402


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
    L = 300
    for F, n in ((3, 1_000_000), (1, 2_000_000)):        # splits of 4.8 / 4.8 GB: far larger than L2
        X, y = synthetic_lists(n, L, F, seed=F, device="cuda")
        for pack in (False, True):
            loader = DeviceLoader(X, y, batch_size=n // 2, shuffle=True, pack_labels=pack)
            idx = torch.randperm(n, device="cuda")[: n // 2].contiguous()
            t = timeit(lambda: loader.gather(idx))
            m = n // 2
            b = m * (8 * L * F + 4 * L + ((L + 31) // 32 * 4 if pack else 4 * L) + 8)
            print(f"gather F={F} labels={'bits ' if pack else 'float'} {m} lists: {t*1e3:7.3f} ms  {m/t/1e6:7.1f} M lists/s  "
                  f"{b/t/1e9:6.0f} GB/s  {b/t/1e9/peak:.2f} of copy peak")
            del loader
        del X, y
    n = 200_000
    _, y = synthetic_lists(n, L, 1, seed=9, device="cuda")
    s = torch.rand(n, L, device="cuda")
    t = timeit(lambda: _rank_metrics(y, s, True, True), it=3)
    print(f"rank_metrics (DCG + AUC) {n} lists x {L}: {t*1e3:7.3f} ms incl. the three D2H result copies  {n/t/1e6:6.2f} M lists/s "
          f"({n*L*L/t/1e12:.2f} T document pairs/s)")


if __name__ == "__main__":
    main()
