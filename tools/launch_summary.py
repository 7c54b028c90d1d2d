"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (share of total device time)."""
import collections
import csv
import re
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v for _, v in agg.values())
    print(f"{n} launches, {tot / 1e6:.3f} ms total device time (cold-cache, serialised: compare shares)")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v / 1e6:10.3f} ms {100 * v / tot:5.1f}%  n={c:4d}  {k[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
