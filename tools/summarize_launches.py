"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of time)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3,
                 "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, val * scale))
    tot = sum(ms for _, ms in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ms in rows:
        agg[name][0] += 1
        agg[name][1] += ms
    print(f"launches {len(rows)}  total {tot:.3f} ms (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'count':>8s} {'ms':>12s} {'share':>7s} {'avg_us':>10s}")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {cnt:8d} {ms:12.3f} {100 * ms / tot:6.2f}% {1e3 * ms / cnt:10.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
