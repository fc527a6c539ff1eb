"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
    rows = rows[skip:]
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    print(f"launches: {len(rows)}  total device time: {tot / 1e6:.3f} ms  (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{name[:70]:70s} {n:8d} {ns / 1e6:10.3f} {ns / n / 1e3:10.2f} {100 * ns / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
