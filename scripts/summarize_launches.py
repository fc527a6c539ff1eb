"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)
into per-kernel totals and shares; with --json also writes the per-kernel averages (time, DRAM bytes per launch)."""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT = {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1,
        "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, json_out=None):
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    per_id = defaultdict(dict)
    names = {}
    for r in csv.DictReader(lines):
        val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1)
        per_id[r["ID"]][r["Metric Name"]] = val
        names[r["ID"]] = re.sub(r"\(.*", "", r["Kernel Name"])
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per_id.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"launches: {len(per_id)}  total device time: {tot / 1e6:.3f} ms  (cold-cache, serialised: compare shares)")
    print(f"{'kernel':64s} {'launches':>8s} {'total ms':>9s} {'avg us':>8s} {'share':>6s} {'rd MB/l':>8s} {'wr MB/l':>8s}")
    for name, (n, ns, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{name[:64]:64s} {n:8d} {ns / 1e6:9.3f} {ns / n / 1e3:8.2f} {100 * ns / tot:5.1f}% {rd / n / 1e6:8.2f} {wr / n / 1e6:8.2f}")
    if json_out:
        out = {name: {"launches": n, "avg_us": ns / n / 1e3, "share": ns / tot, "dram_read_bytes_per_launch": rd / n,
                      "dram_write_bytes_per_launch": wr / n} for name, (n, ns, rd, wr) in agg.items()}
        with open(json_out, "w") as fh:
            json.dump({"source": path, "total_ms": tot / 1e6, "kernels": out}, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None)
