#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
Usage: python scripts/summarize_launches.py gpurun_out/<tag>/launches.csv [skip_first_n] > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        name = re.sub(r">\(.*$", ">", name) if "<" in name else re.sub(r"\(.*$", "", name)
        rows.append((int(r["ID"]), name, ns))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for _, name, ns in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)} (first {skip} skipped), total device time {total/1e6:.3f} ms "
          f"(ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:110]}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% | {ns/n/1e3:.1f} |")


if __name__ == "__main__":
    main()
