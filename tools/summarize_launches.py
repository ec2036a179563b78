#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches.csv [first_id last_id] > profiles/rNN_launch_shares.md

ncu serialises launches and runs them cold-cache, so only the SHARES are meaningful (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("sylph::", "").replace("void ", "")
    return name.strip()


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if lo <= i <= hi:
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) * 1e-3, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, us, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list summary: {path} (launch ids {lo}..{min(hi, lo + len(rows) - 1)}, {len(rows)} launches, "
          f"{total / 1000:.3f} ms serialised)\n")
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {us / total:.3f} |")


if __name__ == "__main__":
    main()
