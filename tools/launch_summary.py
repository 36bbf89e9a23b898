#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for one step and the
per-launch sequence.  Usage: python tools/launch_summary.py gpurun_out/launches.csv <steps> [--seq]"""
import collections
import csv
import re
import sys


def main(path, steps, seq=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    items = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"], r["Block Size"]) for r in rows]
    n = len(items) // steps
    frame = items[n:2 * n] if steps > 1 else items
    tot = sum(d for _, d, _, _ in frame)
    print(f"# {path}: {len(items)} launches profiled, {n} per step; step total {tot:.1f} us (cold-cache, serialised: compare SHARES)")
    agg = collections.OrderedDict()
    for name, d, g, b in frame:
        k = re.sub(r"\(.*", "", name)[:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += d
    for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{d:9.1f} us {100 * d / tot:5.1f}%  x{c:3d}  {k}")
    if seq:
        print()
        for i, (name, d, g, b) in enumerate(frame):
            print(f"{i:4d} {d:8.1f} us  grid {g:18s} block {b:14s} {re.sub(r'[(].*', '', name)[:60]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), "--seq" in sys.argv)
