#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].replace("<unnamed>::", "")[:72]
        agg.setdefault((name, row["Grid Size"]), []).append(float(row["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print("%-74s %-14s %5s %10s %10s %6s" % ("kernel", "grid", "n", "avg_ns", "min_ns", "share"))
    for (name, grid), v in agg.items():
        print("%-74s %-14s %5d %10.0f %10.0f %5.1f%%" % (name, grid, len(v), sum(v) / len(v), min(v), 100 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1])
