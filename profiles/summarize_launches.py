#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("h2::", "")
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[row["Metric Unit"][:2]]
        agg.setdefault(k, []).append(v * scale)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':42s} {'n':>4s} {'mean ms':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:42]:42s} {len(v):4d} {sum(v) / len(v):10.4f} {100 * sum(v) / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
