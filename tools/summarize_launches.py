#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.1f %% |" % (k[:80], c, t, t / c, t / tot * 100))


if __name__ == "__main__":
    main(sys.argv[1])
