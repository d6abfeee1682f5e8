#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path, skip_names=("at::",)):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-58s %6s %11s %11s %7s" % ("kernel", "n", "total ms", "avg us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %6d %11.3f %11.3f %6.1f%%" % (k[:58], n, t, 1e3 * t / n, 100 * t / tot))
    print("%-58s %6s %11.3f" % ("TOTAL", "", tot))


if __name__ == "__main__":
    main(sys.argv[1])
