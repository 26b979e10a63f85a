#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total time, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-44s %5s %12s %7s" % ("kernel", "n", "total ms", "share"))
    for k, (n, t) in agg.items():
        print("%-44s %5d %12.3f %6.1f%%" % (k, n, t, 100 * t / tot))
    print("%-44s %5s %12.3f" % ("all", "", tot))


if __name__ == "__main__":
    main(sys.argv[1])
