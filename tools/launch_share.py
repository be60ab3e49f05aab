#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: launch_share.py in.csv [out.txt]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui].replace("second", "s") if r[ui].endswith("second") else r[ui], v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    out.write(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms of device time (serialised, cold cache)\n")
    for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.write(f"{k[:90]:90s} n={n:4d} total_us={us:12.1f} avg_us={us / n:10.1f} share={100 * us / tot:6.2f}%\n")


if __name__ == "__main__":
    main()
