#!/usr/bin/env python
"""Instruction / stall-sample share per source-line range of an .ncu-rep captured with --import-source on.
Usage: ncu_regions.py report.ncu-rep file.cu start:end:name [start:end:name ...]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, fname = sys.argv[1], sys.argv[2]
    regions = []
    for a in sys.argv[3:]:
        s, e, n = a.split(":")
        regions.append((int(s), int(e), n))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    agg = collections.OrderedDict()
    cur, curfile = None, None
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            curfile = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] != "":
            cur = (curfile, int(r[0]))
            agg.setdefault(cur, [0, 0])
        elif cur is not None and len(r) > 7:
            try:
                agg[cur][0] += int(r[4])
                agg[cur][1] += int(r[7])
            except ValueError:
                pass
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    out = collections.defaultdict(lambda: [0, 0])
    for (f, l), v in agg.items():
        name = "inlined:" + f
        if f == fname:
            name = "other"
            for s, e, n in regions:
                if s <= l <= e:
                    name = n
                    break
        out[name][0] += v[0]
        out[name][1] += v[1]
    for k, v in sorted(out.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} samples {v[0] * 100 / ts:5.1f}%  instructions {v[1] * 100 / ti:5.1f}%")


if __name__ == "__main__":
    main()
