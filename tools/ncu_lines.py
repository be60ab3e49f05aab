#!/usr/bin/env python
"""Per-CUDA-line stall samples / executed instructions from an .ncu-rep captured with --import-source on.
Usage: ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    agg = collections.OrderedDict()
    cur, curfile = None, None
    src = {}
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            curfile = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] != "":
            cur = (curfile, r[0])
            agg.setdefault(cur, [0, 0])
            src[cur] = ",".join(r[1:4])[:90]
        elif cur is not None and len(r) > 7:
            try:
                agg[cur][0] += int(r[4])
                agg[cur][1] += int(r[7])
            except ValueError:
                pass
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"samples {ts}  warp-instructions {ti}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] * 100 / ts:5.1f}% samp {v[1] * 100 / ti:5.1f}% inst  {k[0]}:{k[1]}  {src[k]}")


if __name__ == "__main__":
    main()
