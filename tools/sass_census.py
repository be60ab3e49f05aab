#!/usr/bin/env python
"""SASS census of rag4dyg_b200/libr4d.so: per kernel, the counts of the mnemonics that prove what the kernel is made
of (tcgen05 MMAs = UTCHMMA/UTCQMMA..., TMEM loads = LDTM, TMA = UTMALDG / UBLKCP, tcgen05 barriers = UTCBAR, POPC,
LOP3, ...).  Usage: python tools/sass_census.py [out.txt]   (cuobjdump -sass; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rag4dyg_b200", "libr4d.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP",
         "SYNCS", "POPC", "LOP3", "ATOMS", "ATOMG", "RED", "REDUX", "SHFL", "VOTE", "MATCH", "LDS", "STS", "LDG", "STG",
         "HMMA", "IMMA", "DFMA", "MUFU", "BAR", "ELECT", "WARPSYNC"]


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    out.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel; columns = watched mnemonics\n")
    out.write("# whole library: " + ", ".join(f"{k}={tot[k]}" for k in WATCH if tot[k]) + f", total={tot['_total']}\n\n")
    for name, c in per.items():
        seen = ", ".join(f"{k}={c[k]}" for k in WATCH if c[k])
        out.write(f"{name}\n    total={c['_total']}  {seen}\n")


if __name__ == "__main__":
    main()
