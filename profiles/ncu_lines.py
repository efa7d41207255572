#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > both.csv; ncu_lines.py both.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iinst, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iinst or r[0] == "":
        continue
    try:
        key = (fname, int(r[0]))
    except ValueError:
        continue
    try:
        ni, ns = int(r[iinst] or 0), int(r[isamp] or 0)
    except ValueError:
        continue  # a source line with embedded quotes confused the csv reader
    a = agg.setdefault(key, [r[1], 0, 0])
    a[1] += ni
    a[2] += ns
tot = sum(a[1] for a in agg.values())
tots = sum(a[2] for a in agg.values())
print(f"total warp instructions {tot}, stall samples {tots}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:<4d} {100 * a[1] / tot:5.1f}% inst {100 * a[2] / max(tots, 1):5.1f}% samples  {a[0].strip()[:100]}")
