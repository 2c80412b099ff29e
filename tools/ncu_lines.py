#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an ncu report:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fname, hdr, agg = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ie = hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > si and r[2] == "-":      # a source line (its SASS lines carry an address)
        try:
            agg.append((int(r[si] or 0), int(r[ie] or 0), fname, r[0], r[1].strip()))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
print("total samples", tot)
for s, ie_, f, ln, src in sorted(agg, key=lambda a: -a[0])[:top]:
    print(f"{100.0 * s / tot:5.1f}% {s:7d} inst={ie_:9d} {f}:{ln}  {src[:100]}")
