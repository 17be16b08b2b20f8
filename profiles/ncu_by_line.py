#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; python ncu_by_line.py x.csv [min_pct]"""
import csv, sys
path = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
fname = ""
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or r[0] == "" or not r[0].isdigit():
        continue
    S, I = hdr["# Samples"], hdr["Instructions Executed"]
    try:
        out.append((fname, int(r[0]), r[1], int(r[S]), int(r[I])))
    except ValueError:
        pass
ts = sum(o[3] for o in out) or 1
ti = sum(o[4] for o in out) or 1
print(f"total samples {ts}, total warp instructions {ti}")
print("file:line  samples%  instr%  source")
for f, ln, src, s, i in out:
    if s * 100 / ts >= minpct or i * 100 / ti >= minpct:
        print(f"{f}:{ln:<4} {s*100/ts:6.2f} {i*100/ti:6.2f}  {src.strip()[:120]}")
