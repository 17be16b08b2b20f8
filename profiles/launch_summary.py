#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, mean of the last launches.
usage: python profiles/launch_summary.py launches.csv [last_n]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
last = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = {n: i for i, n in enumerate(rows[hi])}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    name = re.sub(r"\(.*", "", r[h["Kernel Name"]]).replace("void ", "").replace("ivf::", "").replace("<unnamed>::", "")[:56]
    v = float(r[h["Metric Value"]].replace(",", ""))
    u = r[h["Metric Unit"]]
    v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    agg.setdefault(name, []).append(v)
for k, v in agg.items():
    t = v[-last:]
    print(f"{k:56s} launches {len(v):4d}   mean of last {len(t)}: {sum(t)/len(t):9.1f} us")
