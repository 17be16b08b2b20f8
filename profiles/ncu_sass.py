#!/usr/bin/env python
"""Per-SASS-instruction view of an `ncu --page source --print-source cuda,sass --csv` dump: offset, samples,
warp-level executions, dominant stall reasons, instruction.  usage: ncu_sass.py x.csv [lo_hex hi_hex] [min_samples]"""
import csv, sys
path = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hdr = None
ins = {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        names = r
        continue
    if hdr is None or len(r) < 8 or not r[2].startswith("0x"):
        continue
    a = int(r[2], 16)
    try:
        smp, ex = int(r[6]), int(r[7])
    except ValueError:
        continue
    st0 = names.index("stall_barrier")
    stalls = {names[i][6:]: int(r[i]) for i in range(st0, st0 + 17) if r[i].isdigit() and int(r[i]) > 0}
    ins[a] = (smp, ex, stalls, r[3].strip())
base = min(ins)
tot = sum(v[0] for v in ins.values()) or 1
print(f"instructions {len(ins)}, samples {tot}")
for a in sorted(ins):
    o = a - base
    if lo <= o < hi:
        smp, ex, stalls, txt = ins[a]
        if smp < mins:
            continue
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print(f"{o:06x} {smp:6d} {smp*100/tot:5.2f}% {ex:9d}  {txt[:70]:70s} {' '.join(f'{k}:{v}' for k, v in top)}")
if len(sys.argv) > 5 and sys.argv[5] == "agg":
    import collections
    agg = collections.Counter(); agi = collections.Counter()
    for a in ins:
        o = a - base
        if lo <= o < hi:
            smp, ex, stalls, txt = ins[a]
            op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
            op = op.split(".")[0]
            agg[op] += smp; agi[op] += ex
    t2 = sum(agg.values()); i2 = sum(agi.values())
    print(f"range {lo:x}-{hi:x}: samples {t2} ({t2*100/tot:.1f}% of kernel), warp instr {i2}")
    for op, s in agg.most_common(25):
        print(f"  {op:12s} samples {s:6d} {s*100/t2:5.1f}%   instr {agi[op]:10d} {agi[op]*100/i2:5.1f}%")
