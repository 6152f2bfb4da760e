#!/usr/bin/env python
"""Per-source-line instruction/sample totals from `ncu -i rep --page source --print-source cuda,sass --csv -k K`.
usage: ncu_lines.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 8 and r[0].isdigit():
        key = (cur, int(r[0]))
        f = lambda x: int(x) if x.isdigit() else 0
        ie = f(r[hdr["Instructions Executed"]]); s = f(r[hdr["# Samples"]])
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += ie; a[1] += s
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp instr", ti, "samples", ts)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:>16s}:{k[1]:<5d} instr {100*a[0]/ti:5.1f}%  samp {100*a[1]/max(ts,1):5.1f}%  {a[2]}")
