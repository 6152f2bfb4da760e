#!/usr/bin/env python
"""Hot SASS instructions of one kernel from `ncu -i rep --page source --csv --kernel-name K` output.
usage: ncu_hot.py file.csv [min_exec_per_warp_block]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot_s = sum(int(r[ix["# Samples"]]) for r in body)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in body)
mx = max(int(r[ix["Instructions Executed"]]) for r in body)
print("total samples", tot_s, "warp instr", tot_i, "sass lines", len(body))
for k, r in enumerate(body):
    ie, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    if ie >= mx * thr / 100.0 or s > tot_s * 0.01:
        print(f"{k:4d} exec={ie:12d} samp={100*s/tot_s:5.1f}% thr={r[ix['Avg. Threads Executed']]:>5s} "
              f"conf={r[ix['L1 Conflicts Shared N-Way']]:>4s} {r[ix['Source']].strip()[:100]}")
