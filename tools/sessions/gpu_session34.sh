#!/bin/bash
# C4: first count-table size (positions / div) sweep
mkdir -p gpurun_out
for d in 2 4 8 16; do
timeout 300 python bench.py --workload c4 --c4-acc 12 --steps 2 --warmup 1 --opt build_table_div=$d > gpurun_out/s34_d$d.json 2> gpurun_out/s34_d$d.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s34_d$d.json").read().strip().splitlines()[-1])
print("div $d: %.1f Gbp/s"%d["value"], {k:round(v["ms_per_launch"],3) for k,v in d.get("kernels",d.get("roofline",{}).get("kernels",{})).items()} if isinstance(d.get("kernels",None),dict) else list(d.keys())[:30])
PY
done
