#!/bin/bash
# C5 shard shape on one GPU (1,250 accessions, 160-byte rows): query_gather with / without the 64-byte L2 prefetch size
mkdir -p gpurun_out
for v in 0 1; do
  timeout 600 python bench.py --workload c5 --c5-acc 1250 --steps 3 --warmup 1 --opt gather_l2_64b=$v > gpurun_out/s46_c5_pf$v.json 2> gpurun_out/s46_c5_pf$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s46_c5_pf$v.json").read().strip().splitlines()[-1])
print("pf64=$v", "G lookups/s %.3f"%(d["value"]/1e9), "gather ms %.3f frac %.3f"%(d["roofline"]["ms_per_pass"], d["roofline"]["frac"]))
PY
done
for v in 0 1; do
  timeout 600 python bench.py --only-search --opt gather_l2_64b=$v > gpurun_out/s46_c3_pf$v.json 2> gpurun_out/s46_c3_pf$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s46_c3_pf$v.json").read().strip().splitlines()[-1])
print("C3 pf64=$v ms/pass %.2f"%d["ms_per_pass"], {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["kernels"].items()}, "frac %.3f"%d["roofline"]["frac"], "perfect", {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["perfect_search"]["kernels"].items()})
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q -k "search or gene or stride or perfect or shard" 2>&1 | tail -2
