#!/bin/bash
# query_gather with the bit-sliced cross-partial flush: parity + C3/C5-shard measurement
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or gene" > gpurun_out/s25_pytest.txt 2>&1; tail -5 gpurun_out/s25_pytest.txt
timeout 600 python bench.py --only-search > gpurun_out/s25_c3.json 2> gpurun_out/s25_c3.err; tail -2 gpurun_out/s25_c3.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s25_c3.json").read().strip().splitlines()[-1])
print("C3 ms/pass %.2f  G lookups/s %.2f"%(d["ms_per_pass"], d["lookups_per_s"]/1e9), {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["kernels"].items()}, "frac %.3f"%d["roofline"]["frac"])
print("perfect", d["perfect_search"]["ms_per_pass_wall"], d["perfect_search"]["kernels"], d["perfect_search"]["self_query_violations"])
PY
timeout 600 python bench.py --workload c5 --c5-acc 1250 --steps 3 --warmup 1 > gpurun_out/s25_c5.json 2> gpurun_out/s25_c5.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s25_c5.json").read().strip().splitlines()[-1])
print("C5 shard", d["value"]/1e9, d["ms_per_step"], d["kernels"], d["roofline"]["frac"], d["parity"])
PY
