#!/bin/bash
# optimistic count table for read-set queries: parity + C1 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q  > gpurun_out/s50_pytest.txt 2>&1; tail -3 gpurun_out/s50_pytest.txt | cut -c1-250
CID_TRACE=1 timeout 600 python bench.py --workload c1 > gpurun_out/s50_c1.json 2> gpurun_out/s50_c1.err; grep "cid trace" gpurun_out/s50_c1.err | tail -2 | head -1
python - <<PY
import json
d=json.loads(open("gpurun_out/s50_c1.json").read().strip().splitlines()[-1])
print("Gbp/s %.2f ms %.1f kern_ms %.1f lookups/s %.3g build %.2f"%(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["lookups_per_s"], d["build"]["gbp_per_s"]))
print({k:round(v["ms_per_launch"]*v["launches_per_step"],2) for k,v in d["kernels"].items()}, d["cpu_baseline"]["matches_gpu_report"])
PY
