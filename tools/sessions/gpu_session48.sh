#!/bin/bash
# count-table compaction for large queries: parity + C1 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_cli_gpu.py -m gpu -x -q > gpurun_out/s48_pytest.txt 2>&1; tail -4 gpurun_out/s48_pytest.txt | cut -c1-250
for c in 0 1; do
CID_TRACE=1 timeout 600 python bench.py --workload c1 --opt query_compact=$c > gpurun_out/s48_c1_$c.json 2> gpurun_out/s48_c1_$c.err; grep "cid trace" gpurun_out/s48_c1_$c.err | tail -2 | head -1
python - <<PY
import json
d=json.loads(open("gpurun_out/s48_c1_$c.json").read().strip().splitlines()[-1])
print("compact=$c Gbp/s %.2f ms %.1f kern_ms %.1f lookups/s %.3g build %.2f"%(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["lookups_per_s"], d["build"]["gbp_per_s"]))
print({k:round(v["ms_per_launch"]*v["launches_per_step"],2) for k,v in d["kernels"].items()}, d["cpu_baseline"]["matches_gpu_report"])
PY
done
