#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_cli_gpu.py -m gpu -x -q > gpurun_out/s43_pytest.txt 2>&1; tail -25 gpurun_out/s43_pytest.txt | cut -c1-250
timeout 600 python bench.py --only-search > gpurun_out/s43_c3.json 2> gpurun_out/s43_c3.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s43_c3.json").read().strip().splitlines()[-1])
print("C3 ms/pass %.2f  G lookups/s %.2f"%(d["ms_per_pass"], d["lookups_per_s"]/1e9), {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["kernels"].items()}, "frac %.3f"%d["roofline"]["frac"])
print("perfect", d["perfect_search"]["ms_per_pass_wall"], {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["perfect_search"]["kernels"].items()}, d["perfect_search"]["self_query_violations"])
PY
