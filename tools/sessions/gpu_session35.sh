#!/bin/bash
# C4 with adaptive count-table sizing + build parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_minimizer.py -m gpu -x -q -k "build" > gpurun_out/s35_pytest.txt 2>&1; tail -3 gpurun_out/s35_pytest.txt
timeout 300 python bench.py --workload c4 > gpurun_out/s35_c4.json 2> gpurun_out/s35_c4.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s35_c4.json").read().strip().splitlines()[-1])
print("adaptive: %.1f Gbp/s"%d["value"], {k:round(v["ms_per_launch"],3) for k,v in d["kernels"].items()}, d.get("parity"))
PY
