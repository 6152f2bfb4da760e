#!/bin/bash
# packed count table: parity + C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_minimizer.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/s37_pytest.txt 2>&1; tail -6 gpurun_out/s37_pytest.txt | cut -c1-300
for o in 1 0; do
timeout 300 python bench.py --workload c4 --opt build_packed=$o > gpurun_out/s37_c4_$o.json 2> gpurun_out/s37_c4_$o.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s37_c4_$o.json").read().strip().splitlines()[-1])
print("packed=$o: %.1f Gbp/s"%d["value"], {k:round(v["ms_per_launch"],3) for k,v in d["kernels"].items()}, d.get("parity"))
PY
done
