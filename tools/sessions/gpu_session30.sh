#!/bin/bash
# set-only build path: parity (all suites that build) + build rates
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s30_pytest.txt 2>&1; tail -4 gpurun_out/s30_pytest.txt
timeout 600 python bench.py --workload c5 --c5-acc 1250 --steps 3 --warmup 1 > gpurun_out/s30_c5.json 2> gpurun_out/s30_c5.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s30_c5.json").read().strip().splitlines()[-1])
print("C5 shard build", d["build"], d["parity"])
PY
timeout 600 python bench.py --workload c5 --c5-acc 1250 --steps 3 --warmup 1 --opt build_set=0 > gpurun_out/s30_c5_old.json 2> gpurun_out/s30_c5_old.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s30_c5_old.json").read().strip().splitlines()[-1])
print("C5 shard build (count table)", d["build"])
PY
