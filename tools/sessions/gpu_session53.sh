#!/bin/bash
# column-sharded read_id, templated narrow kernel + two-process test; C2 kernel timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/s53_sharded.txt 2>&1; tail -15 gpurun_out/s53_sharded.txt | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_minimizer.py tests/test_cli_gpu.py -m gpu -x -q -k "read_id or readid or cli" > gpurun_out/s53_parity.txt 2>&1; tail -3 gpurun_out/s53_parity.txt | cut -c1-300
timeout 600 python bench.py --no-search --no-cpu-baseline --steps 5 > gpurun_out/s53_bench.json 2> gpurun_out/s53_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s53_bench.json").read().strip().splitlines()[-1])
print("value %.1fM e2e %.1fM ms %.2f"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), {k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()})
PY
