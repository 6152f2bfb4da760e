#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s3_pytest.txt
tail -5 gpurun_out/s3_pytest.txt
CID_TRACE=1 python bench.py --no-search --steps 10 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
grep "cid trace" gpurun_out/s3_bench.err | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'], d['cpu_baseline'] and d['cpu_baseline'].get('matches_gpu_classification'))
for k,v in d['roofline']['kernels'].items(): print(k,v['ms_per_launch'])
PY
