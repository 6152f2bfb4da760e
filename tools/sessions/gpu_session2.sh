#!/bin/bash
# GPU session 2: pipelined read_id parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s2_pytest.txt
tail -15 gpurun_out/s2_pytest.txt
( time timeout 900 python bench.py --no-search ) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
tail -5 gpurun_out/s2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'])
PY
