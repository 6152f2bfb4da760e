#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s9_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s9_pytest.txt
tail -25 gpurun_out/s9_pytest.txt
CID_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
grep "cid trace" gpurun_out/s9_bench.err | tail -9
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s9_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
for k,v in d['roofline']['kernels'].items(): print(k,v['ms_per_launch'])
s=d['search_c3']; print({k:s[k] for k in ('lookups_per_s','ms_per_pass','build_gbp_per_s')}, s.get('roofline'), s['kernels'])
PY
python - <<PY
import json
d=json.loads(open("gpurun_out/s9_bench.json").read().strip().splitlines()[-1])
print(d["roofline"]["random_access"])
PY
