#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s15_pytest.txt 2>&1; tail -4 gpurun_out/s15_pytest.txt
timeout 900 python bench.py --workload c4 --c4-acc 12 --steps 8 --warmup 2 > gpurun_out/s15_c4.json 2> gpurun_out/s15_c4.err
tail -3 gpurun_out/s15_c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s15_c4.json').read().strip().splitlines()[-1])
print('c4 value',d['value'],'ms/step',d['ms_per_step'],{k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()}, d['parity'], d['auto_cutoff_used'][:3])
PY
timeout 600 python bench.py --workload c5 --c5-acc 1250 --steps 3 --warmup 1 > gpurun_out/s15_c5.json 2> gpurun_out/s15_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s15_c5.json').read().strip().splitlines()[-1])
print('c5 build',d['build'])
PY
