#!/bin/bash
# 2 GPUs: sharded tests, then C5 (2,500 accessions) with the column-sharded default report and read_id timed
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 240 > gpurun_out/s55_sharded.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s55_sharded.txt | head -30 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 2 --workload c5 --quick --c5-acc 200 --c5-extra --steps 2 --warmup 1 > gpurun_out/s55_c5_quick.json 2> gpurun_out/s55_c5_quick.err; echo "quick rc=$?"; grep -E "Error|error|assert" gpurun_out/s55_c5_quick.err | tail -5 | cut -c1-300
timeout 700 $TR --master-port 29542 bench.py --gpus 2 --workload c5 --c5-acc 2500 --c5-extra --steps 3 --warmup 1 > gpurun_out/s55_c5_n2.json 2> gpurun_out/s55_c5_n2.err; echo "full rc=$?"; grep -E "Error|error|assert" gpurun_out/s55_c5_n2.err | tail -5 | cut -c1-300
python - <<PY
import json
for f in ("gpurun_out/s55_c5_quick.json", "gpurun_out/s55_c5_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.2f G lookups/s"%(d["value"]/1e9), json.dumps(d["sharded_default_report_and_read_id"])[:1500])
    except Exception as e: print(f, "ERR", e)
PY
