#!/bin/bash
# 2 GPUs: sharded tests after the slots_popcount rewrite + C5 extra timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 240 > gpurun_out/s65_sharded.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s65_sharded.txt | head -20 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29542 bench.py --gpus 2 --workload c5 --c5-acc 2500 --c5-extra --steps 3 --warmup 1 > gpurun_out/s65_c5_n2.json 2> gpurun_out/s65_c5_n2.err; echo "full rc=$?"; grep "rank0\]:" gpurun_out/s65_c5_n2.err | grep -v Warning | tail -3 | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/s65_c5_n2.json").read().strip().splitlines()[-1])
x=d["sharded_default_report_and_read_id"]
print("%.2f G lookups/s"%(d["value"]/1e9), "report ms", x["default_report"]["ms_per_query_wall"], "read_id ms", x["read_id"]["ms_read_id_batch_wall"], x["read_id"]["ms_gather_merge_classify_wall"])
PY
