#!/bin/bash
# 2 GPUs: C5 (2,500 accessions) with the column-sharded default report and read_id timed
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29542 bench.py --gpus 2 --workload c5 --c5-acc 2500 --c5-extra --steps 3 --warmup 1 > gpurun_out/s70_c5_n2.json 2> gpurun_out/s70_c5_n2.err; echo "full rc=$?"; grep "rank0\]:" gpurun_out/s70_c5_n2.err | grep -v Warning | tail -3 | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/s70_c5_n2.json").read().strip().splitlines()[-1])
print("%.2f G lookups/s"%(d["value"]/1e9), json.dumps(d["sharded_default_report_and_read_id"], indent=1)[:2500])
PY
