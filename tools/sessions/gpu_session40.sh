#!/bin/bash
# 8 GPUs: C5 at full size, NCCL all_gather vs count exchange fused into the gather kernel (peer stores over NVSwitch)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --workload c5 --steps 5 --warmup 2 > gpurun_out/s40_c5_n8.json 2> gpurun_out/s40_c5_n8.err
tail -3 gpurun_out/s40_c5_n8.err | cut -c1-300; python - <<PY
import json
d=json.loads(open("gpurun_out/s40_c5_n8.json").read().strip().splitlines()[-1])
print(d["value"]/1e9, d["ms_per_step"], d["build"]["gbp_per_s"], d["parity"], d["fused_count_exchange"])
PY
