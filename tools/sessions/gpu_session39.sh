#!/bin/bash
# 2 GPUs: C5 with the count exchange fused into the gather kernel (peer stores over NVLink) vs NCCL all_gather
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --workload c5 --c5-acc 2500 --steps 3 --warmup 1 > gpurun_out/s39_c5_n2.json 2> gpurun_out/s39_c5_n2.err
tail -5 gpurun_out/s39_c5_n2.err | cut -c1-300; python - <<PY
import json
d=json.loads(open("gpurun_out/s39_c5_n2.json").read().strip().splitlines()[-1])
print(d["value"]/1e9, d["ms_per_step"], d["fused_count_exchange"])
PY
timeout 600 python bench.py --only-search > gpurun_out/s39_c3.json 2> gpurun_out/s39_c3.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s39_c3.json").read().strip().splitlines()[-1])
print("C3 ms/pass %.2f  G lookups/s %.2f"%(d["ms_per_pass"], d["lookups_per_s"]/1e9), {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["kernels"].items()}, "frac %.3f"%d["roofline"]["frac"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or gene" 2>&1 | tail -2
