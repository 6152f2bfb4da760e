#!/bin/bash
# 8-GPU C2 line with ranks bound to their GPU's NUMA node (e2e was 210.7M pairs/s unbound), then C5 with the new front end
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/s26_topo.txt 2>&1
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-search > gpurun_out/s26_c2_n8.json 2> gpurun_out/s26_c2_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s26_c2_n8.json").read().strip().splitlines()[-1])
print("N=8 value %.1fM e2e %.1fM"%(d["value"]/1e6, d["e2e"]["value"]/1e6), d["e2e"].get("host_cores_bound_per_rank"), d["cpu_baseline"]["cores"])
PY
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --workload c5 --steps 3 --warmup 1 > gpurun_out/s26_c5_n8.json 2> gpurun_out/s26_c5_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/s26_c5_n8.json").read().strip().splitlines()[-1])
print("C5 N=8", d["value"]/1e9, d["ms_per_step"], d["build"]["gbp_per_s"], d["parity"], {k:round(v["ms_per_launch"]*v["launches_per_pass"],2) for k,v in d["kernels"].items()})
PY
