#!/bin/bash
# 8-GPU: C5 at full size (10,000 accessions x 5 Mbp, 62.6 GB signature matrix column-sharded over 8 GPUs, NCCL count gather)
# and the default C2 line at N=8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s22_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus 8 --workload c5 --steps 3 --warmup 1 > gpurun_out/s22_c5_n8.json 2> gpurun_out/s22_c5_n8.err
tail -5 gpurun_out/s22_c5_n8.err; cat gpurun_out/s22_c5_n8.json
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-search > gpurun_out/s22_c2_n8.json 2> gpurun_out/s22_c2_n8.err
tail -c 900 gpurun_out/s22_c2_n8.json
