#!/bin/bash
# 2-GPU check: tests, bench under torchrun (N=2), C5 column-sharded workload with NCCL
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s10_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.txt 2>&1; tail -3 gpurun_out/s10_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-search > gpurun_out/s10_bench_n2.json 2> gpurun_out/s10_bench_n2.err
tail -c 600 gpurun_out/s10_bench_n2.json; echo
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --workload c5 --c5-acc 2500 --steps 3 --warmup 1 > gpurun_out/s10_c5_n2.json 2> gpurun_out/s10_c5_n2.err
tail -5 gpurun_out/s10_c5_n2.err; cat gpurun_out/s10_c5_n2.json
timeout 600 python bench.py --workload c5 --c5-acc 1250 --steps 3 --warmup 1 > gpurun_out/s10_c5_n1.json 2> gpurun_out/s10_c5_n1.err
tail -3 gpurun_out/s10_c5_n1.err; cat gpurun_out/s10_c5_n1.json
timeout 600 python bench.py --only-search > gpurun_out/s10_c3.json 2> gpurun_out/s10_c3.err; cat gpurun_out/s10_c3.json
