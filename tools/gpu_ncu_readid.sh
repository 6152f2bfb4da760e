#!/bin/bash
# ncu --set full on the read_id kernels (one launch each), C2 workload
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"readid_(order|kmerize|vote)" -s 6 -c 3 -f -o gpurun_out/prof_readid \
   python bench.py --steps 1 --warmup 2 --no-search --no-cpu-baseline > gpurun_out/ncu_readid.log 2>&1
tail -3 gpurun_out/ncu_readid.log
ls -la gpurun_out/
