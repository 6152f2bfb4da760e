#!/bin/bash
# ncu --set full with source on readid_kmerize / readid_order (C2)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:readid_(kmerize|order_small_kernel<256>)" -s 4 -c 2 -f -o gpurun_out/prof_readid_r1g \
   python bench.py --steps 1 --warmup 2 --no-search --no-cpu-baseline > gpurun_out/s28_ncu.log 2>&1
tail -2 gpurun_out/s28_ncu.log | cut -c1-200
