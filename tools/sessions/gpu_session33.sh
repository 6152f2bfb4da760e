#!/bin/bash
# ncu --set full on the C4 build kernels (count table of a 30x read set)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:kmerize_insert|region_to_bloom|region_histogram" -s 0 -c 14 -f -o gpurun_out/prof_c4_r1g \
   python bench.py --workload c4 --c4-acc 6 --steps 1 --warmup 1 > gpurun_out/s33_ncu.log 2>&1
tail -2 gpurun_out/s33_ncu.log | cut -c1-300
