#!/bin/bash
# ncu --set full (with source) on the C3 search kernels after the query_front change
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:query_(gather|front)" -s 4 -c 5 -f -o gpurun_out/prof_query_r1f \
   python bench.py --only-search > gpurun_out/s24_ncu_query.log 2>&1
tail -2 gpurun_out/s24_ncu_query.log | cut -c1-300
ls -la gpurun_out | grep r1f
