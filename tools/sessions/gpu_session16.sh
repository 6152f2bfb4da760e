#!/bin/bash
# profiling artefacts of the current kernels: filtered launch list of the bench command, ncu --set full captures
mkdir -p gpurun_out
KRE='regex:readid_|kmerize_|query_|sched_|transpose_|region_|table_|rownz_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/s16_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-search --no-cpu-baseline > gpurun_out/s16_ncu_bench.log 2>&1
tail -2 gpurun_out/s16_launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:readid_(order|kmerize|vote)" -s 10 -c 5 -f -o gpurun_out/prof_readid_r1e \
   python bench.py --steps 1 --warmup 2 --no-search --no-cpu-baseline > gpurun_out/s16_ncu_readid.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:query_(gather|hash)" -s 2 -c 2 -f -o gpurun_out/prof_query_r1e \
   python bench.py --only-search > gpurun_out/s16_ncu_query.log 2>&1
ls -la gpurun_out | grep -E "r1e|s16"
