#!/bin/bash
# full GPU suite, smoke, default bench line, launch list + ncu --set full captures of the current kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s27_pytest.txt 2>&1; tail -3 gpurun_out/s27_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/s27_smoke.txt 2>&1; tail -1 gpurun_out/s27_smoke.txt
timeout 900 python bench.py > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err; tail -c 300 gpurun_out/s27_bench.json; echo
KRE='regex:readid_|kmerize_|query_|sched_|transpose_|region_|table_|rownz_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/s27_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s27_ncu_bench.log 2>&1
tail -1 gpurun_out/s27_launches.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:query_(gather|front)" -s 4 -c 6 -f -o gpurun_out/prof_query_r1g \
   python bench.py --only-search > gpurun_out/s27_ncu_query.log 2>&1
ls -la gpurun_out | grep -E "r1g|s27"
