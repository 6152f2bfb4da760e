#!/bin/bash
# gather probe under ncu (DRAM bytes per gather), ncu --set full of the current kernels, filtered launch list
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum"
for g in 0 32 128; do
  timeout 300 ncu $M --clock-control none --csv --log-file gpurun_out/s5_probe_g$g.csv tools/gather_probe.bin $g 50 8 256 > gpurun_out/s5_probe_g$g.txt 2>&1
done
timeout 120 tools/gather_probe.bin 0 50 8 1024 > gpurun_out/s5_probe_plain.txt 2>&1
timeout 120 tools/gather_probe.bin 0 30 128 256 >> gpurun_out/s5_probe_plain.txt 2>&1
timeout 120 tools/gather_probe.bin 0 50 160 256 >> gpurun_out/s5_probe_plain.txt 2>&1
cat gpurun_out/s5_probe_plain.txt
KRE='regex:readid_|kmerize_|query_|sched_|transpose_|region_|table_|rownz_'
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:readid_(order|kmerize|vote|classify)" -s 10 -c 5 -f -o gpurun_out/prof_readid_r1c \
   python bench.py --steps 1 --warmup 2 --no-search --no-cpu-baseline > gpurun_out/s5_ncu_readid.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:query_counts" -s 1 -c 1 -f -o gpurun_out/prof_query_r1c \
   python bench.py --only-search > gpurun_out/s5_ncu_query.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/s5_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-search --no-cpu-baseline > gpurun_out/s5_ncu_bench.log 2>&1
tail -3 gpurun_out/s5_launches.csv
ls -la gpurun_out
