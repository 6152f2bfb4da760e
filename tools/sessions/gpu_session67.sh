#!/bin/bash
# ncu --set full: the wide-row vote kernel (1,000 accessions) and the read-set search kernels of the c1 workload (one launch each)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"readid_vote_wide" -s 3 -c 1 -f -o gpurun_out/prof_vote_wide \
   python bench.py --n-acc 1000 --batch-pairs 200000 --pool-batches 2 --steps 1 --warmup 3 --no-search --no-cpu-baseline --rep-cap 128 > gpurun_out/ncu_vote_wide.log 2>&1
tail -2 gpurun_out/ncu_vote_wide.log | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"(region_compact|kmerize_insert_kernel<0, 0>|query_counts_kernel|uniq_hist|slots)" --kernel-name-base demangled -s 5 -c 5 -f -o gpurun_out/prof_c1 \
   python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1
tail -2 gpurun_out/ncu_c1.log | cut -c1-200
python profiles/summarize_ncu.py gpurun_out/prof_vote_wide.ncu-rep > gpurun_out/s67_vote_wide.txt 2>&1; head -30 gpurun_out/s67_vote_wide.txt
python profiles/summarize_ncu.py gpurun_out/prof_c1.ncu-rep > gpurun_out/s67_c1.txt 2>&1; grep -E "^==|duration|dram read|dram write|throughput" gpurun_out/s67_c1.txt | head -40
ls -la gpurun_out/*.ncu-rep | tail -3
