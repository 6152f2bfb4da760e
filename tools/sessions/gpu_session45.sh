#!/bin/bash
# compute-sanitizer passes over the smoke path and a slice of the parity tests (round-1 hygiene check)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 500 $CS --tool $tool --print-limit 20 --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/s45_${tool}_smoke.txt 2>&1
  echo "$tool smoke rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/s45_${tool}_smoke.txt | head -8
done
timeout 900 $CS --tool memcheck --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "perfect or edge or minimizer or sharded or stride" > gpurun_out/s45_memcheck_parity.txt 2>&1
echo "memcheck parity rc=$?"; tail -5 gpurun_out/s45_memcheck_parity.txt | cut -c1-300
CID_TRACE=1 timeout 300 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/s45_c1_trace.json 2> gpurun_out/s45_c1_trace.err; grep "cid trace" gpurun_out/s45_c1_trace.err | tail -3
