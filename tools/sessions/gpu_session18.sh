#!/bin/bash
# re-entry sanity: full GPU parity suite, smoke, default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.txt 2>&1; tail -3 gpurun_out/s18_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/s18_smoke.txt 2>&1; tail -2 gpurun_out/s18_smoke.txt
timeout 900 python bench.py > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err; tail -c 600 gpurun_out/s18_bench.json
