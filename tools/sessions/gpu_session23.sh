#!/bin/bash
# query_front (shared-memory dedup front end of small queries): parity + C3 measurement
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -x -q > gpurun_out/s23_pytest.txt 2>&1; tail -15 gpurun_out/s23_pytest.txt
timeout 600 python bench.py --only-search > gpurun_out/s23_c3.json 2> gpurun_out/s23_c3.err; tail -3 gpurun_out/s23_c3.err; cat gpurun_out/s23_c3.json | cut -c1-1800

