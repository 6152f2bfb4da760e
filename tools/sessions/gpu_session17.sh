#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli_gpu.py -m gpu -x -q > gpurun_out/s17_pytest.txt 2>&1; tail -3 gpurun_out/s17_pytest.txt
timeout 1500 python tools/cli_e2e.py 1000000 > gpurun_out/s17_cli.json 2> gpurun_out/s17_cli.err; grep -A12 "^build\|^read_id\|^search" gpurun_out/s17_cli.err | head -60; cat gpurun_out/s17_cli.json
