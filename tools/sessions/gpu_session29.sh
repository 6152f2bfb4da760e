#!/bin/bash
# 2 GPUs: batch_id with one worker per GPU; CLI suites
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_cli_minimizer_gpu.py -m gpu -x -q > gpurun_out/s29_pytest.txt 2>&1; tail -5 gpurun_out/s29_pytest.txt
