#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fuzz.py -m gpu -q > gpurun_out/s36_pytest.txt 2>&1; tail -40 gpurun_out/s36_pytest.txt | cut -c1-300
