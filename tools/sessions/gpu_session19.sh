#!/bin/bash
# minimizer (.mxi) device path: parity tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_minimizer.py -m gpu -x -q > gpurun_out/s19_pytest.txt 2>&1; tail -30 gpurun_out/s19_pytest.txt
