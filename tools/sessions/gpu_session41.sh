#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/s41_pytest.txt 2>&1; tail -25 gpurun_out/s41_pytest.txt | cut -c1-250
