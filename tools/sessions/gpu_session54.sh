#!/bin/bash
# column-sharded default report: one-process and two-process tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 240 > gpurun_out/s54_sharded.txt 2>&1; tail -30 gpurun_out/s54_sharded.txt | cut -c1-400
