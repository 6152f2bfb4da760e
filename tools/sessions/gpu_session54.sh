#!/bin/bash
# column-sharded default report: one-process and two-process tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 240 -k "one_process" > gpurun_out/s54_one.txt 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/s54_one.txt | head -40 | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 240 -k "not one_process" > gpurun_out/s54_two.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s54_two.txt | head -60 | cut -c1-300
