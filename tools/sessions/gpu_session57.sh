#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 400 -k "read_id" > gpurun_out/s57.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s57.txt | head -30 | cut -c1-400
