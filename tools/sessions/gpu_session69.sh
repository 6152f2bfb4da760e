#!/bin/bash
# last check of the session: the rebuilt library on a fresh box (smoke + the quickest parity files)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q --timeout 300 2>&1 | tail -1
