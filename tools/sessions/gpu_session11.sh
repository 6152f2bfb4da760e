#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "read_id" > gpurun_out/s11_pytest.txt 2>&1; tail -3 gpurun_out/s11_pytest.txt
timeout 900 python tools/l2_probe.py > gpurun_out/s11_l2.txt 2>&1; cat gpurun_out/s11_l2.txt
