#!/bin/bash
# minimizer CLI tests + whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s20_pytest.txt 2>&1; tail -30 gpurun_out/s20_pytest.txt
