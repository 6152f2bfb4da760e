#!/bin/bash
# reads up to 1000 bases, fuzz over the new search paths, everything else still green
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/s56_pytest.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s56_pytest.txt | head -30 | cut -c1-300
