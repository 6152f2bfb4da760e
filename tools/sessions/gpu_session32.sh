#!/bin/bash
mkdir -p gpurun_out
CID_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-search --no-cpu-baseline > gpurun_out/s32.json 2> gpurun_out/s32.err
grep "cid trace" gpurun_out/s32.err | tail -14
