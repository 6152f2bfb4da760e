#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload c4 --c4-acc 12 --steps 8 --warmup 2 > gpurun_out/s12_c4.json 2> gpurun_out/s12_c4.err
tail -5 gpurun_out/s12_c4.err; cat gpurun_out/s12_c4.json
