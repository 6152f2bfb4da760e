#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/cli_e2e.py 1000000 > gpurun_out/s38_cli.json 2> gpurun_out/s38_cli.err; grep "^build\|^read_id\|^search\|^batch" -A14 gpurun_out/s38_cli.err | head -90; cat gpurun_out/s38_cli.json
