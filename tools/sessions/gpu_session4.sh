#!/bin/bash
# re-entry check: GPU parity tests, smoke, full bench, ncu launch list, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s4_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s4_pytest.txt
tail -5 gpurun_out/s4_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.txt 2>&1; tail -2 gpurun_out/s4_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
tail -c 3000 gpurun_out/s4_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-search --no-cpu-baseline > gpurun_out/s4_ncu_bench.log 2>&1
tail -3 gpurun_out/s4_launches.csv
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4_ref.json 2> gpurun_out/s4_ref.err
cat gpurun_out/s4_ref.json
nproc
