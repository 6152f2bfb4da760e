#!/bin/bash
# end-of-session validation: every GPU test, smoke, default bench line, reference arm, ncu launch list of the c1 workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/s66_pytest.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s66_pytest.txt | head -20 | cut -c1-250
for i in 1 2 3 4 5; do timeout 200 python -m pytest tests/test_cli_gpu.py -m gpu -x -q -k batch_id 2>&1 | tail -1; done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/s66_bench.json 2> gpurun_out/s66_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/s66_bench.json").read().strip().splitlines()[-1])
print("value %.1fM e2e %.1fM ms %.2f launches %d frac %.3f ra %.3f cpu %.0f (%d cores, match %s)"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["random_access"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["matches_gpu_classification"]))
s=d["search_c3"]; print("C3 %.2f G lookups/s frac %.3f"%(s["lookups_per_s"]/1e9, s["roofline"]["frac"]), d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:(kmerize|region|query|uniq|table_clear|transpose|rownz|slots)" -c 300 --csv --log-file gpurun_out/s66_c1_launches.csv python bench.py --workload c1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s66_c1_ncu.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/s66_c1_launches.csv
