#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload c1 --quick > gpurun_out/s44_c1_quick.json 2> gpurun_out/s44_c1_quick.err; echo "quick rc=$?"; tail -c 600 gpurun_out/s44_c1_quick.err
timeout 600 python bench.py --workload c1 > gpurun_out/s44_c1.json 2> gpurun_out/s44_c1.err; echo "full rc=$?"; tail -c 600 gpurun_out/s44_c1.err
python - <<PY
import json
for f in ("gpurun_out/s44_c1_quick.json","gpurun_out/s44_c1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "Gbp/s %.2f ms %.1f kern_ms %.1f lookups/s %.3g build %.2f"%(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["lookups_per_s"], d["build"]["gbp_per_s"]))
        print({k:round(v["ms_per_launch"]*v["launches_per_step"],2) for k,v in d["kernels"].items()})
        print(d["cpu_baseline"], d["parity"], d["config"]["cutoff_used"])
    except Exception as e: print(f, "ERR", e)
PY
