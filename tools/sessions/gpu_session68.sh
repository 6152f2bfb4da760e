#!/bin/bash
# wide-row vote kernel (indexes with more than 64 accessions): lane-parallel hashing vs the previous build, 1,000-accession index;
# parity of everything that touches it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_sharded.py tests/test_gpu_minimizer.py -m gpu -x -q --timeout 400 -k "read_id or readid or fuzz or sharded" > gpurun_out/s68_pytest.txt 2>&1; grep -E "^E  |passed|failed" gpurun_out/s68_pytest.txt | head -20 | cut -c1-300
run() {
  timeout 600 python bench.py --n-acc 1000 --batch-pairs 200000 --pool-batches 4 --steps 4 --warmup 2 --no-search --no-cpu-baseline --rep-cap 128 > gpurun_out/s68_$1.json 2> gpurun_out/s68_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s68_$1.json").read().strip().splitlines()[-1])
print("$1", "value %.2fM pairs/s e2e %.2fM ms/step %.2f"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), {k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()}, "trunc", d["report_truncated_reads_last_step"])
PY
}
run new
run2() {
  timeout 600 python bench.py --n-acc 2000 --batch-pairs 200000 --pool-batches 4 --steps 4 --warmup 2 --no-search --no-cpu-baseline --rep-cap 128 > gpurun_out/s68_n2000.json 2> gpurun_out/s68_n2000.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s68_n2000.json").read().strip().splitlines()[-1])
print("n2000", "value %.2fM pairs/s e2e %.2fM ms/step %.2f"%(d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), {k:round(v["ms_per_launch"],2) for k,v in d["roofline"]["kernels"].items()}, "trunc", d["report_truncated_reads_last_step"])
PY
}
run2
