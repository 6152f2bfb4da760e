#!/bin/bash
# e2e read_id pipeline: kernels of consecutive chunks serialised (copies still overlap) vs free overlap; chunk size sweep
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 6 --warmup 3 --no-search --no-cpu-baseline "$@" > gpurun_out/s31_$tag.json 2> gpurun_out/s31_$tag.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s31_$tag.json").read().strip().splitlines()[-1])
print("$tag", "value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6))
PY
}
run serial
run overlap --opt readid_serialize=0
run serial_c128k --opt readid_chunk_reads=131072
run serial_c512k --opt readid_chunk_reads=524288
run overlap_c512k --opt readid_serialize=0 --opt readid_chunk_reads=524288
