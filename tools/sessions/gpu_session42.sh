#!/bin/bash
# e2e read_id pipeline: first-chunk size and chunk ceiling sweep
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-search --no-cpu-baseline "$@" > gpurun_out/s42_$tag.json 2> gpurun_out/s42_$tag.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s42_$tag.json").read().strip().splitlines()[-1])
print("$tag", "value %.1fM e2e %.1fM"%(d["value"]/1e6,d["e2e"]["value"]/1e6))
PY
}
run c0_32k
run c0_48k --opt readid_chunk0_reads=49152
run c0_64k --opt readid_chunk0_reads=65536
run c0_24k --opt readid_chunk0_reads=24576
run c0_48k_max384k --opt readid_chunk0_reads=49152 --opt readid_chunk_reads=393216
run c0_64k_max512k --opt readid_chunk0_reads=65536 --opt readid_chunk_reads=524288
