#!/bin/bash
# experiment: kmerize/order of chunk c+1 sharing the SMs with the vote kernel of chunk c (two streams, capped grids)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 6 --warmup 3 --no-search --no-cpu-baseline "$@" > gpurun_out/s21_$tag.json 2> gpurun_out/s21_$tag.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s21_$tag.json").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("$tag", "value %.1fM ms/step %.2f e2e %.1fM"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6), {n:round(v["ms_per_launch"]*v["launches_per_step"],2) for n,v in k.items()})
PY
}
run base
run s2 --opt readid_streams=2
run s2_k3_v6 --opt readid_streams=2 --opt readid_kmerize_ctas=3 --opt readid_vote_ctas=6
run s2_k3_v8 --opt readid_streams=2 --opt readid_kmerize_ctas=3 --opt readid_vote_ctas=8
run s2_k4_v6 --opt readid_streams=2 --opt readid_kmerize_ctas=4 --opt readid_vote_ctas=6
run s2_k2_v8 --opt readid_streams=2 --opt readid_kmerize_ctas=2 --opt readid_vote_ctas=8
run s2_k4_v4 --opt readid_streams=2 --opt readid_kmerize_ctas=4 --opt readid_vote_ctas=4
run s2_k6_v6 --opt readid_streams=2 --opt readid_kmerize_ctas=6 --opt readid_vote_ctas=6
