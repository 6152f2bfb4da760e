# e2e of the read_id host pipeline against its chunk schedule (bench.py --opt ...): tools for profiles/r2_pipeline_sweep.txt
for ser in 1; do for c0 in 32768 65536 131072; do for c in 131072 262144 524288; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-search --opt readid_serialize=$ser --opt readid_chunk0_reads=$c0 --opt readid_chunk_reads=$c > gpurun_out/sw.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/sw.json").read().strip().split("\n")[-1])
print("serialize $ser chunk0 $c0 chunk $c value %.1f e2e %.1f ascii %.1f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_ascii"]["value"]/1e6))
PY
done; done; done
