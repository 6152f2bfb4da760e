#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/h2d_bw.txt 2>&1
import torch, time
for mb in (64, 256, 600):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(mb << 20, dtype=torch.uint8, device='cuda')
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    t = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt2 = (time.perf_counter() - t) / 5
    print(f"{mb} MB pinned: H2D {mb/1024/dt:.1f} GB/s, D2H {mb/1024/dt2:.1f} GB/s")
import os
print("cpus", os.cpu_count())
PY
cat gpurun_out/h2d_bw.txt
CID_TRACE=1 python bench.py --no-search --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/trace_bench.json 2> gpurun_out/trace_bench.err
grep "cid trace" gpurun_out/trace_bench.err | tail -6
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
