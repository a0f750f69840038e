#!/bin/bash
# round-2 GPU call 18: cp.async gather with 4 vs 8 producer warps (arrival count per stage 129 vs 257)
mkdir -p gpurun_out
for n in 8 4; do
  echo "== bench FF3D_CPA_NPW=$n"
  FF3D_CPA_NPW=$n timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c18_bench_npw$n.err | tail -1 > gpurun_out/c18_bench_npw$n.json
  grep -E "ms  x" gpurun_out/c18_bench_npw$n.err | grep spconv | head -9
  python -c "
import json; d=json.load(open('gpurun_out/c18_bench_npw$n.json')); print(d['value'], d['e2e']['value'], d['stage_ms']['sparse_encoder'])"
done
FF3D_CPA_NPW=4 timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 200 -k "sparse or tma" 2>&1 | tail -3
