#!/bin/bash
# round-2 GPU call 22: probe records row-major ([cap, kvol]) + tap-mask sort on the 18 rarest bits (2 radix passes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -x 2>&1 | tail -4 | tee gpurun_out/c22_ops.log
if grep -q "failed\|rror\|Timeout" gpurun_out/c22_ops.log; then echo "ops failed: stopping"; exit 1; fi
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c22_e2e.log
for d in 9 0 13; do
  echo "== bench FF3D_SORT_DROP=$d"
  FF3D_SORT_DROP=$d timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c22_bench_$d.err | tail -1 > gpurun_out/c22_bench_$d.json
  grep -E "ms  x" gpurun_out/c22_bench_$d.err | grep spconv | head -4
  python -c "
import json; d=json.load(open('gpurun_out/c22_bench_$d.json')); print(d['value'], d['e2e']['value'], d['stage_ms']['sparse_encoder'], d['roofline']['sparse_executed_vs_algorithmic'])"
done
