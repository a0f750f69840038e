#!/bin/bash
# round-2 GPU call 29: radix sort with 1024-element blocks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -x 2>&1 | tail -4 | tee gpurun_out/c29_ops.log
if grep -q "failed\|rror\|Timeout" gpurun_out/c29_ops.log; then echo "ops failed: stopping"; exit 1; fi
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c29_e2e.log
for o in 1 0; do
FF3D_SPARSE_OVERLAP=$o timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c29_bench_$o.err | tail -1 > gpurun_out/c29_bench_$o.json
python -c "
import json; d=json.load(open('gpurun_out/c29_bench_$o.json')); print('overlap=$o', d['value'], d['e2e']['value'], d['stage_ms']['sparse_encoder'])"
done
