#!/bin/bash
# round-2 GPU call 19: cp.async gather with the neighbour maps staged in shared memory by an index warp
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -x 2>&1 | tail -4 | tee gpurun_out/c19_ops.log
if grep -q "failed\|rror\|Timeout" gpurun_out/c19_ops.log; then echo "ops failed: stopping"; exit 1; fi
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c19_e2e.log
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c19_bench.err | tail -1 > gpurun_out/c19_bench.json
grep -E "ms  x" gpurun_out/c19_bench.err | grep spconv | head -9
python -c "
import json; d=json.load(open('gpurun_out/c19_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
