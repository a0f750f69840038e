#!/bin/bash
# round-2 GPU call 12: chunking only for long K + probe-free rulebook re-ordering
mkdir -p gpurun_out
echo "== ops (sparse, sort, tma, long-K)"
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -s 2>&1 | grep -E "long-K|passed|failed|Error|error" | tee gpurun_out/c12_ops.log
echo "== e2e + fullsize"
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -5 | tee gpurun_out/c12_e2e.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c12_bench.err | tail -1 > gpurun_out/c12_bench.json
grep -E "ms  x" gpurun_out/c12_bench.err | head -8
python -c "
import json; d=json.load(open('gpurun_out/c12_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
