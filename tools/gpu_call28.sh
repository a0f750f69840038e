#!/bin/bash
# round-2 GPU call 28: branch-free epilogue arithmetic + bias staged in shared memory (tmagemm)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -x 2>&1 | tail -8 | tee gpurun_out/c28_ops.log
if grep -q "failed\|rror\|Timeout" gpurun_out/c28_ops.log; then echo "ops failed: stopping"; exit 1; fi
timeout 1500 python -m pytest tests/test_gpu_decoder_stage.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_camera.py -q -m gpu --timeout 400 2>&1 | tail -6 | tee gpurun_out/c28_e2e.log
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c28_bench.err | tail -1 > gpurun_out/c28_bench.json
grep -E "ms  x" gpurun_out/c28_bench.err | head -16
python -c "
import json; d=json.load(open('gpurun_out/c28_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
