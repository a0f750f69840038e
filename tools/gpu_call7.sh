#!/bin/bash
# round-2 GPU call 7: cp.async sparse gather + input-side kernels
mkdir -p gpurun_out
echo "== sparse TMA/cp.async tests + preprocess"
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_preprocess.py -q -m gpu --timeout 120 -k "tma_sparse or preprocess or assemble or image" 2>&1 | tail -15 | tee gpurun_out/c7_ops.log
echo "== bench (cp.async gather)"
timeout 600 python bench.py --steps 6 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c7_bench.err | tail -1 > gpurun_out/c7_bench.json
grep -E "spconv|conv3x3s1\[128->128" gpurun_out/c7_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c7_bench.json')); print(d['value'], d['stage_ms'])"
echo "== e2e quick"
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu --timeout 200 -x 2>&1 | tail -5 | tee gpurun_out/c7_e2e.log
