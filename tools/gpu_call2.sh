#!/bin/bash
# round-2 GPU call 2: mask-sorted rulebooks (old GEMM, no tap skipping yet) + sort tests + full-size parity / accuracy tables
mkdir -p gpurun_out
echo "== pytest gpu ops (new rulebook API)"
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -q -m gpu --timeout 300 -x 2>&1 | tail -25 | tee gpurun_out/c2_tests.log
echo "== full size + accuracy (tf32)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_accuracy.py -q -m gpu --timeout 400 -s 2>&1 | grep -E "parity:|accuracy|dense_heatmap|sparse_bev|second|conv_feat|stage_feat|extra|passed|failed|Error|error" | cut -c1-1500 | tee gpurun_out/c2_fullsize_tf32.log
cp gpurun_out/accuracy_vs_fp64.json gpurun_out/c2_accuracy_tf32.json 2>/dev/null
echo "== full size + accuracy (f16 harness)"
timeout 900 env FF3D_EXPERIMENTAL_F16=1 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_accuracy.py -q -m gpu --timeout 400 -s 2>&1 | grep -E "parity:|accuracy|dense_heatmap|sparse_bev|second|conv_feat|stage_feat|extra|passed|failed|Error|error" | cut -c1-1500 | tee gpurun_out/c2_fullsize_f16.log
cp gpurun_out/accuracy_vs_fp64.json gpurun_out/c2_accuracy_f16.json 2>/dev/null
echo "== bench (steps 10)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/c2_bench.err | tail -1 > gpurun_out/c2_bench.json
python - <<'P'
import json
d=json.load(open('gpurun_out/c2_bench.json'))
print(d['value'], d['e2e'], d['stage_ms'])
P
grep -E "ms  x" gpurun_out/c2_bench.err | head -30
