#!/bin/bash
# round-2 GPU call 3: unified tcgemm (fp16 hi/lo default + TF32) with per-tile tap skipping
mkdir -p gpurun_out
echo "== pytest gpu (f16 default)"
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -x 2>&1 | tail -25 | tee gpurun_out/c3_tests_f16.log
echo "== pytest gpu ops+e2e (tf32)"
timeout 600 env FF3D_GEMM=tf32 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_camera.py -q -m gpu --timeout 300 2>&1 | tail -8 | tee gpurun_out/c3_tests_tf32.log
echo "== full size + accuracy tables (f16)"
timeout 1500 env FF3D_SLOW_TESTS=1 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_accuracy.py "tests/test_gpu_ops.py::test_tcgemm_long_k_error_budget" -q -m gpu --timeout 900 -s 2>&1 | grep -E "long-K|img_feat|camera_bev|parity:|accuracy|dense_heatmap|sparse_bev|second|conv_feat|stage_feat|extra|passed|failed|Error|error" | cut -c1-1500 | tee gpurun_out/c3_fullsize_f16.log
echo "== bench f16 (steps 10)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/c3_bench_f16.err | tail -1 > gpurun_out/c3_bench_f16.json
python - <<'P'
import json
d=json.load(open('gpurun_out/c3_bench_f16.json'))
print(d['value'], d['e2e'], d['stage_ms'])
P
grep -E "ms  x" gpurun_out/c3_bench_f16.err | head -24
echo "== bench tf32 (steps 10)"
timeout 600 env FF3D_GEMM=tf32 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/c3_bench_tf32.err | tail -1 > gpurun_out/c3_bench_tf32.json
python - <<'P'
import json
d=json.load(open('gpurun_out/c3_bench_tf32.json'))
print(d['value'], d['e2e'], d['stage_ms'])
P
grep -E "ms  x" gpurun_out/c3_bench_tf32.err | head -12
