#!/bin/bash
# round-2 GPU call 11: chunked accumulation, LSS bevencode on TMA, mha_core key split, division-free site insertion
mkdir -p gpurun_out
echo "== ops"
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -s 2>&1 | grep -E "long-K|passed|failed|Error|error" | tee gpurun_out/c11_ops.log
echo "== e2e + fullsize + camera"
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_camera.py tests/test_gpu_preprocess.py tests/test_gpu_reference_ext.py -q -m gpu --timeout 400 -s 2>&1 | grep -E "parity:|passed|failed|Error|error|assert" | cut -c1-1100 | tee gpurun_out/c11_e2e.log
echo "== accuracy tables"
timeout 1200 env FF3D_SLOW_TESTS=1 python -m pytest tests/test_gpu_accuracy.py -q -m gpu --timeout 900 -s 2>&1 | grep -E "accuracy|dense_heatmap|sparse_bev|second|conv_feat|stage_feat|extra|img_feat|camera_bev|passed|failed|Error" | cut -c1-400 | tee gpurun_out/c11_accuracy.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c11_bench.err | tail -1 > gpurun_out/c11_bench.json
grep -E "ms  x" gpurun_out/c11_bench.err | head -8
python -c "
import json; d=json.load(open('gpurun_out/c11_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
echo "== bench lc"
timeout 600 python bench.py --config lc --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/c11_bench_lc.err | tail -1 > gpurun_out/c11_bench_lc.json
python -c "
import json; d=json.load(open('gpurun_out/c11_bench_lc.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
