#!/bin/bash
# round-2 GPU call 10: down_sites without a contended counter, aggregated sort histograms, lane-per-query mha_core, warp-per-point roi_sample
mkdir -p gpurun_out
echo "== ops"
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 2>&1 | tail -8 | tee gpurun_out/c10_ops.log
echo "== e2e + fullsize L"
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 -s -k "not fusion_lc_full and not waymo_l_full" 2>&1 | grep -E "parity:|passed|failed|Error|error|assert" | cut -c1-600 | tee gpurun_out/c10_e2e.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c10_bench.err | tail -1 > gpurun_out/c10_bench.json
grep -E "ms  x" gpurun_out/c10_bench.err | head -12
python -c "
import json; d=json.load(open('gpurun_out/c10_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
