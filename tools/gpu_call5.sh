#!/bin/bash
# round-2 GPU call 5: TMA-fed GEMM (split activations) -- kernel tests, then e2e / full-size parity, then bench
mkdir -p gpurun_out
echo "== TMA kernel tests"
echo skip
echo "== all ops"
echo skip
echo "== e2e + fullsize"
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_camera.py -q -m gpu --timeout 300 -s 2>&1 | grep -E "parity:|passed|failed|Error|error|assert" | cut -c1-1200 | tee gpurun_out/c5_e2e.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-gpu-standin 2> gpurun_out/c5_bench.err | tail -1 > gpurun_out/c5_bench.json
grep -E "ms  x" gpurun_out/c5_bench.err | head -30
python - <<'P'
import json
d=json.load(open('gpurun_out/c5_bench.json'))
for k in ('value','ms_per_step','e2e','eager_ms_per_step','launches_per_step','stage_ms','parity'):
    print(k, d.get(k))
P
