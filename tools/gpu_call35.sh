#!/bin/bash
# round-2 GPU call 35: scheduling of the rulebook items: 0 = as before, 1 = site sets on their own stream, 2 = 1 + just-in-time release
mkdir -p gpurun_out
for m in 0 1 2; do
FF3D_PLAN_MODE=$m FF3D_SPARSE_MARKS=1 timeout 600 python bench.py --steps 20 --warmup 5 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c35_bench_$m.err | tail -1 > gpurun_out/c35_bench_$m.json
python -c "
import json; d=json.load(open('gpurun_out/c35_bench_$m.json')); print('mode=$m', round(d['value'],1), round(d['e2e']['value'],1), {k:v for k,v in d['stage_ms'].items() if k.startswith('sp')})"
done
FF3D_PLAN_MODE=2 timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -3
