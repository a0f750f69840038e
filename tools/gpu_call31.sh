#!/bin/bash
# round-2 GPU call 31: site sets of the deeper levels on their own stream (input-level maps no longer queued behind them)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_camera.py -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c31_e2e.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c31_bench_$i.err | tail -1 > gpurun_out/c31_bench_$i.json
python -c "
import json; d=json.load(open('gpurun_out/c31_bench_$i.json')); print(d['value'], d['e2e']['value'], d['eager_ms_per_step'], d['stage_ms']['sparse_encoder'])"
done
