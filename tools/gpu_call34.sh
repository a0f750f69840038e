#!/bin/bash
# round-2 GPU call 34: GPU-side timeline of the sparse encoder (host given a head start by a spin kernel), overlap on / off
mkdir -p gpurun_out
for o in 1 0; do
FF3D_SPARSE_MARKS=1 FF3D_SPARSE_OVERLAP=$o timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c34_bench_$o.err | tail -1 > gpurun_out/c34_bench_$o.json
python -c "
import json; d=json.load(open('gpurun_out/c34_bench_$o.json')); print('overlap=$o', round(d['value'],1), {k:v for k,v in d['stage_ms'].items()})"
done
