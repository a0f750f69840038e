#!/bin/bash
# round-2 GPU call 30: per-level timeline of the sparse encoder (main stream: waits for the rulebook streams vs gather-GEMMs)
mkdir -p gpurun_out
for o in 1 0; do
FF3D_SPARSE_MARKS=1 FF3D_SPARSE_OVERLAP=$o timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c30_bench_$o.err | tail -1 > gpurun_out/c30_bench_$o.json
python -c "
import json; d=json.load(open('gpurun_out/c30_bench_$o.json')); print('overlap=$o', d['value'], {k:v for k,v in d['stage_ms'].items() if k.startswith('sp') or k in ('sparse_encoder','voxelize+vfe')})"
done
