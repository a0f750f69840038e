#!/bin/bash
# round-2 GPU call 33: occupancy of the rulebook kernels (blocks per SM) vs the slow-down of the concurrent gather-GEMMs
mkdir -p gpurun_out
for b in 16 32 64 128; do
FF3D_CHAIN_BPS=$b timeout 600 python bench.py --steps 20 --warmup 5 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c33_bench_$b.err | tail -1 > gpurun_out/c33_bench_$b.json
python -c "
import json; d=json.load(open('gpurun_out/c33_bench_$b.json')); print('bps=$b', round(d['value'],1), round(d['e2e']['value'],1), d['stage_ms']['sparse_encoder'])"
done
