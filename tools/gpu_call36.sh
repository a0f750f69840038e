#!/bin/bash
# final bench lines of the committed tree (default flags + reference arm)
mkdir -p gpurun_out
timeout 900 python bench.py 2> gpurun_out/ev4_bench.err | tail -1 > gpurun_out/ev4_bench.json
python -c "
import json; d=json.load(open('gpurun_out/ev4_bench.json')); print(d['value'], d['e2e']['value'], d['parity']['pass'], d['stage_ms'], d['bs_sweep'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/ev4_bench_ref.err | tail -1 > gpurun_out/ev4_bench_ref.json
cut -c1-200 gpurun_out/ev4_bench_ref.json
