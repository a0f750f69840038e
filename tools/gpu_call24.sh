#!/bin/bash
# round-2 GPU call 24: racecheck after the __syncwarp fix in the radix scatter; bench with the all-lane barrier poll
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -x 2>&1 | tail -3 | tee gpurun_out/c24_ops.log
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c24_bench.err | tail -1 > gpurun_out/c24_bench.json
grep -E "ms  x" gpurun_out/c24_bench.err | grep spconv | head -4
python -c "
import json; d=json.load(open('gpurun_out/c24_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms']['sparse_encoder'])"
echo "== racecheck"
timeout 700 compute-sanitizer --tool racecheck --print-limit 450 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/c24_racecheck.log 2>&1
echo "rc=$?" >> gpurun_out/c24_racecheck.log; tail -3 gpurun_out/c24_racecheck.log
grep -E "^========= (Error|Warning)" gpurun_out/c24_racecheck.log | sed 's/at __shared__ 0x[0-9a-f]* in block.*//' | sort | uniq -c | sort -rn | head
grep -oE "at ff3d::[a-z_0-9A-Z]+" gpurun_out/c24_racecheck.log | sort | uniq -c | sort -rn | head -20
