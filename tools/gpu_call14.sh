#!/bin/bash
# round-2 GPU call 14: one-launch decoder stage (decstage.cu) vs the unfused path, then the parity suites and the bench
mkdir -p gpurun_out
echo "== decoder stage"
timeout 300 python -m pytest tests/test_gpu_decoder_stage.py -q -m gpu --timeout 120 -x 2>&1 | tail -15 | tee gpurun_out/c14_decstage.log
if grep -q "failed\|error\|Error\|Timeout" gpurun_out/c14_decstage.log; then echo "decoder stage test failed: stopping"; exit 1; fi
echo "== e2e + fullsize"
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -8 | tee gpurun_out/c14_e2e.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c14_bench.err | tail -1 > gpurun_out/c14_bench.json
grep -E "ms  x" gpurun_out/c14_bench.err | head -12
python -c "
import json; d=json.load(open('gpurun_out/c14_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
