#!/bin/bash
# round-2 GPU call 16: 16-warp decoder stage with batched MSDA loads; bench; ncu of the stage kernel + the sparse gather-GEMMs
mkdir -p gpurun_out
echo "== decoder stage"
timeout 300 python -m pytest tests/test_gpu_decoder_stage.py -q -m gpu --timeout 120 -x 2>&1 | tail -15 | tee gpurun_out/c16_decstage.log
if grep -q "failed\|error\|Error\|Timeout" gpurun_out/c16_decstage.log; then echo "decoder stage test failed: stopping"; exit 1; fi
echo "== e2e + fullsize"
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -8 | tee gpurun_out/c16_e2e.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c16_bench.err | tail -1 > gpurun_out/c16_bench.json
grep -E "ms  x" gpurun_out/c16_bench.err | head -6
python -c "
import json; d=json.load(open('gpurun_out/c16_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'decoder_stage_kernel' -c 1 \
  -o gpurun_out/c16_decstage -f python tools/profile_forward.py 1 > gpurun_out/c16_ncu_a.log 2>&1
tail -1 gpurun_out/c16_ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'tmagemm_kernel<\(int\)2, \(int\)(64|32|128)' -s 2 -c 9 -o gpurun_out/c16_sparse -f python tools/profile_forward.py 1 > gpurun_out/c16_ncu_b.log 2>&1
tail -1 gpurun_out/c16_ncu_b.log
