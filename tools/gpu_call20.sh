#!/bin/bash
# round-2 GPU call 20: index warp staging the neighbour maps by 4-byte cp.async; bench; ncu of the BN=64 / BN=16 gather kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -x 2>&1 | tail -4 | tee gpurun_out/c20_ops.log
if grep -q "failed\|rror\|Timeout" gpurun_out/c20_ops.log; then echo "ops failed: stopping"; exit 1; fi
timeout 1200 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c20_e2e.log
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c20_bench.err | tail -1 > gpurun_out/c20_bench.json
grep -E "ms  x" gpurun_out/c20_bench.err | grep spconv | head -9
python -c "
import json; d=json.load(open('gpurun_out/c20_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
FF3D_SPARSE_OVERLAP=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'tmagemm_kernel<\(int\)2, \(int\)(64|16)' -s 3 -c 4 -o gpurun_out/c20_sparse -f python tools/profile_forward.py 1 > gpurun_out/c20_ncu.log 2>&1
tail -1 gpurun_out/c20_ncu.log
