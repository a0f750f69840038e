#!/bin/bash
# round-2 GPU call 40: stride-2 convs on the TMA-fed kernel (traversal stride in the tensor map) + lattice test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 200 -k "stride2 or lattice or tma_conv2d" 2>&1 | tail -15 | tee gpurun_out/c40_ops.log
if grep -q "failed\|rror\|Timeout" gpurun_out/c40_ops.log; then echo "new tests failed: stopping"; exit 1; fi
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c40_e2e.log
timeout 600 python bench.py --steps 20 --warmup 5 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c40_bench.err | tail -1 > gpurun_out/c40_bench.json
python -c "
import json; d=json.load(open('gpurun_out/c40_bench.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['stage_ms'])"
