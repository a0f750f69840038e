#!/bin/bash
# round-2 GPU call 6: gather4 producer-warp count experiment + NMS / merge tests
mkdir -p gpurun_out
echo "== NMS / merge / sparse TMA tests"
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -k "nms or iou or merge or tma_sparse" 2>&1 | tail -15 | tee gpurun_out/c6_ops.log
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -m gpu --timeout 120 -k "with_nms" 2>&1 | tail -8 | tee gpurun_out/c6_e2e_nms.log
for npw in 1 4 8; do
  echo "== bench FF3D_TMA_NPW=$npw"
  timeout 600 env FF3D_TMA_NPW=$npw python bench.py --steps 6 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c6_bench_npw$npw.err | tail -1 > gpurun_out/c6_bench_npw$npw.json
  grep -E "spconv|conv3x3s1\[128->128" gpurun_out/c6_bench_npw$npw.err
  python -c "
import json; d=json.load(open('gpurun_out/c6_bench_npw$npw.json')); print(d['value'], d['stage_ms']['sparse_encoder'])"
done
