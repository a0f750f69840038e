#!/bin/bash
# round-2 GPU call 17: rulebooks on a side stream + spatially coherent mask groups + ldmatrix decoder stage
mkdir -p gpurun_out
echo "== ops + decoder stage + e2e + fullsize (defaults: overlap on, spatial sort on)"
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decoder_stage.py tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 400 2>&1 | tail -8 | tee gpurun_out/c17_tests.log
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  echo "== bench overlap=$1 spatial=$2"
  FF3D_SPARSE_OVERLAP=$1 FF3D_SPATIAL_SORT=$2 timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c17_bench_$1$2.err | tail -1 > gpurun_out/c17_bench_$1$2.json
  grep -E "ms  x" gpurun_out/c17_bench_$1$2.err | grep spconv | head -5
  python -c "
import json; d=json.load(open('gpurun_out/c17_bench_$1$2.json')); print(d['value'], d['e2e']['value'], d['stage_ms'], d.get('parity'))"
done
