#!/bin/bash
# round-2 GPU call 4: new bench.py (CUDA-graph replay, pipelined e2e, parity key, bs sweep, gpu stand-in) + smoke + ncu of the f16 kernels
mkdir -p gpurun_out
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/c4_smoke.log
echo "== ops tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 200 2>&1 | tail -4 | tee gpurun_out/c4_ops.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/c4_bench.err | tail -1 > gpurun_out/c4_bench.json
tail -30 gpurun_out/c4_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/c4_bench.json'))
for k in ('value','ms_per_step','e2e','eager_ms_per_step','launches_per_step','stage_ms','parity','bs_sweep','gpu_standin','cpu_baseline','clocks'):
    print(k, d.get(k))
print('roof', {k:v for k,v in d['roofline'].items() if k in ('kernel','achieved','frac','executed','sparse_executed_vs_algorithmic','ms_per_launch')})
P
echo "== ncu f16 sparse128 + conv128"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tcgemm_kernel<2, 128, 2, 2, true>|tcgemm_kernel<1, 128, 3, 1, true>|tcgemm_kernel<2, 64, 2, 1, true>' \
  -s 12 -c 6 -o gpurun_out/c4_gemm -f python tools/profile_forward.py 2 > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/c4_ncu.log
ls -la gpurun_out | tail -8
