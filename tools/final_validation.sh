#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, bench (both arms), ncu launch list, camera / fusion profiles.
# Usage (from the repo root, under gpurun): bash tools/final_validation.sh
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 150 2>&1 | tail -4 | tee gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json
cut -c1-400 gpurun_out/final_bench.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/final_bench_ref.err | tail -1 > gpurun_out/final_bench_ref.json
cut -c1-400 gpurun_out/final_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 245 -c 330 --csv --log-file gpurun_out/launches_final.csv \
  python tools/profile_forward.py 3 > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log
timeout 100 python tools/profile_camera.py 1 5 deformformer3d_c_r50 2>&1 | tail -1 | cut -c1-300
timeout 120 python tools/profile_camera.py 2 4 focalformer3d_lc 2>&1 | tail -1 | cut -c1-300
