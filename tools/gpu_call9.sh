#!/bin/bash
# round-2 GPU call 9: ROI / value_proj / FocalEncoder / heat-map convs on the TMA kernel + launch-time list of one forward
mkdir -p gpurun_out
echo "== e2e + fullsize L"
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py tests/test_gpu_camera.py -q -m gpu --timeout 300 -s -k "not fusion_lc_full and not waymo_l_full" -x 2>&1 | grep -E "parity:|passed|failed|Error|error|assert" | cut -c1-900 | tee gpurun_out/c9_e2e.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c9_bench.err | tail -1 > gpurun_out/c9_bench.json
grep -E "ms  x" gpurun_out/c9_bench.err | head -30
python -c "
import json; d=json.load(open('gpurun_out/c9_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 330 --csv --log-file gpurun_out/c9_launches.csv python tools/profile_forward.py 3 > gpurun_out/c9_ncu.log 2>&1
tail -2 gpurun_out/c9_ncu.log
