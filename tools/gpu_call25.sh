#!/bin/bash
# round-2 GPU call 25: ncu --set full of every kernel family, summarised on the box; racecheck with the TMA gather4 engine
bash tools/gpu_evidence_b.sh ev2 ncu
echo "== racecheck, FF3D_SPARSE_GATHER=tma (no cp.async ring)"
FF3D_SPARSE_GATHER=tma timeout 700 compute-sanitizer --tool racecheck --print-limit 100 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/ev2_racecheck_tma.log 2>&1
echo "rc=$?" >> gpurun_out/ev2_racecheck_tma.log; tail -3 gpurun_out/ev2_racecheck_tma.log
grep -oE "at ff3d::[a-z_0-9A-Z]+" gpurun_out/ev2_racecheck_tma.log | sort | uniq -c | sort -rn | head -10
