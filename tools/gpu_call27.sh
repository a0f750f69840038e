#!/bin/bash
# round-2 GPU call 27: ncu --set full of the ROWS TMA GEMM (value_proj, roi_mlp) and of the final decoder stage kernel, summarised on the box
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'tmagemm_kernel<\(int\)0, \(int\)128|decoder_stage_kernel' -c 5 -o /tmp/ncu/c27 -f python tools/profile_forward.py 1 > gpurun_out/c27_ncu.log 2>&1
tail -1 gpurun_out/c27_ncu.log
python tools/ncu_summary.py /tmp/ncu/c27.ncu-rep > gpurun_out/c27_ncu.txt 2>&1
python tools/ncu_stalls_all.py /tmp/ncu/c27.ncu-rep 24 > gpurun_out/c27_stalls.txt 2>&1
python tools/ncu_table.py gpurun_out/c27_ncu.txt 5
ls -la /tmp/ncu
