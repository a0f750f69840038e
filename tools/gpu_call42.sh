#!/bin/bash
# final ncu --set full captures of the dense TMA GEMMs (new epilogue), the strided conv, the lattice GEMM and the depthwise conv
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'tmagemm_kernel<\(int\)0, \(int\)128|tmagemm_kernel<\(int\)1, \(int\)128|dwconv3x3_kernel' -c 14 -o /tmp/ncu/c42 -f python tools/profile_forward.py 1 > gpurun_out/c42_ncu.log 2>&1
tail -1 gpurun_out/c42_ncu.log
python tools/ncu_summary.py /tmp/ncu/c42.ncu-rep > gpurun_out/c42_ncu.txt 2>&1
python tools/ncu_stalls_all.py /tmp/ncu/c42.ncu-rep 14 > gpurun_out/c42_stalls.txt 2>&1
python tools/ncu_table.py gpurun_out/c42_ncu.txt 14
