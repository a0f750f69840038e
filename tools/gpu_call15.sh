#!/bin/bash
# round-2 GPU call 15: ncu --set full of the decoder stage kernel and the cp.async sparse gather-GEMM kernels (one launch each)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decoder_stage_kernel' -c 1 \
  -o gpurun_out/c15_decstage -f python tools/profile_forward.py 1 > gpurun_out/c15_ncu_a.log 2>&1
tail -2 gpurun_out/c15_ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'tmagemm_kernel<2, (64|32|128)' -s 2 -c 9 -o gpurun_out/c15_sparse -f python tools/profile_forward.py 1 > gpurun_out/c15_ncu_b.log 2>&1
tail -2 gpurun_out/c15_ncu_b.log
ls -la gpurun_out/c15*
