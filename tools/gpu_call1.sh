#!/bin/bash
# round-2 GPU call 1: baseline suite + full-size parity + experimental f16 harness + ncu evidence for the non-GEMM kernels + sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/c1_smi.txt
echo "== pytest gpu (tcgen05 tf32 path)"; 
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -s 2>&1 | tail -40 | tee gpurun_out/c1_tests.log
echo "== pytest experimental f16 (ops)"
timeout 400 env FF3D_EXPERIMENTAL_F16=1 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 2>&1 | tail -30 | tee gpurun_out/c1_f16_ops.log
echo "== pytest experimental f16 (e2e)"
timeout 600 env FF3D_EXPERIMENTAL_F16=1 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 2>&1 | tail -30 | tee gpurun_out/c1_f16_e2e.log
echo "== ncu set full: non-GEMM kernels"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'hip_select|hip_nms|hip_heat|msda|mha_core|roi_sample|vox_assign|vox_gather|vox_hash|sp_down_sites|sp_subm_map|sp_down_map|sp_hash_build|dwconv3x3' \
  -s 45 -c 48 -o gpurun_out/c1_nongemm -f python tools/profile_forward.py 2 > gpurun_out/c1_ncu_nongemm.log 2>&1
tail -2 gpurun_out/c1_ncu_nongemm.log
echo "== sanitizer memcheck"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/c1_memcheck.log 2>&1
tail -5 gpurun_out/c1_memcheck.log
echo "== sanitizer racecheck"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/c1_racecheck.log 2>&1
tail -5 gpurun_out/c1_racecheck.log
ls -la gpurun_out
