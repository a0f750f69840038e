#!/bin/bash
# Round-2 evidence, part B (one B200): ncu --set full of every kernel family (summarised on the box: the reports are too
# large to travel back) and compute-sanitizer runs.   Usage: bash tools/gpu_evidence_b.sh [tag] [ncu|san|all]
T=${1:-ev}; W=${2:-all}
mkdir -p gpurun_out /tmp/ncu
if [ "$W" = "all" ] || [ "$W" = "ncu" ]; then
echo "== ncu set full: GEMM family + decoder stage"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'tmagemm_kernel<\(int\)2, \(int\)(16|32|64|128)|tmagemm_kernel<\(int\)1, \(int\)128|tmagemm_kernel<\(int\)0, \(int\)128|decoder_stage_kernel|tcgemm_kernel<\(int\)1' \
  -c 40 -o /tmp/ncu/${T}_gemm -f python tools/profile_forward.py 1 > gpurun_out/${T}_ncu_gemm.log 2>&1
tail -1 gpurun_out/${T}_ncu_gemm.log
python tools/ncu_summary.py /tmp/ncu/${T}_gemm.ncu-rep > gpurun_out/${T}_ncu_gemm.txt 2>&1
python tools/ncu_stalls_all.py /tmp/ncu/${T}_gemm.ncu-rep 16 > gpurun_out/${T}_ncu_gemm_stalls.txt 2>&1
echo "== ncu set full: non-GEMM kernels"
timeout 900 ncu --set full --clock-control none \
  -k regex:'hip_select|hip_nms|hip_heat|roi_sample|vox_assign|vox_gather|vox_hash|sp_sites_insert|sp_sites_assign|sp_tap_keys|sp_nbr_permute|sp_level_permute|sp_hash_build|rs_hist|rs_scatter|dwconv3x3|add_bcast_rows_split|split_rows|sine_embed|head_update|box_decode' \
  -c 60 -o /tmp/ncu/${T}_nongemm -f python tools/profile_forward.py 1 > gpurun_out/${T}_ncu_nongemm.log 2>&1
tail -1 gpurun_out/${T}_ncu_nongemm.log
python tools/ncu_summary.py /tmp/ncu/${T}_nongemm.ncu-rep > gpurun_out/${T}_ncu_nongemm.txt 2>&1
ls -la /tmp/ncu
fi
if [ "$W" = "all" ] || [ "$W" = "san" ]; then
echo "== sanitizer memcheck (fused decoder stage on)"
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/${T}_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_memcheck.log; tail -4 gpurun_out/${T}_memcheck.log
echo "== sanitizer racecheck (fused decoder stage on)"
timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 400 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/${T}_racecheck.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_racecheck.log; tail -4 gpurun_out/${T}_racecheck.log
grep -E "hazard detected|Race reported" gpurun_out/${T}_racecheck.log | sed 's/.*in //' | sort | uniq -c | sort -rn | head -30
echo "== sanitizer synccheck"
timeout 420 compute-sanitizer --tool synccheck --print-limit 20 python tools/tiny_forward.py focalformer3d_l 1 > gpurun_out/${T}_synccheck.log 2>&1
echo "rc=$?" >> gpurun_out/${T}_synccheck.log; tail -4 gpurun_out/${T}_synccheck.log
fi
ls -la gpurun_out/${T}_* | head -30
