#!/bin/bash
# ncu --set full of the camera / fusion kernels (local-context attention, fused lift-splat), summarised on the box
mkdir -p gpurun_out /tmp/ncu
timeout 400 ncu --set full --clock-control none -k regex:'local_attention|lss_splat' -c 4 -o /tmp/ncu/c43 -f python tools/profile_camera.py 2 1 focalformer3d_lc > gpurun_out/c43_ncu.log 2>&1
tail -1 gpurun_out/c43_ncu.log
python tools/ncu_summary.py /tmp/ncu/c43.ncu-rep > gpurun_out/c43_ncu.txt 2>&1
python tools/ncu_table.py gpurun_out/c43_ncu.txt 4
