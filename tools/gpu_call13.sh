#!/bin/bash
# round-2 GPU call 13: fresh launch list of one forward (current kernels)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 340 --csv --log-file gpurun_out/c13_launches.csv python tools/profile_forward.py 3 > gpurun_out/c13_ncu.log 2>&1
tail -2 gpurun_out/c13_ncu.log
