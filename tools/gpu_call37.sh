#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder_stage.py -q -m gpu --timeout 200 2>&1 | tail -12 | tee gpurun_out/c37_decstage.log
