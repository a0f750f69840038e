#!/bin/bash
# 2 GPUs: bench under torchrun -- weak scaling + multi-GPU output parity (every rank's top-k re-computed on rank 0)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --bs-sweep "" --no-cpu-baseline 2> gpurun_out/c41_bench_2gpu.err | tail -1 > gpurun_out/c41_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/c41_bench_2gpu.json')); print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e']['value'], d.get('parity'))"
