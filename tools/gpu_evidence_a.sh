#!/bin/bash
# Round-2 evidence, part A (one B200): full GPU suite, smoke, bench (both arms, all configs), launch list.
# Usage (repo root, under gpurun): bash tools/gpu_evidence_a.sh [tag]
T=${1:-ev}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/${T}_smi.txt
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 2>&1 | tail -6 | tee gpurun_out/${T}_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
echo "== bench (default flags)"
timeout 900 python bench.py 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json
python - <<P
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
for k in ('value','ms_per_step','e2e','eager_ms_per_step','launches_per_step','stage_ms','parity','bs_sweep','gpu_standin','cpu_baseline','clocks','overflow_flags'):
    print(k, str(d.get(k))[:600])
print('roof', {k:v for k,v in d['roofline'].items() if k in ('kernel','achieved','frac','ms_per_launch','share_of_step','sparse_executed_vs_algorithmic')})
P
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_bench_ref.err | tail -1 > gpurun_out/${T}_bench_ref.json
cut -c1-500 gpurun_out/${T}_bench_ref.json
for c in waymo_l lc c_r50; do
  echo "== bench --config $c"
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 2> gpurun_out/${T}_bench_$c.err | tail -1 > gpurun_out/${T}_bench_$c.json
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$c.json')); print(d.get('value'), d.get('e2e'), d.get('stage_ms'), d.get('parity'), str(d.get('cpu_baseline'))[:200])"
done
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 210 -c 230 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_forward.py 3 > gpurun_out/${T}_ncu_list.log 2>&1
tail -1 gpurun_out/${T}_ncu_list.log
