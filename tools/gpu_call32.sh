#!/bin/bash
# final sanity of the committed tree: full GPU suite, smoke, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 400 2>&1 | tail -4 | tee gpurun_out/c32_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/c32_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --bs-sweep "" 2> gpurun_out/c32_bench.err | tail -1 > gpurun_out/c32_bench.json
python -c "
import json; d=json.load(open('gpurun_out/c32_bench.json')); print(d['value'], d['e2e']['value'], d['parity']['pass'], d['clocks'])"
