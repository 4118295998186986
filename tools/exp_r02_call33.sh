#!/bin/bash
# round 2, GPU call 33: final state: full -m gpu suite (74 tests incl. attention v6 and the narrow fused conv), C4 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_call33.txt
timeout 300 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_C4.err | tail -1 > gpurun_out/r02_bench_C4_call33.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_C4_call33.json'))
print('C4', round(d['value'],4), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'roofline', round(d['roofline']['frac'],3), d.get('last_step_ms'), d['clocks'])
" || tail -5 gpurun_out/bench_C4.err
