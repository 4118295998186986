#!/bin/bash
# round 2, GPU call 15: full -m gpu suite, conv bench (role layout v6), VAE timing, bench.py for every BASELINE config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu_call15.txt
echo "=== conv bench"
python tools/bench_conv.py 2>&1 | tee gpurun_out/r02_bench_conv_v6.txt | cut -c1-400
echo "=== VAE timing"
for fused in 1 0; do
  DK_VAE_FUSED=$fused timeout 300 python tools/profile_vae.py 4 3 2>&1 | tail -1 | sed "s/^/fused=$fused /"
  DK_VAE_FUSED=$fused timeout 300 python tools/profile_vae.py 1 3 2>&1 | tail -1 | sed "s/^/fused=$fused /"
done | tee gpurun_out/r02_vae_timing_v6.txt
echo "=== bench"
for wl in C4 C2 C3 C5; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 2>gpurun_out/bench_$wl.err | tail -1 > gpurun_out/r02_bench_$wl.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02_bench_$wl.json'))
print('$wl', d['value'], d['unit'], 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'roofline', round(d['roofline']['frac'],3), 'split', d.get('last_step_ms'), 'clk', d['clocks'].get('sm_mhz'))
" || tail -3 gpurun_out/bench_$wl.err
done
