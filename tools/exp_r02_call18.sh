#!/bin/bash
# round 2, GPU call 18 (first of the re-entry session): validate HEAD (full -m gpu suite), per-kernel `ncu --set full`
# capture at the real shapes (N2), ncu launch list of one C4 step, C4 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_call18.txt
echo "=== ncu --set full, every kernel of the C4 path at its real shape"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:(gemm|attention|conv|groupnorm|softmax_rows|ln_modulate).*_kernel' -c 80 -f -o gpurun_out/prof_r02_kernels \
  python tools/profile_kernels_r02.py > gpurun_out/ncu_kernels.log 2>&1
tail -2 gpurun_out/ncu_kernels.log
python tools/ncu_to_json.py gpurun_out/prof_r02_kernels.ncu-rep gpurun_out/r02_ncu_kernels.json 2>&1 | cut -c1-260
echo "=== launch list of one C4 step"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r02_launches_c4_step.csv python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_c4_step.csv 2>/dev/null | tee gpurun_out/r02_launches_c4_step_summary.txt | head -32
gzip -f gpurun_out/r02_launches_c4_step.csv
echo "=== bench C4"
timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_C4.err | tail -1 > gpurun_out/r02_bench_C4_call18.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_C4_call18.json'))
print('C4', d['value'], 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'roofline', round(d['roofline']['frac'],3), d.get('last_step_ms'), d['clocks'])
" || tail -5 gpurun_out/bench_C4.err
