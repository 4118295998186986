#!/bin/bash
# round 2, GPU call 28: validation of the round-end state: full -m gpu suite, bench.py for every BASELINE config,
# ncu --set full of the kernels whose default changed (attention, both head dims), launch list of one C4 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_call28.txt
echo "=== bench"
for wl in C4 C2 C3 C5; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 2>gpurun_out/bench_$wl.err | tail -1 > gpurun_out/r02_bench_${wl}_call28.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02_bench_${wl}_call28.json'))
print('$wl', round(d['value'],4), d['unit'], 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'roofline', round(d['roofline']['frac'],3), 'mmdit_frac_sust', round(d.get('mmdit_tensor_frac_sustained',0),3), 'split', d.get('last_step_ms'), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
" || tail -3 gpurun_out/bench_$wl.err
done
echo "=== ncu attention"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:attention.*_kernel' -c 4 -f -o gpurun_out/prof_r02_attention_final \
  python tools/profile_kernels_r02.py > gpurun_out/ncu_att.log 2>&1
python tools/ncu_to_json.py gpurun_out/prof_r02_attention_final.ncu-rep gpurun_out/r02_ncu_attention_final.json 2>&1 | cut -c1-200
echo "=== launch list of one C4 step"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r02_launches_c4_step_final.csv python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_c4_step_final.csv 2>/dev/null | tee gpurun_out/r02_launches_c4_step_final_summary.txt | head -12
gzip -f gpurun_out/r02_launches_c4_step_final.csv
