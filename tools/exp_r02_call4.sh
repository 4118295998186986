#!/bin/bash
# round 2, GPU call 4: attention v5 ablations (item order, exchange, wait/arrive style) + ncu source-level captures of v5 and v3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== v5 check (L2-friendly order)"
timeout 300 python tools/run_gpu_checks.py +experimental attention_v5 2>&1 | tail -2
echo "=== ablations"
{
DK_ATTENTION_IMPL=3 TAG="v3       " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
for dbg in 0 2 4 6 1; do
  DK_ATTENTION_IMPL=5 DK_ATT_DEBUG=$dbg TAG="v5 dbg=$dbg" timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
done
DK_ATTENTION_IMPL=3 TAG="v3 again " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
} | tee gpurun_out/r02_att5_ablation.txt
echo "=== ncu"
DK_ATTENTION_IMPL=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 2 -c 1 -f -o gpurun_out/prof_att5 python tools/one_attention.py > gpurun_out/ncu_att5.log 2>&1
DK_ATTENTION_IMPL=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 2 -c 1 -f -o gpurun_out/prof_att3 python tools/one_attention.py > gpurun_out/ncu_att3.log 2>&1
ls -la gpurun_out/*.ncu-rep
