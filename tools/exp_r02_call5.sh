#!/bin/bash
# round 2, GPU call 5: attention v5 (mask-free hot path, cheap polling) A/B; fused VAE conv validation (both descriptor encodings)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== v5 check"
timeout 300 python tools/run_gpu_checks.py +experimental attention_v5 2>&1 | tail -2
echo "=== attention A/B"
{
DK_ATTENTION_IMPL=3 TAG="v3       " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
for dbg in 0 4 2; do
  DK_ATTENTION_IMPL=5 DK_ATT_DEBUG=$dbg TAG="v5 dbg=$dbg" timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
done
DK_ATTENTION_IMPL=3 TAG="v3 again " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
} | tee gpurun_out/r02_att5_ab2.txt
echo "=== conv_fused"
timeout 600 python tools/run_gpu_checks.py +experimental check_conv_fused 2>&1 | tail -4
cp gpurun_out/kernel_checks.json gpurun_out/r02_checks_conv_fused.json
