#!/bin/bash
# round 2, GPU call 16: attention variants on d=128 (poly exp2 share, split P publication), conv role layout re-check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== v3s check"
timeout 300 python tools/run_gpu_checks.py +experimental attention_v3s 2>&1 | tail -2 | cut -c1-300
echo "=== attention A/B"
{
for round in 1 2; do
DK_ATTENTION_IMPL=3 TAG="v3 poly0 " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
DK_ATTENTION_IMPL=3 DK_ATT_POLY=1 TAG="v3 poly1 " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
DK_ATTENTION_IMPL=3 DK_ATT_POLY=2 TAG="v3 poly2 " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
DK_ATTENTION_IMPL=3s TAG="v3s      " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
DK_ATTENTION_IMPL=5 TAG="v5       " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
done
} | tee gpurun_out/r02_att_ab4.txt
echo "=== conv bench (transform on warps 0-3,12,13 again)"
python tools/bench_conv.py 2>&1 | head -3 | tee gpurun_out/r02_bench_conv_v7.txt | cut -c1-400
DK_VAE_FUSED=1 timeout 300 python tools/profile_vae.py 4 3 2>&1 | tail -1
