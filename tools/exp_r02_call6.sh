#!/bin/bash
# round 2, GPU call 6: attention (v3 with the mask-free hot path vs v5), fused VAE decoder: model checks, timing, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attention checks + A/B"
timeout 600 python tools/run_gpu_checks.py +experimental check_attention_d check_attention_large check_edge check_attention_v5 check_fullsize_attention 2>&1 | tail -8 | cut -c1-300
{
for round in 1 2; do
DK_ATTENTION_IMPL=3 TAG="v3m      " timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
for dbg in 0 4; do
  DK_ATTENTION_IMPL=5 DK_ATT_DEBUG=$dbg TAG="v5 dbg=$dbg" timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
done
done
} | tee gpurun_out/r02_att_ab3.txt
echo "=== VAE checks (fused path)"
timeout 900 python tools/run_gpu_checks.py check_conv_fused check_fullsize_conv_fused check_vae check_full_size_vae check_product_vs_reference check_pipeline_flux_tiny check_pipeline_img2img 2>&1 | tail -14 | cut -c1-400
cp gpurun_out/kernel_checks.json gpurun_out/r02_checks_vae_fused.json
echo "=== VAE timing (fused / unfused, batch 4 and 1)"
for fused in 1 0; do
  DK_VAE_FUSED=$fused timeout 300 python tools/profile_vae.py 4 4 2>&1 | sed "s/^/fused=$fused /"
  DK_VAE_FUSED=$fused timeout 300 python tools/profile_vae.py 1 4 2>&1 | sed "s/^/fused=$fused /"
done | tee gpurun_out/r02_vae_timing.txt
echo "=== VAE launch list (fused, eager)"
DK_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_vae_fused.csv python tools/profile_vae.py 4 2 > gpurun_out/ncu_vae.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_vae_fused.csv 2>/dev/null | head -30
