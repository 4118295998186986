#!/bin/bash
# round 2, GPU call 27: VAE decode with GroupNorm-apply+SiLU inside the fused convolutions vs as a separate HBM pass in
# front of them; attention default (streamed + poly 1) vs the previous default, longer timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for mode in 0 1 0 1; do
  DK_VAE_NORM_IN_CONV=$mode timeout 200 python tools/profile_vae.py 4 4 2>&1 | tail -1 | sed "s/^/norm_in_conv=$mode /"
done
DK_VAE_NORM_IN_CONV=0 timeout 200 python tools/profile_vae.py 1 4 2>&1 | tail -1 | sed "s/^/norm_in_conv=0 /"
DK_VAE_NORM_IN_CONV=1 timeout 200 python tools/profile_vae.py 1 4 2>&1 | tail -1 | sed "s/^/norm_in_conv=1 /"
} | tee gpurun_out/r02_vae_norm_mode.txt
timeout 300 python tools/run_gpu_checks.py check_vae_decode_512 check_attention_v3 2>&1 | tail -6 | cut -c1-400
timeout 120 python tools/exp_attention_variants.py 2 2>&1 | tee gpurun_out/r02_att_variants_final.txt | cut -c1-200
