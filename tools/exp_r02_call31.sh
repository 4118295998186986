#!/bin/bash
# round 2, GPU call 31: narrow-output fused conv (conv_norm_out + SiLU + conv_out in one kernel), parallel GroupNorm
# finalize: checks + decode timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/run_gpu_checks.py check_conv_fused check_groupnorm check_vae_decode 2>&1 | tail -9 | cut -c1-330
for B in 4 1; do
  timeout 120 python tools/profile_vae.py $B 4 2>&1 | tail -1
done | tee gpurun_out/r02_vae_timing_call31.txt
