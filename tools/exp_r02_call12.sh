#!/bin/bash
# round 2, GPU call 12: warp-role re-mapping (issuer / producer on the highest warp ids): GEMM + conv checks, GEMM sustained A/B,
# conv bench, VAE timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== checks"
timeout 900 python tools/run_gpu_checks.py check_gemm check_conv check_fullsize_gemm check_fullsize_conv 2>&1 | tail -4 | cut -c1-300
echo "=== GEMM sustained A/B (roles)"
{
for r in 1 2; do
BS_ONLY_OURS=1 BS_ITERS=250 BS_TAG=roles_new timeout 200 python tools/bench_sustained.py 2>&1 | grep ours | sed "s/^/new /"
DK_GEMM_ROLES=0 BS_ONLY_OURS=1 BS_ITERS=250 BS_TAG=roles_old timeout 200 python tools/bench_sustained.py 2>&1 | grep ours | sed "s/^/old /"
done
} | tee gpurun_out/r02_gemm_roles_ab.txt
echo "=== one GEMM isolated"
for r in 1 0; do DK_GEMM_ROLES=$r TAG="roles=$r" python tools/one_gemm.py 16384 12288 3072 gelu; DK_GEMM_ROLES=$r TAG="roles=$r" python tools/one_gemm.py 17408 3072 15360; done
echo "=== conv bench"
python tools/bench_conv.py 2>&1 | tee gpurun_out/r02_bench_conv_v3.txt
echo "=== VAE timing"
for fused in 1 0; do
  DK_VAE_FUSED=$fused timeout 300 python tools/profile_vae.py 4 3 2>&1 | tail -1 | sed "s/^/fused=$fused /"
  DK_VAE_FUSED=$fused timeout 300 python tools/profile_vae.py 1 3 2>&1 | tail -1 | sed "s/^/fused=$fused /"
done | tee gpurun_out/r02_vae_timing_v3.txt
