#!/bin/bash
# round 2, GPU call 2: attention v5 validation + A/B; pair-GEMM L2-policy / band-height A/B (sustained) + DRAM bytes (ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attention v5 check"
timeout 300 python tools/run_gpu_checks.py +experimental attention_v5 2>&1 | tail -4
cp gpurun_out/kernel_checks.json gpurun_out/r02_checks_att5.json
echo "=== attention A/B"
for round in 1 2; do
  for impl in 3 5; do
    DK_ATTENTION_IMPL=$impl TAG="impl=$impl" timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
  done
done | tee gpurun_out/r02_ab_attention_v5.txt
echo "=== GEMM sustained A/B"
run_cfg() {  # tag, env...
  local tag=$1; shift
  env "$@" BS_ONLY_OURS=1 BS_ITERS=250 BS_TAG=$tag timeout 200 python tools/bench_sustained.py 2>&1 | grep ours | sed "s/^/$tag /"
}
{
run_cfg base
run_cfg w1a2 DK_GEMM_HINT_W=1 DK_GEMM_HINT_A=2
run_cfg w2a1_gm4 DK_GEMM_HINT_W=2 DK_GEMM_HINT_A=1 DK_GEMM_GM=4
run_cfg gm32 DK_GEMM_GM=32
run_cfg gm32_w1a2 DK_GEMM_GM=32 DK_GEMM_HINT_W=1 DK_GEMM_HINT_A=2
run_cfg gm8 DK_GEMM_GM=8
run_cfg base_again
} | tee gpurun_out/r02_gemm_sustained_ab.txt
echo "=== GEMM DRAM bytes (ncu)"
ncu_cfg() {
  local tag=$1; shift
  for shape in "16384 12288 3072 gelu" "17408 3072 15360" "16384 3072 12288"; do
    env "$@" TAG=$tag timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
      --clock-control none -k regex:gemm2 -s 4 -c 1 --csv python tools/one_gemm.py $shape 2>/dev/null \
      | grep -E "dram__bytes|hit_rate|gpu__time" | awk -F'","' -v t="$tag" -v s="$shape" '{gsub(/"/,"",$NF); print t, "|", s, "|", $(NF-2), $(NF-1), $NF}'
  done
}
{
ncu_cfg base
ncu_cfg w1a2 DK_GEMM_HINT_W=1 DK_GEMM_HINT_A=2
ncu_cfg w2a1_gm4 DK_GEMM_HINT_W=2 DK_GEMM_HINT_A=1 DK_GEMM_GM=4
ncu_cfg gm32 DK_GEMM_GM=32
ncu_cfg gm32_w1a2 DK_GEMM_GM=32 DK_GEMM_HINT_W=1 DK_GEMM_HINT_A=2
} | tee gpurun_out/r02_gemm_dram_ab.txt
