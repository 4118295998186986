#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== checks"
timeout 600 python tools/run_gpu_checks.py +experimental check_attention 2>&1 | tail -3 | cut -c1-200
echo "=== attention A/B (split x poly)"
{
for round in 1 2; do
for sp in 0 1; do for po in 0 1 2; do
DK_ATT_SPLIT=$sp DK_ATT_POLY=$po TAG="split=$sp poly=$po" timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
done; done
done
} | tee gpurun_out/r02_att_ab5.txt
