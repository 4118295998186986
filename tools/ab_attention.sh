#!/bin/bash
# Same-box A/B of the attention kernel variants (one process each, alternating, two rounds): default v3, 3b (pairwise
# barriers + 4 max chains), 3r (register-resident scores, 112 registers), 4 (one Q tile, double-buffered S, alternating softmax sets), 2 (legacy).  Run under gpurun:
#   gpurun --timeout 300 -- 'python tools/run_gpu_checks.py +experimental attention_v3 attention_v4 | tail -3; bash tools/ab_attention.sh'
cd "$(dirname "$0")/.."
for round in 1 2; do
  for impl in 3 3b 3r 4 2; do
    DK_ATTENTION_IMPL=$impl TAG="impl=$impl" timeout 60 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
  done
done
