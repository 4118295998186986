#!/bin/bash
# round 2, GPU call 20: attention tensor-side decomposition (no-split path; DK_ATT_DEBUG 4 = MMA free-running,
# 5 = QK^T MMAs only, 6 = PV MMAs only; nominal TFLOP/s of the full problem)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dbg in 0 4 5 6 1 2; do
  DK_ATT_SPLIT=0 DK_ATT_DEBUG=$dbg TAG="nosplit dbg=$dbg" timeout 100 python tools/bench_attention.py 2>&1 | grep -E "c4|sd3"
done | tee gpurun_out/r02_att_decomp.txt
