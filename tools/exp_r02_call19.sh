#!/bin/bash
# round 2, GPU call 19: attention v3 variants (MUFU token, per-warp arrivals, 3-input max) x poly share, same process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python tools/exp_attention_flags.py 2 2>&1 | tee gpurun_out/r02_att_flags.txt | cut -c1-250
