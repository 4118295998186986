#!/bin/bash
# round 2, GPU call 25: attention v3 half-resident exponential pass (flags 16) x poly; trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/exp_attention_flags.py 2 2>&1 | tee gpurun_out/r02_att_flags5.txt | cut -c1-250
DK_ATT_FLAGS=16 timeout 60 python tools/trace_attention.py gpurun_out/r02_att_trace_halfres.txt 2>&1 | tail -1 | cut -c1-100
