#!/bin/bash
# round 2, GPU call 22: attention v3 streamed exponential pass (flags 8) x 3-input max (4) x poly; trace of the streamed form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/exp_attention_flags.py 2 2>&1 | tee gpurun_out/r02_att_flags2.txt | cut -c1-250
DK_ATT_FLAGS=8 timeout 60 python tools/trace_attention.py gpurun_out/r02_att_trace_streamed.txt 2>&1 | tail -1 | cut -c1-100
