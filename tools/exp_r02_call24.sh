#!/bin/bash
# round 2, GPU call 24: attention v3 register-resident exponential pass (flags 16, 112 registers) x poly; trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/exp_attention_flags.py 2 2>&1 | tee gpurun_out/r02_att_flags4.txt | cut -c1-250
DK_ATT_FLAGS=16 timeout 60 python tools/trace_attention.py gpurun_out/r02_att_trace_resident.txt 2>&1 | tail -1 | cut -c1-100
