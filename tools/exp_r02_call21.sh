#!/bin/bash
# round 2, GPU call 21: SM-clock timestamp trace of one attention CTA (hand-over latencies of the softmax / MMA chain)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python tools/trace_attention.py gpurun_out/r02_att_trace.txt 2>&1 | tail -3 | cut -c1-300
