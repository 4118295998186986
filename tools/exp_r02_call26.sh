#!/bin/bash
# round 2, GPU call 26: attention v3 streamed pass + spinning issuer (1) / per-warp arrivals (2); traces of each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python tools/exp_attention_flags.py 2 2>&1 | tee gpurun_out/r02_att_flags6.txt | cut -c1-250
for fl in 9 10 11; do
  DK_ATT_FLAGS=$fl timeout 60 python tools/trace_attention.py gpurun_out/r02_att_trace_f$fl.txt 2>&1 | tail -1 | cut -c1-60
done
