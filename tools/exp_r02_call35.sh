#!/bin/bash
# round 2, GPU call 35: attention v6, one thread per row + one MMA issuer warp per Q tile (352 threads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for c in "1 64 1 128" "2 300 2 128" "1 1280 3 128 256" "2 1178 2 64 1024" "1 4400 4 128"; do
  V6_MODE=4 timeout 40 python tools/exp_attention_v6.py check $c 2>&1 | grep -E "v6 mode|Error|error|assert|watchdog" | head -4 || echo "case $c: timeout / no output"
done
timeout 100 python tools/exp_attention_v6.py bench 2>&1 | tail -9
} | tee gpurun_out/r02_att_v6_two_issuers.txt | cut -c1-260
