#!/bin/bash
# round 2, GPU call 32: attention v6 (64-key steps, double-buffered score accumulators): parity cases one per process
# (a trapped kernel kills only its own context), then same-process timing against the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for c in "1 128 1 128" "1 64 1 128" "2 300 2 128" "1 1280 3 128 256" "2 1178 2 64 1024" "1 333 2 64" "1 4400 4 128"; do
  timeout 45 python tools/exp_attention_v6.py check $c 2>&1 | grep -E "v6 B|Error|error|assert|watchdog" | head -4 || echo "case $c: timeout / no output"
done
timeout 90 python tools/exp_attention_v6.py bench 2>&1 | tail -9
} | tee gpurun_out/r02_att_v6.txt | cut -c1-220
