#!/bin/bash
# round 2, GPU call 3: v5 (96-register layout) validation + A/B, the whole -m gpu suite, bench C4 with both attention kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== attention v5 check"
timeout 300 python tools/run_gpu_checks.py +experimental attention_v5 2>&1 | tail -3
echo "=== attention A/B"
for round in 1 2; do
  for impl in 3 5; do
    DK_ATTENTION_IMPL=$impl TAG="impl=$impl" timeout 120 python tools/bench_attention.py 2>&1 | grep -E "c4|c2|sd3"
  done
done | tee gpurun_out/r02_ab_attention_v5b.txt
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_call3.txt
echo "=== bench C4 (v3 / v5)"
DK_ATTENTION_IMPL=3 timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_c4_v3.err | tail -1 | tee gpurun_out/r02_bench_c4_att3.json | cut -c1-400
DK_ATTENTION_IMPL=5 timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_c4_v5.err | tail -1 | tee gpurun_out/r02_bench_c4_att5.json | cut -c1-400
