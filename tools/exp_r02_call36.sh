#!/bin/bash
# round 2, GPU call 36: the attention checks of the pytest suite on the final code (v6 issuer loop refactored in call 35)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -3 | tee gpurun_out/r02_pytest_attention_call36.txt
timeout 60 python tools/run_gpu_checks.py +experimental check_attention_v6_one 2>&1 | tail -2 | cut -c1-300
