#!/bin/bash
# round 2, GPU call 29 (2 GPUs): the torchrun path of bench.py as the driver launches it (weights broadcast over NCCL,
# batch sharding, max-over-ranks timing)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/r02_bench_C4_n2_call29.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_C4_n2_call29.json'))
print('C4 n=2', round(d['value'],3), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'per-rank', d.get('per_rank_step_ms'), 'weights', d.get('weights_timing'))
" || tail -5 gpurun_out/bench_n2.err
