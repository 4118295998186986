"""Sustained (multi-second, power-capped) throughput of the GEMM mix of one FLUX C4 block sequence:
ours (pair / single kernels) vs torch.matmul (cuBLAS) on the same shapes with rotating weights (no L2 reuse)."""
import json
import math
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops  # noqa: E402

DEV = "cuda:0"
dt = torch.bfloat16
SHAPES = [(16384, 9216, 3072), (16384, 3072, 3072), (16384, 12288, 3072), (16384, 3072, 12288),
          (17408, 9216, 3072), (17408, 12288, 3072), (17408, 3072, 15360)]
NW = 6  # weight sets per shape


def smi():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                         capture_output=True, text=True).stdout.strip().split(",")
    return float(out[0]), float(out[1])


def run(fn, seconds):
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    t0 = time.time()
    it = 0
    clocks = []
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(4):
            fn(it)
            it += 1
        clocks.append(smi())
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    return it, s.elapsed_time(e) / 1e3, clocks


def main():
    res = {}
    bufs = []
    for (M, N, K) in SHAPES:
        A = torch.randn((M, K), device=DEV, dtype=dt)
        Ws = [torch.randn((N, K), device=DEV, dtype=dt) / math.sqrt(K) for _ in range(NW)]
        out = torch.empty((M, N), device=DEV, dtype=dt)
        bufs.append((A, Ws, out))
    flops_seq = sum(2.0 * M * N * K for (M, N, K) in SHAPES)

    def ours(i):
        for (A, Ws, out) in bufs:
            ops.gemm(A, Ws[i % NW], out=out)

    def cublas(i):
        for (A, Ws, out) in bufs:
            torch.matmul(A, Ws[i % NW].t(), out=out)

    for name, fn in [("ours", ours), ("cublas", cublas), ("ours_again", ours)]:
        it, sec, clocks = run(fn, 5.0)
        mhz = sorted(c[0] for c in clocks)[len(clocks) // 2]
        watts = sorted(c[1] for c in clocks)[len(clocks) // 2]
        res[name] = {"tflops": flops_seq * it / sec / 1e12, "sm_mhz_median": mhz, "power_w_median": watts, "iters": it}
        print(name, res[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("DK_GEMM_PAIR", "1")
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"bench_sustained_pair{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
