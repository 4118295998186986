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


def run(fn, iters):
    """enqueue `iters` sequences back to back (no host sync inside), sample clocks/power in the background"""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    q = "clocks.sm,power.draw"
    proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                            stdout=subprocess.PIPE, text=True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    proc.terminate()
    lines = proc.stdout.read().strip().splitlines()
    clocks = []
    for l in lines[2:]:
        try:
            a, b = l.split(",")
            clocks.append((float(a), float(b)))
        except ValueError:
            pass
    return iters, s.elapsed_time(e) / 1e3, clocks or [(0.0, 0.0)]


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

    legs = [("ours", ours), ("cublas", cublas), ("ours_again", ours)]
    if os.environ.get("BS_ONLY_OURS"):
        legs = [("ours", ours)]
    for name, fn in legs:
        it, sec, clocks = run(fn, int(os.environ.get("BS_ITERS", "500")))
        mhz = sorted(c[0] for c in clocks)[len(clocks) // 2]
        watts = sorted(c[1] for c in clocks)[len(clocks) // 2]
        res[name] = {"tflops": flops_seq * it / sec / 1e12, "sm_mhz_median": mhz, "power_w_median": watts, "iters": it}
        print(name, res[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("BS_TAG") or (os.environ.get("DK_GEMM_PAIR", "1") + "_tmastore" + os.environ.get("DK_GEMM_TMA_STORE", "1"))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"bench_sustained_pair{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
