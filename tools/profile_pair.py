import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
DEV = "cuda:0"; dt = torch.bfloat16
M, N, K = 16384, 9216, 3072
A = torch.randn((M, K), device=DEV, dtype=dt); W = torch.randn((N, K), device=DEV, dtype=dt) / math.sqrt(K)
b = torch.randn((N,), device=DEV, dtype=dt); out = torch.empty((M, N), device=DEV, dtype=dt)
for _ in range(3):
    ops.gemm(A, W, out=out, bias=b)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.gemm(A, W, out=out, bias=b); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print("ms", min(ts), "TF/s", 2.0 * M * N * K / min(ts) / 1e9)
