"""Run one GEMM shape a few times (for ncu metric passes and env-knob A/Bs): python tools/one_gemm.py M N K [gelu]"""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
from diffusionkit_b200._lib import ACT_GELU_ERF, ACT_NONE
DEV = "cuda:0"; dt = torch.bfloat16
M, N, K = (int(v) for v in sys.argv[1:4])
act = ACT_GELU_ERF if len(sys.argv) > 4 and sys.argv[4] == "gelu" else ACT_NONE
A = torch.randn((M, K), device=DEV, dtype=dt); W = torch.randn((N, K), device=DEV, dtype=dt) / math.sqrt(K)
b = torch.randn((N,), device=DEV, dtype=dt); out = torch.empty((M, N), device=DEV, dtype=dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ts = []
for i in range(6):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.gemm(A, W, out=out, bias=b, act=act); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print(os.environ.get("TAG", ""), f"{M}x{N}x{K}", "ms", round(min(ts[2:]), 4), "TF/s", round(2.0 * M * N * K / min(ts[2:]) / 1e9, 1), flush=True)
