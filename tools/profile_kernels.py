"""A few launches of the dominant kernels at C4 shapes, for `ncu --set full` captures (one GPU, short)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops  # noqa: E402
from diffusionkit_b200._lib import ACT_GELU_ERF  # noqa: E402

DEV = "cuda:0"
dt = torch.bfloat16
M, N, K = 16384, 12288, 3072
A = torch.randn((M, K), device=DEV, dtype=dt)
W = torch.randn((N, K), device=DEV, dtype=dt) / math.sqrt(K)
b = torch.randn((N,), device=DEV, dtype=dt)
out = torch.empty((M, N), device=DEV, dtype=dt)
for _ in range(2):
    ops.gemm(A, W, out=out, bias=b, act=ACT_GELU_ERF)
B, S, H, d = 4, 4352, 24, 128
qkv = torch.randn((B * S, 3 * H * d), device=DEV, dtype=dt)
o = torch.empty((B * S, H * d), device=DEV, dtype=dt)
for _ in range(2):
    ops.attention(qkv, B, S, H, d, o)
x = torch.randn((1, 512, 512, 256), device=DEV, dtype=dt)
w = torch.randn((256, 3, 3, 256), device=DEV, dtype=dt) * 0.02
bb = torch.zeros((256,), device=DEV, dtype=dt)
for _ in range(2):
    ops.conv3x3(x, w, bias=bb)
xx = torch.randn((17408, 3072), device=DEV, dtype=dt)
mod = torch.randn((4, 6 * 3072), device=DEV, dtype=dt)
for _ in range(2):
    ops.ln_modulate(xx, mod[:, :3072], mod[:, 3072:6144], 4352)
torch.cuda.synchronize()
print("done")
