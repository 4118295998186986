"""One attention shape a few times (for ncu): python tools/one_attention.py [B S heads d]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
B, S, heads, d = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (4, 4352, 24, 128)
dt = torch.bfloat16 if d == 128 else torch.float16
qkv = torch.randn((B * S, 3 * heads * d), device="cuda:0", dtype=dt)
o = torch.empty((B * S, heads * d), device="cuda:0", dtype=dt)
for _ in range(4):
    ops.attention(qkv, B, S, heads, d, o)
torch.cuda.synchronize()
print("done")
