"""Timestamp trace of one attention CTA (diagnostic instantiation, DK_ATT_TRACE): python tools/trace_attention.py out.txt"""
import os, sys
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/att_trace.txt"
os.environ["DK_ATT_TRACE"] = out
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
B, S, heads, d = 4, 4352, 24, 128
qkv = torch.randn((B * S, 3 * heads * d), device="cuda:0", dtype=torch.bfloat16)
o = torch.empty((B * S, heads * d), device="cuda:0", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, B, S, heads, d, o)
torch.cuda.synchronize()
print(open(out).read()[:400])
