"""A few launches of the fused conv at one shape (for ncu): python tools/one_conv_fused.py B H W Cin Cout [norm]"""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
B, H, W, Cin, Cout = (int(v) for v in sys.argv[1:6])
norm = len(sys.argv) > 6 and sys.argv[6] == "norm"
DEV = "cuda:0"; dt = torch.bfloat16
x = torch.randn((B, H, W, Cin), device=DEV).to(dt)
w = (torch.randn((Cout, 3, 3, Cin), device=DEV) / math.sqrt(9 * Cin)).to(dt)
b = torch.randn((Cout,), device=DEV).to(dt)
gamma, beta = torch.ones(Cin, device=DEV).to(dt), torch.zeros(Cin, device=DEV).to(dt)
stats = ops.groupnorm_stats(x, 32, 1e-5)
part = torch.empty((B, H * W // 128, 32, 2), dtype=torch.float32, device=DEV)
out = torch.empty((B, H, W, Cout), device=DEV, dtype=dt)
for _ in range(4):
    ops.conv3x3_fused(x, w, bias=b, out=out, gn=(stats, gamma, beta, 32) if norm else None, silu=norm, out_partial=part, out_G=32)
torch.cuda.synchronize()
print("done")
