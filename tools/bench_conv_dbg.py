import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_attention import timeit
DEV = "cuda:0"; dt = torch.bfloat16
for (B, H, W, Cin, Cout) in [(4, 1024, 1024, 128, 128), (4, 256, 256, 512, 512)]:
    x = torch.randn((B, H, W, Cin), device=DEV).to(dt)
    w = (torch.randn((Cout, 3, 3, Cin), device=DEV) / math.sqrt(9 * Cin)).to(dt)
    b = torch.randn((Cout,), device=DEV).to(dt)
    gamma, beta = torch.ones(Cin, device=DEV).to(dt), torch.zeros(Cin, device=DEV).to(dt)
    stats = ops.groupnorm_stats(x, 32, 1e-5)
    out = torch.empty((B, H, W, Cout), device=DEV, dtype=dt)
    fl = 2.0 * B * H * W * 9 * Cin * Cout
    t0 = timeit(lambda: ops.conv3x3_fused(x, w, bias=b, out=out)) * 1e3
    t1 = timeit(lambda: ops.conv3x3_fused(x, w, bias=b, out=out, gn=(stats, gamma, beta, 32), silu=True)) * 1e3
    print(os.environ.get("DK_CF_DEBUG", "0"), f"{H}x{W} {Cin}->{Cout}: plain {t0:.3f} ms  norm+silu {t1:.3f} ms", flush=True)
