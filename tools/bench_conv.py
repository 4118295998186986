"""conv3x3 (round-1 kernel) vs conv3x3_fused at the decoder's shapes: CUDA-event times, TFLOP/s of the nominal 3x3 conv."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
DEV = "cuda:0"; dt = torch.bfloat16
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
G = 32
for (B, H, W, Cin, Cout) in [(4, 1024, 1024, 128, 128), (4, 512, 512, 256, 256), (4, 256, 256, 512, 512), (4, 128, 128, 512, 512), (4, 512, 512, 512, 256)]:
    x = torch.randn((B, H, W, Cin), device=DEV).to(dt)
    w = (torch.randn((Cout, 3, 3, Cin), device=DEV) / math.sqrt(9 * Cin)).to(dt)
    b = torch.randn((Cout,), device=DEV).to(dt)
    gamma, beta = torch.ones(Cin, device=DEV).to(dt), torch.zeros(Cin, device=DEV).to(dt)
    stats = ops.groupnorm_stats(x, G, 1e-5)
    part = torch.empty((B, H * W // 128, G, 2), dtype=torch.float32, device=DEV)
    out = torch.empty((B, H, W, Cout), device=DEV, dtype=dt)
    fl = 2.0 * B * H * W * 9 * Cin * Cout
    res = {}
    res["old conv"] = timeit(lambda: ops.conv3x3(x, w, b, out=out))
    res["old apply+conv+stats"] = timeit(lambda: (ops.conv3x3(ops.groupnorm_apply(x, stats, gamma, beta, G, True), w, b, out=out), ops.groupnorm_stats(out, G, 1e-5)))
    res["fused plain"] = timeit(lambda: ops.conv3x3_fused(x, w, bias=b, out=out))
    res["fused +stats"] = timeit(lambda: ops.conv3x3_fused(x, w, bias=b, out=out, out_partial=part, out_G=G))
    res["fused norm+silu"] = timeit(lambda: ops.conv3x3_fused(x, w, bias=b, out=out, gn=(stats, gamma, beta, G), silu=True))
    res["fused norm+silu+stats"] = timeit(lambda: ops.conv3x3_fused(x, w, bias=b, out=out, gn=(stats, gamma, beta, G), silu=True, out_partial=part, out_G=G))
    print(f"B{B} {H}x{W} {Cin}->{Cout}: " + "  ".join(f"{k} {v:.3f} ms ({fl / v / 1e9:.0f} TF/s)" for k, v in res.items()), flush=True)
for (B, H, W, C) in [(4, 512, 512, 256), (4, 256, 256, 512), (4, 128, 128, 512)]:
    x = torch.randn((B, H, W, C), device=DEV).to(dt)
    w = (torch.randn((C, 3, 3, C), device=DEV) / math.sqrt(9 * C)).to(dt)
    b = torch.randn((C,), device=DEV).to(dt)
    wp = ops.conv_up_weights(w)
    out = torch.empty((B, 2 * H, 2 * W, C), device=DEV, dtype=dt)
    fl = 2.0 * B * 4 * H * W * 9 * C * C
    t_old = timeit(lambda: ops.conv3x3(ops.upsample_nearest2x(x), w, b, out=out))
    t_new = timeit(lambda: ops.conv3x3_fused(x, wp, bias=b, out=out, up=True))
    print(f"up B{B} {H}x{W}->x2 {C}: old upsample+conv {t_old:.3f} ms ({fl / t_old / 1e9:.0f} TF/s nominal)  fused {t_new:.3f} ms ({fl / t_new / 1e9:.0f} TF/s nominal)", flush=True)
