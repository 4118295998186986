import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops
DEV = "cuda:0"
def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters / 1e3
if __name__ == "__main__":
  for name, (B, S, heads, d) in {"c4": (4, 4352, 24, 128), "c2": (1, 1280, 24, 128), "sd3": (8, 4685, 24, 64)}.items():
    dt = torch.bfloat16 if d == 128 else torch.float16
    qkv = torch.randn((B * S, 3 * heads * d), device=DEV, dtype=dt)
    o = torch.empty((B * S, heads * d), device=DEV, dtype=dt)
    t = timeit(lambda: ops.attention(qkv, B, S, heads, d, o))
    print(f"{os.environ.get('TAG','')} {name}: {t*1e3:.3f} ms {4.0*B*heads*S*S*d/t/1e12:.0f} TF/s", flush=True)
