"""attention v6 (64-key steps, double-buffered scores) against the fp32 reference and against the default kernel:
  python tools/exp_attention_v6.py check <B> <S> <heads> <d> [split]     one parity case (own process: a trap kills only it)
  python tools/exp_attention_v6.py bench                                  same-process timing, v3 default vs v6"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from diffusionkit_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
DEV = "cuda:0"
MODE = int(os.environ.get("V6_MODE", "2"))   # dk_attention_tuning stream value: 2 = two threads per row, 3 = one


def timeit(fn, iters=100, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters / 1e3


if sys.argv[1] == "check":
    import kernel_checks as kc

    B, S, heads, d = (int(v) for v in sys.argv[2:6])
    split = int(sys.argv[6]) if len(sys.argv) > 6 else None
    kc._setup()
    dt = torch.bfloat16 if d == 128 else torch.float16
    for poly in (0, 1):
        lib.dk_attention_tuning(1, poly, MODE)
        err = kc._attention_case(B, S, heads, d, dt, split=split, name=f"att6_B{B}_S{S}_h{heads}_d{d}_poly{poly}")
        torch.cuda.synchronize()
        print(f"v6 mode{MODE} B{B} S{S} heads{heads} d{d} split{split} poly{poly}: rel_l2 {err:.2e}", flush=True)
else:
    shapes = {"c4": (4, 4352, 24, 128), "c2": (1, 1280, 24, 128), "sd3": (8, 4685, 24, 64), "c5": (1, 4608, 24, 128)}
    for rnd in range(2):
        for name, (B, S, heads, d) in shapes.items():
            dt = torch.bfloat16 if d == 128 else torch.float16
            qkv = torch.randn((B * S, 3 * heads * d), device=DEV, dtype=dt)
            o = torch.empty((B * S, heads * d), device=DEV, dtype=dt)
            line = [f"r{rnd} {name}"]
            for tag, (sp, po, st) in {"v3": (-1, -1, -1), "v6 p0": (1, 0, 2), "v6 p1": (1, 1, 2), "v6-1T p0": (1, 0, 3),
                                      "v6-1T p1": (1, 1, 3), "v6-2I p0": (1, 0, 4), "v6-2I p1": (1, 1, 4)}.items():
                lib.dk_attention_tuning(sp, po, st)
                t = timeit(lambda: ops.attention(qkv, B, S, heads, d, o))
                line.append(f"{tag} {4.0 * B * heads * S * S * d / t / 1e12:6.0f} TF/s")
            print(" | ".join(line), flush=True)
    lib.dk_attention_tuning(-1, -1, -1)
