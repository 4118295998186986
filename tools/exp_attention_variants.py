"""Same-process A/B of the K3 (attention v3) tuning variants through dk_attention_tuning: every (poly, stream) pair is
checked against the default kernel's output on the same inputs (bit-exact for equal poly; rel-L2 across poly) and timed.
  python tools/exp_attention_variants.py [rounds]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
lib = _lib.load()
SHAPES = {"c4": (4, 4352, 24, 128), "c4tail": (1, 4400, 24, 128), "sd3": (8, 4685, 24, 64), "c2": (1, 1280, 24, 128)}
ORDER = [0, 1]   # streamed exponential pass off / on (split publication on in both)


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


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    data = {}
    for name, (B, S, heads, d) in SHAPES.items():
        dt = torch.bfloat16 if d == 128 else torch.float16
        torch.manual_seed(1)
        qkv = torch.randn((B * S, 3 * heads * d), device=DEV, dtype=dt)
        data[name] = (qkv, torch.empty((B * S, heads * d), device=DEV, dtype=dt))
    base = {}
    for rnd in range(rounds):
        for fl in ORDER:
            for poly in (0, 1, 2):
                line = [f"r{rnd} stream={fl} poly={poly}"]
                for name, (B, S, heads, d) in SHAPES.items():
                    qkv, o = data[name]
                    lib.dk_attention_tuning(1, poly, fl)
                    o.zero_()
                    ops.attention(qkv, B, S, heads, d, o)
                    torch.cuda.synchronize()
                    if rnd == 0:
                        if fl == 0:
                            base[(name, poly)] = o.clone()
                            ref = base[(name, 0)].float()
                            err = ((o.float() - ref).norm() / ref.norm()).item()
                            chk = f"vs_poly0={err:.1e}"
                        else:
                            same = torch.equal(o, base[(name, poly)])
                            chk = "same" if same else "DIFF(%.1e)" % ((o.float() - base[(name, poly)].float()).abs().max().item())
                    else:
                        chk = ""
                    if name in ("c4", "sd3", "c2"):
                        t = timeit(lambda: ops.attention(qkv, B, S, heads, d, o))
                        line.append(f"{name} {4.0 * B * heads * S * S * d / t / 1e12:6.0f} TF/s {chk}")
                    else:
                        line.append(f"{name} {chk}")
                print(" | ".join(line), flush=True)
    lib.dk_attention_tuning(-1, -1, -1)


if __name__ == "__main__":
    main()
