"""Micro-benchmarks of the hot kernels at the FLUX C4 shapes (B=4, S=4096+256, h=3072): TFLOP/s with CUDA events.
Writes gpurun_out/bench_kernels.json."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops  # noqa: E402
from diffusionkit_b200._lib import ACT_GELU_ERF  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    res = {}
    dt = torch.bfloat16
    gemms = {
        "qkv_img  16384x9216x3072": (16384, 9216, 3072),
        "o_img    16384x3072x3072": (16384, 3072, 3072),
        "fc1_img  16384x12288x3072": (16384, 12288, 3072),
        "fc2_img  16384x3072x12288": (16384, 3072, 12288),
        "qkv_txt  1024x9216x3072": (1024, 9216, 3072),
        "single_out 17408x3072x15360": (17408, 3072, 15360),
        "single_fc1 17408x12288x3072": (17408, 12288, 3072),
        "c2_qkv  1280x9216x3072": (1280, 9216, 3072),
        "cublas_ref_shape 8192x8192x8192": (8192, 8192, 8192),
    }
    for name, (M, N, K) in gemms.items():
        A = torch.randn((M, K), device=DEV, dtype=dt)
        W = torch.randn((N, K), device=DEV, dtype=dt) * (1 / math.sqrt(K))
        out = torch.empty((M, N), device=DEV, dtype=dt)
        bias = torch.randn((N,), device=DEV, dtype=dt)
        med, best = timeit(lambda: ops.gemm(A, W, out=out, bias=bias))
        fl = 2.0 * M * N * K
        res["gemm " + name] = {"ms": med * 1e3, "tflops": fl / med / 1e12, "tflops_best": fl / best / 1e12}
        if "fc1_img" in name:
            med, best = timeit(lambda: ops.gemm(A, W, out=out, bias=bias, act=ACT_GELU_ERF))
            res["gemm+gelu " + name] = {"ms": med * 1e3, "tflops": fl / med / 1e12}
            med, best = timeit(lambda: torch.matmul(A, W.t(), out=out))
            res["cublas " + name] = {"ms": med * 1e3, "tflops": fl / med / 1e12}
        if "8192" in name:
            med, best = timeit(lambda: torch.matmul(A, W.t(), out=out))
            res["cublas " + name] = {"ms": med * 1e3, "tflops": fl / med / 1e12, "tflops_best": fl / best / 1e12}
        del A, W, out
    for name, (B, S, heads, d) in {"flux_c4 B4 S4352 H24 d128": (4, 4352, 24, 128),
                                   "flux_c2 B1 S1280 H24 d128": (1, 1280, 24, 128),
                                   "sd3_c3 B8 S4685 H24 d64": (8, 4685, 24, 64)}.items():
        h = heads * d
        dtt = torch.bfloat16 if d == 128 else torch.float16
        qkv = torch.randn((B * S, 3 * h), device=DEV, dtype=dtt)
        o = torch.empty((B * S, h), device=DEV, dtype=dtt)
        med, best = timeit(lambda: ops.attention(qkv, B, S, heads, d, o))
        fl = 4.0 * B * heads * S * S * d
        res["attention " + name] = {"ms": med * 1e3, "tflops": fl / med / 1e12, "tflops_best": fl / best / 1e12}
        del qkv, o
    # memory-bound kernels
    x = torch.randn((17408, 3072), device=DEV, dtype=dt)
    mod = torch.randn((4, 6 * 3072), device=DEV, dtype=dt)
    y = torch.empty_like(x)
    med, _ = timeit(lambda: ops.ln_modulate(x, mod[:, :3072], mod[:, 3072:6144], 4352, out=y))
    res["ln_modulate 17408x3072"] = {"ms": med * 1e3, "gbs": 2 * x.numel() * 2 / med / 1e9}
    qkv = torch.randn((17408, 9216), device=DEV, dtype=dt)
    w = torch.ones(128, device=DEV, dtype=dt)
    rope = torch.rand((4352, 64, 2), device=DEV)
    med, _ = timeit(lambda: ops.qk_norm_rope(qkv, 4352, 24, 128, 256, w, w, w, w, rope))
    res["qk_norm_rope 17408x(2x3072)"] = {"ms": med * 1e3, "gbs": 2 * 17408 * 6144 * 2 / med / 1e9}
    del x, y, qkv
    # VAE conv shapes
    for name, (B, H, W, Cin, Cout) in {"conv 128x128 512->512": (1, 128, 128, 512, 512),
                                       "conv 512x512 256->256": (1, 512, 512, 256, 256),
                                       "conv 1024x1024 128->128": (1, 1024, 1024, 128, 128)}.items():
        xx = torch.randn((B, H, W, Cin), device=DEV, dtype=dt)
        ww = torch.randn((Cout, 3, 3, Cin), device=DEV, dtype=dt) * 0.02
        bb = torch.zeros((Cout,), device=DEV, dtype=dt)
        oo = torch.empty((B, H, W, Cout), device=DEV, dtype=dt)
        med, best = timeit(lambda: ops.conv3x3(xx, ww, bias=bb, out=oo), iters=5)
        fl = 2.0 * B * H * W * Cout * 9 * Cin
        res[name] = {"ms": med * 1e3, "tflops": fl / med / 1e12}
        del xx, ww, oo
    # VAE HBM-bound kernels at the 1024x1024 level (B=1): GroupNorm statistics / apply, nearest upsample, image post
    for name, (H, C) in {"512x512x256": (512, 256), "1024x1024x128": (1024, 128)}.items():
        xx = torch.randn((1, H, H, C), device=DEV, dtype=dt)
        gam = torch.ones(C, device=DEV, dtype=dt)
        bet = torch.zeros(C, device=DEV, dtype=dt)
        ws = torch.empty(1 << 20, device=DEV, dtype=torch.float32)
        yy = torch.empty_like(xx)
        med, _ = timeit(lambda: ops.groupnorm_stats(xx, 32, 1e-5, ws=ws), iters=5)
        res["groupnorm_stats " + name] = {"ms": med * 1e3, "gbs": xx.numel() * 2 / med / 1e9}
        st = ops.groupnorm_stats(xx, 32, 1e-5, ws=ws)
        med, _ = timeit(lambda: ops.groupnorm_apply(xx, st, gam, bet, 32, True, out=yy), iters=5)
        res["groupnorm_apply+silu " + name] = {"ms": med * 1e3, "gbs": 2 * xx.numel() * 2 / med / 1e9}
        del xx, yy
    xs = torch.randn((1, 512, 512, 256), device=DEV, dtype=dt)
    ys = torch.empty((1, 1024, 1024, 256), device=DEV, dtype=dt)
    med, _ = timeit(lambda: ops.upsample_nearest2x(xs, out=ys), iters=5)
    res["upsample2x 512->1024 x256"] = {"ms": med * 1e3, "gbs": (xs.numel() + ys.numel()) * 2 / med / 1e9}
    del xs, ys
    sc = torch.randn((16384, 16384), device=DEV, dtype=dt)
    med, _ = timeit(lambda: ops.softmax_rows(sc, 0.044), iters=5)
    res["softmax_rows 16384x16384"] = {"ms": med * 1e3, "gbs": 2 * sc.numel() * 2 / med / 1e9}
    del sc
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(f"{k:45s} " + " ".join(f"{a}={b:.2f}" for a, b in v.items()))


if __name__ == "__main__":
    main()
