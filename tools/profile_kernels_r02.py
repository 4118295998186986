"""Two launches of every kernel on the C4 / 1024^2-decode path at its real shape, for ONE `ncu --set full` capture
(tools/ncu_to_json.py keeps the second launch of each kernel):
  ncu --set full --clock-control none --import-source on -k 'regex:(gemm|attention|conv|groupnorm|softmax_rows|ln_modulate).*_kernel' \
      -c 80 -o gpurun_out/prof_r02_kernels python tools/profile_kernels_r02.py
(ncu's -k matches the FUNCTION name without its namespace: `regex:dk::` selects nothing)"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionkit_b200 import ops  # noqa: E402
from diffusionkit_b200._lib import ACT_GELU_ERF  # noqa: E402

DEV = "cuda:0"
bf = torch.bfloat16


def r(shape, dt=bf, s=1.0):
    return (torch.randn(shape, device=DEV) * s).to(dt)


def twice(fn):
    if os.environ.get("DK_PROFILE_ONCE", "1") != "1":
        fn()
    fn()


# ---- MMDiT: pair GEMMs (fc1 + GELU, single-block output), small-M GEMM, LN-modulate, attention (both head dims)
A, W, b = r((16384, 3072)), r((12288, 3072), s=1 / math.sqrt(3072)), r((12288,))
out = torch.empty((16384, 12288), device=DEV, dtype=bf)
twice(lambda: ops.gemm(A, W, out=out, bias=b, act=ACT_GELU_ERF))
A2, W2 = r((17408, 15360)), r((3072, 15360), s=1 / math.sqrt(15360))
x2, g2 = r((17408, 3072)), r((4, 3072))
twice(lambda: ops.gemm(A2, W2, out=x2, bias=b[:3072], gate=g2, res=x2, rows_per_batch=4352, out_batch_rows=4352))
A3, W3 = r((1024, 3072)), r((9216, 3072), s=1 / math.sqrt(3072))
twice(lambda: ops.gemm(A3, W3, bias=b[:9216]))
mod = r((4, 6 * 3072))
twice(lambda: ops.ln_modulate(x2, mod[:, :3072], mod[:, 3072:6144], 4352))
qkv = r((4 * 4352, 3 * 3072))
o = torch.empty((4 * 4352, 3072), device=DEV, dtype=bf)
twice(lambda: ops.attention(qkv, 4, 4352, 24, 128, o))
qkv64 = r((4 * 4685, 3 * 1536), torch.float16)
o64 = torch.empty((4 * 4685, 1536), device=DEV, dtype=torch.float16)
twice(lambda: ops.attention(qkv64, 4, 4685, 24, 64, o64))
del A, W, out, A2, W2, qkv, o, qkv64, o64

# ---- VAE decoder at 1024^2 (batch 2): fused convs, the remaining stand-alone GroupNorm / softmax / conv_out kernels
G = 32
x = r((2, 1024, 1024, 128))
w = r((128, 3, 3, 128), s=1 / math.sqrt(9 * 128))
gam, bet = torch.ones(128, device=DEV, dtype=bf), torch.zeros(128, device=DEV, dtype=bf)
stats = ops.groupnorm_stats(x, G, 1e-5)
part = torch.empty((2, 1024 * 1024 // 128, G, 2), dtype=torch.float32, device=DEV)
y = torch.empty_like(x)
twice(lambda: ops.conv3x3_fused(x, w, bias=b[:128], res=x, out=y, gn=(stats, gam, bet, G), silu=True, out_partial=part, out_G=G))
twice(lambda: ops.groupnorm_finalize(part, 2, G, 1024 * 1024 // 128, float(1024 * 1024 * 4), 1e-5))
twice(lambda: ops.groupnorm_apply(x, stats, gam, bet, G, True, out=y))
w8 = r((8, 3, 3, 128), s=1 / math.sqrt(9 * 128))
twice(lambda: ops.conv3x3(y, w8, b[:8]))
x5 = r((2, 512, 512, 256))
w5 = r((256, 3, 3, 256), s=1 / math.sqrt(9 * 256))
y5 = torch.empty_like(x5)
part5 = torch.empty((2, 512 * 512 // 128, G, 2), dtype=torch.float32, device=DEV)
twice(lambda: ops.conv3x3_fused(x5, w5, bias=b[:256], out=y5, out_partial=part5, out_G=G))
wp = ops.conv_up_weights(w5)
yu = torch.empty((2, 1024, 1024, 256), device=DEV, dtype=bf)
twice(lambda: ops.conv3x3_fused(x5, wp, bias=b[:256], out=yu, up=True))
del x, y, yu, x5, y5
q, k = r((16384, 512)), r((16384, 512))
sc = torch.empty((16384, 16384), device=DEV, dtype=bf)
twice(lambda: ops.gemm(q, k, out=sc))
twice(lambda: ops.softmax_rows(sc, 1.0))
ov = torch.empty((16384, 512), device=DEV, dtype=bf)
twice(lambda: ops.gemm(sc, k, out=ov, w_n_major=True))
torch.cuda.synchronize()
print("done")
