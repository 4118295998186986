"""ORACLE (test infrastructure — never imported by the product path).

numpy restatement of MLX's affine weight quantisation as the reference uses it for the `*-4bit-quantized` model versions
(`nn.quantize(model)` with MLX defaults group_size=64, bits=4 — python/src/diffusionkit/mlx/model_io.py:728-734, 772-775;
the arithmetic lives in the un-vendored dependency mlx==0.17.3, `mx.quantize` / `mx.dequantize` / `mx.quantized_matmul`).

PARITY UNPINNED: MLX cannot run here and the repo holds no quantised fixtures.  What the engine has to match is the
published DEQUANTISATION rule, which is unambiguous:
    w[n, k] = scales[n, k // group] * q[n, k] + biases[n, k // group],   q[n, k] = (wq[n, k // 8] >> 4 * (k % 8)) & 0xF
`quantize_q4` below (plain min/max affine) only manufactures test checkpoints; MLX's own quantiser may pick slightly
different scales for the same weights, which does not matter for reading a checkpoint MLX wrote.
"""
from __future__ import annotations

import numpy as np


def quantize_q4(w: np.ndarray, group_size: int = 64):
    """w [N, K] float -> (wq uint32 [N, K/8], scales fp32 [N, K/group], biases fp32 [N, K/group])"""
    N, K = w.shape
    assert K % group_size == 0 and group_size % 8 == 0
    g = w.astype(np.float32).reshape(N, K // group_size, group_size)
    lo, hi = g.min(axis=-1), g.max(axis=-1)
    scales = np.maximum((hi - lo) / 15.0, 1e-7).astype(np.float32)
    biases = lo.astype(np.float32)
    q = np.clip(np.rint((g - biases[..., None]) / scales[..., None]), 0, 15).astype(np.uint32).reshape(N, K // 8, 8)
    wq = np.zeros((N, K // 8), dtype=np.uint32)
    for j in range(8):
        wq |= q[:, :, j] << np.uint32(4 * j)
    return wq, scales, biases


def unpack_q4(wq: np.ndarray) -> np.ndarray:
    """wq uint32 [N, K/8] -> q uint8 [N, K]"""
    N, W = wq.shape
    q = np.empty((N, W, 8), dtype=np.uint8)
    for j in range(8):
        q[:, :, j] = (wq >> np.uint32(4 * j)) & np.uint32(0xF)
    return q.reshape(N, W * 8)


def dequantize_q4(wq: np.ndarray, scales: np.ndarray, biases: np.ndarray, group_size: int = 64) -> np.ndarray:
    """-> fp32 [N, K]; scale * q + bias evaluated as one fused multiply-add in fp32 (q <= 15 and a 16-bit scale make the
    product exact in fp32, so FMA and mul+add agree)"""
    q = unpack_q4(wq).astype(np.float32)
    N, K = q.shape
    s = np.repeat(scales.astype(np.float32), group_size, axis=1)
    b = np.repeat(biases.astype(np.float32), group_size, axis=1)
    return (s.astype(np.float64) * q.astype(np.float64) + b.astype(np.float64)).astype(np.float32)
