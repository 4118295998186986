"""ORACLE (test infrastructure — never imported by the product path).

CPU PyTorch fp32 restatement of the reference text encoders: CLIPTextModel (python/src/diffusionkit/mlx/clip.py:27-120)
and SD3T5Encoder (mlx/t5.py:21-243, 316-325), plus the token batching of DiffusionPipeline._tokenize / encode_text
(mlx/__init__.py:174-251, 642-671).

PARITY STATUS: pinned twice.  (1) Against the reference's MLX source (mlx/clip.py, mlx/t5.py) executed from
/root/reference on the torch-backed MLX stand-in (tests/golden/mlx_standin.py): tests/test_reference_mlxsrc_pin_cpu.py,
live in the build container, 3e-4 / 5e-4.  (2) Against transformers' own CLIPTextModel / T5EncoderModel (installed
here, CPU) on random small configs — an independent implementation of the same published architectures — in
tests/test_text_cpu.py, which also runs on the GPU box.  MLX's kernel numerics themselves cannot be pinned (MLX does
not run here).

Parameters: flat dicts with the reference's module-tree names.  `dt` (optional torch dtype) rounds every op output to
that type, mimicking the reference's 16-bit activations; dt=None computes in fp32.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _r(x, dt):
    return x if dt is None else x.to(dt).to(torch.float32)


def quick_gelu(x):
    """mlx nn.gelu_fast_approx: x * sigmoid(1.702 x)   (clip.py:11)"""
    return x * torch.sigmoid(1.702 * x)


class CLIPTextModelRef:
    def __init__(self, params: Dict[str, torch.Tensor], num_layers: int, num_heads: int, hidden_act: str = "quick_gelu",
                 dt: Optional[torch.dtype] = None):
        self.p = {k: v.float() for k, v in params.items()}
        self.L, self.H, self.dt = num_layers, num_heads, dt
        self.act = quick_gelu if hidden_act == "quick_gelu" else (lambda x: F.gelu(x))

    def _lin(self, x, name):
        y = x @ self.p[name + ".weight"].t()
        if (name + ".bias") in self.p:
            y = y + self.p[name + ".bias"]
        return _r(y, self.dt)

    def _ln(self, x, name):
        return _r(F.layer_norm(x, (x.shape[-1],), self.p[name + ".weight"], self.p[name + ".bias"], 1e-5), self.dt)

    def _attention(self, y, name, mask):
        """mlx nn.MultiHeadAttention: softmax((q * d^-1/2) k^T + mask) v   (clip.py:52)"""
        B, N, D = y.shape
        H = self.H
        q = self._lin(y, name + ".query_proj").reshape(B, N, H, -1).transpose(1, 2)
        k = self._lin(y, name + ".key_proj").reshape(B, N, H, -1).transpose(1, 2)
        v = self._lin(y, name + ".value_proj").reshape(B, N, H, -1).transpose(1, 2)
        scale = math.sqrt(1 / q.shape[-1])
        s = (q * scale) @ k.transpose(-1, -2) + mask
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, N, D)
        return self._lin(_r(o, self.dt), name + ".out_proj")

    def __call__(self, tokens: torch.Tensor):
        """-> (pooled_output, last_hidden_state, hidden_states)   (clip.py:91-120)"""
        B, N = tokens.shape
        eos = tokens.argmax(-1)
        x = _r(self.p["token_embedding.weight"][tokens] + self.p["position_embedding.weight"][:N], self.dt)
        idx = torch.arange(N)
        mask = (idx[:, None] < idx[None]).float() * (-6e4 if self.dt in (torch.float16, torch.bfloat16) else -1e9)
        hidden: List[torch.Tensor] = []
        for i in range(self.L):
            l = f"layers.{i}"
            y = self._ln(x, l + ".layer_norm1")
            x = _r(self._attention(y, l + ".attention", mask) + x, self.dt)
            y = self._ln(x, l + ".layer_norm2")
            y = _r(self.act(self._lin(y, l + ".linear1")), self.dt)
            x = _r(self._lin(y, l + ".linear2") + x, self.dt)
            hidden.append(x)
        last = self._ln(x, "final_layer_norm")
        pooled = last[torch.arange(B), eos]
        if "text_projection.weight" in self.p:
            pooled = self._lin(pooled, "text_projection")
        return pooled, last, hidden


def relative_position_bucket(relative_position: torch.Tensor, num_buckets: int = 32, max_distance: int = 128):
    """bidirectional branch of mlx/t5.py:21-64 (HF's T5 bucket rule; the reference truncates the log term through
    int16 and multiplies log(n / max_exact) by a precomputed fp32 scale)"""
    nb = num_buckets // 2
    ret = (relative_position > 0).long() * nb
    n = relative_position.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    scale = torch.tensor((nb - max_exact) / np.log(max_distance / max_exact), dtype=torch.float32)
    large = max_exact + (torch.log(n.clamp(min=1).float() / max_exact) * scale).to(torch.int16).long()
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


class T5EncoderRef:
    def __init__(self, params: Dict[str, torch.Tensor], num_layers: int, num_heads: int, eps: float = 1e-6,
                 num_buckets: int = 32, max_distance: int = 128, dt: Optional[torch.dtype] = None):
        self.p = {k: v.float() for k, v in params.items()}
        self.L, self.H, self.eps, self.dt = num_layers, num_heads, eps, dt
        self.nb, self.md = num_buckets, max_distance

    def _rms(self, x, name):
        """t5.py:150-170 — fp32 norm, then weight"""
        d = x.shape[-1]
        n = x * torch.rsqrt((x * (1.0 / math.sqrt(d))).square().sum(-1, keepdim=True) + self.eps)
        return self.p[name + ".weight"] * n

    def position_bias(self, L: int):
        """t5.py:79-102 -> (heads, L, L)"""
        ctx = torch.arange(L)[:, None]
        mem = torch.arange(L)[None, :]
        b = relative_position_bucket(mem - ctx, self.nb, self.md)
        return self.p["encoder.relative_attention_bias.embeddings.weight"][b].permute(2, 0, 1)

    def __call__(self, tokens: torch.Tensor) -> torch.Tensor:
        B, L = tokens.shape
        H = self.H
        x = _r(self.p["wte.weight"][tokens], self.dt)
        bias = self.position_bias(L)
        for i in range(self.L):
            l = f"encoder.layers.{i}"
            y = _r(self._rms(x, l + ".ln1"), self.dt)                                  # cast to the weight dtype, :216
            q = _r(y @ self.p[l + ".attention.query_proj.weight"].t(), self.dt).reshape(B, L, H, -1).transpose(1, 2)
            k = _r(y @ self.p[l + ".attention.key_proj.weight"].t(), self.dt).reshape(B, L, H, -1).transpose(1, 2)
            v = _r(y @ self.p[l + ".attention.value_proj.weight"].t(), self.dt).reshape(B, L, H, -1).transpose(1, 2)
            s = _r(_r(q @ k.transpose(-1, -2), self.dt) + bias, self.dt)                # no 1/sqrt(d) in T5, :134-136
            pr = _r(torch.softmax(s, dim=-1), self.dt)
            o = _r(pr @ v, self.dt).transpose(1, 2).reshape(B, L, -1)
            x = x + _r(o @ self.p[l + ".attention.out_proj.weight"].t(), self.dt)        # fp32 residual, :218-219
            y = self._rms(x, l + ".ln2")                                                  # stays fp32 (:221-223)
            g = F.gelu(y @ self.p[l + ".dense.wi_0.weight"].t()) * (y @ self.p[l + ".dense.wi_1.weight"].t())
            x = x + g @ self.p[l + ".dense.wo.weight"].t()
        return _r(self._rms(x, "encoder.ln"), self.dt)


def tokenize_pair(tokenizer, text: str, negative_text: Optional[str]) -> torch.Tensor:
    """DiffusionPipeline._tokenize (mlx/__init__.py:174-195)"""
    if negative_text is None:
        negative_text = ""
    pad = tokenizer.eos_token if tokenizer.pad_with_eos else 0
    tokens = [list(tokenizer.tokenize(text))]
    if tokenizer.pad_to_max_length:
        tokens[0].extend([pad] * (tokenizer.max_length - len(tokens[0])))
    tokens += [list(tokenizer.tokenize(negative_text))]
    n = max(len(t) for t in tokens)
    return torch.tensor([t + [pad] * (n - len(t)) for t in tokens], dtype=torch.int64)


def encode_text_sd3(clip_l: CLIPTextModelRef, clip_g: CLIPTextModelRef, t5: Optional["T5EncoderRef"], tokens_l, tokens_g,
                    tokens_t5=None):
    """DiffusionPipeline.encode_text after tokenisation (mlx/__init__.py:212-251): hidden_states[-2] of both CLIPs side by
    side, zero-padded to 4096 channels, T5 sequence appended along the token axis; pooled outputs side by side"""
    pl, _, hl = clip_l(tokens_l)
    pg, _, hg = clip_g(tokens_g)
    cond = torch.cat([hl[-2], hg[-2]], dim=-1)
    pooled = torch.cat([pl, pg], dim=-1)
    cond = torch.cat([cond, torch.zeros(cond.shape[0], cond.shape[1], 4096 - cond.shape[2])], dim=-1)
    t5c = t5(tokens_t5) if t5 is not None else torch.zeros_like(cond)
    return torch.cat([cond, t5c], dim=1), pooled


def encode_text_flux(clip_l: CLIPTextModelRef, t5: "T5EncoderRef", tokens_l, tokens_t5, t5_max_length: int):
    """FluxPipeline.encode_text after tokenisation (mlx/__init__.py:642-671): positive prompt only; the T5 ids are copied
    into a zero buffer of T5_MAX_LENGTH"""
    pooled, _, _ = clip_l(tokens_l[[0]])
    padded = torch.zeros((1, t5_max_length), dtype=tokens_t5.dtype)
    padded[:, : tokens_t5.shape[1]] = tokens_t5[[0]]
    return t5(padded), pooled
