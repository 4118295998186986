"""ORACLE (test infrastructure — never imported by the product path).

CPU PyTorch fp32 restatement of the reference VAE decoder and encoder (python/src/diffusionkit/mlx/vae.py:20-149,
336-401, 404-467), of decode_latents_to_image (mlx/__init__.py:581-584) and of read_image / encode_image_to_latents
(mlx/__init__.py:536-551, 586-594).

PARITY STATUS: decoder AND encoder are pinned against the reference's MLX source (python/src/diffusionkit/mlx/vae.py)
executed from /root/reference on the torch-backed MLX stand-in (tests/golden/mlx_standin.py):
tests/golden/reference_mlxsrc_vae.npz, reproduced to 3e-4 by tests/test_reference_mlxsrc_pin_cpu.py; the decoder
additionally against the reference's PyTorch twin (python/src/diffusionkit/torch/vae.py, `gn_eps` = 1e-6 there):
tests/golden/reference_torch_vae_decoder.npz, tests/test_reference_pin_cpu.py.  MLX's own kernel numerics cannot be
pinned here (MLX does not run in this container).

Parameters: flat dict with the reference's names; conv weights (O, kh, kw, I), Linear (out, in).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _r(x, dt):
    return x if dt is None else x.to(dt).to(torch.float32)


def conv3x3(x, w, b, dt=None):
    """mlx nn.Conv2d(k=3, stride 1, padding 1) on NHWC, weight (O, kh, kw, I) (App. A.3)"""
    y = F.conv2d(x.permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b.float(), padding=1)
    return _r(y.permute(0, 2, 3, 1), dt)


def conv3x3_s2(x, w, b, dt=None):
    """EncoderDecoderBlock2D downsample (vae.py:142-144): mx.pad [(0,0),(0,1),(0,1),(0,0)] then Conv2d(k=3, stride 2,
    padding 0)"""
    xp = F.pad(x.permute(0, 3, 1, 2), (0, 1, 0, 1))
    y = F.conv2d(xp, w.float().permute(0, 3, 1, 2), b.float(), stride=2)
    return _r(y.permute(0, 2, 3, 1), dt)


def group_norm(x, w, b, groups=32, eps=1e-5, dt=None):
    """mlx nn.GroupNorm(32, C, pytorch_compatible=True), default eps 1e-5 (vae.py:34,72,78; quirk Q7)"""
    y = F.group_norm(x.permute(0, 3, 1, 2), groups, w.float(), b.float(), eps)
    return _r(y.permute(0, 2, 3, 1), dt)


def upsample_nearest(x, scale=2):
    """vae.py:20-25"""
    return x.repeat_interleave(scale, dim=1).repeat_interleave(scale, dim=2)


class _VAEBlocksRef:
    # GroupNorm epsilon: 1e-5 on the MLX path (mlx nn.GroupNorm default, quirk Q7).  The reference's PyTorch twin uses
    # 1e-6 (python/src/diffusionkit/torch/vae.py:20); tests/test_reference_pin_cpu.py sets it to pin this oracle
    # against that module.
    gn_eps = 1e-5

    def __init__(self, params: Dict[str, torch.Tensor], act_dtype: Optional[torch.dtype],
                 block_out_channels, layers_per_block: int, groups: int):
        self.p = params
        self.dt = act_dtype
        self.boc = list(block_out_channels)
        self.layers = layers_per_block
        self.groups = groups

    def _gn(self, x, name):
        return group_norm(x, self.p[name + ".weight"], self.p[name + ".bias"], self.groups, self.gn_eps, self.dt)

    def _conv(self, x, name):
        return conv3x3(x, self.p[name + ".weight"], self.p[name + ".bias"], self.dt)

    def _lin(self, x, name):
        return _r(x @ self.p[name + ".weight"].float().t() + self.p[name + ".bias"].float(), self.dt)

    def resnet(self, x, name):
        """ResnetBlock2D.__call__ (vae.py:86-101)"""
        y = self._gn(x, name + ".norm1")
        y = _r(F.silu(y), self.dt)
        y = self._conv(y, name + ".conv1")
        y = self._gn(y, name + ".norm2")
        y = _r(F.silu(y), self.dt)
        y = self._conv(y, name + ".conv2")
        if (name + ".conv_shortcut.weight") in self.p:
            x = self._lin(x, name + ".conv_shortcut")
        return _r(y + x, self.dt)

    def attention(self, x, name):
        """Attention.__call__ (vae.py:40-57): single head, scores materialised, softmax in the activation dtype"""
        B, H, W, C = x.shape
        y = self._gn(x, name + ".group_norm")
        q = self._lin(y, name + ".query_proj").reshape(B, H * W, C)
        k = self._lin(y, name + ".key_proj").reshape(B, H * W, C)
        v = self._lin(y, name + ".value_proj").reshape(B, H * W, C)
        scale = 1 / math.sqrt(C)
        scores = _r(_r(q * scale, self.dt) @ k.transpose(1, 2), self.dt)
        attn = _r(torch.softmax(scores, dim=-1), self.dt)
        y = _r(attn @ v, self.dt).reshape(B, H, W, C)
        y = self._lin(y, name + ".out_proj")
        return _r(x + y, self.dt)


class VAEDecoderRef(_VAEBlocksRef):
    def __init__(self, params: Dict[str, torch.Tensor], act_dtype: Optional[torch.dtype] = None,
                 block_out_channels=(128, 256, 512, 512), layers_per_block: int = 3, groups: int = 32):
        super().__init__(params, act_dtype, block_out_channels, layers_per_block, groups)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """VAEDecoder.__call__ (vae.py:386-401): x (B, H, W, 16) -> (B, 8H, 8W, 3)"""
        x = self._conv(_r(x.float(), self.dt), "conv_in")
        x = self.resnet(x, "mid_blocks.0")
        x = self.attention(x, "mid_blocks.1")
        x = self.resnet(x, "mid_blocks.2")
        n = len(self.boc)
        for j in reversed(range(n)):                      # reversed(self.up_blocks) (vae.py:393)
            for l in range(self.layers):
                x = self.resnet(x, f"up_blocks.{j}.resnets.{l}")
            if (f"up_blocks.{j}.upsample.weight") in self.p:
                x = self._conv(upsample_nearest(x), f"up_blocks.{j}.upsample")   # vae.py:146-147
        x = self._gn(x, "conv_norm_out")
        x = _r(F.silu(x), self.dt)
        return self._conv(x, "conv_out")


class VAEEncoderRef(_VAEBlocksRef):
    def __init__(self, params: Dict[str, torch.Tensor], act_dtype: Optional[torch.dtype] = None,
                 block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, groups: int = 32):
        super().__init__(params, act_dtype, block_out_channels, layers_per_block, groups)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """VAEEncoder.__call__ (vae.py:453-467): x (B, H, W, 3) in [-1, 1] -> (B, H/8, W/8, 32) = (mean | logvar)"""
        x = self._conv(_r(x.float(), self.dt), "conv_in")
        for i in range(len(self.boc)):
            for l in range(self.layers):
                x = self.resnet(x, f"down_blocks.{i}.resnets.{l}")
            name = f"down_blocks.{i}.downsample"
            if (name + ".weight") in self.p:
                x = conv3x3_s2(x, self.p[name + ".weight"], self.p[name + ".bias"], self.dt)
        x = self.resnet(x, "mid_blocks.0")
        x = self.attention(x, "mid_blocks.1")
        x = self.resnet(x, "mid_blocks.2")
        x = self._gn(x, "conv_norm_out")
        x = _r(F.silu(x), self.dt)
        return self._conv(x, "conv_out")


def read_image_array(img_u8: torch.Tensor) -> torch.Tensor:
    """read_image after PIL (mlx/__init__.py:548-551): uint8 (H, W, >=3) -> (1, H, W, 3) fp32 in [-1, 1]"""
    return ((img_u8[:, :, :3].to(torch.float32) / 255) * 2 - 1.0).unsqueeze(0)


def encode_image_to_latents(encoder: VAEEncoderRef, image: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """mlx/__init__.py:586-594 with the seeded draw passed in: mean + exp(0.5 * clip(logvar, -30, 20)) * noise"""
    hidden = encoder(image)
    mean, logvar = hidden.split(hidden.shape[-1] // 2, dim=-1)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return mean + std * noise.float()


def decode_latents_to_image(decoder: VAEDecoderRef, latent: torch.Tensor) -> torch.Tensor:
    """__init__.py:581-584: clip(decoder(x) / 2 + 0.5, 0, 1)"""
    x = decoder(latent)
    return _r(torch.clamp(_r(x / 2, decoder.dt) + 0.5, 0, 1), decoder.dt)


def to_uint8(img: torch.Tensor, dt=None) -> torch.Tensor:
    """__init__.py:526: (x * 255).astype(uint8) — truncation"""
    return _r(img * 255, dt).to(torch.uint8)
