"""Test-side glue between the product's config/weights and the oracle (only tests/, smoke() and bench.py's CPU-baseline
leg touch oracle/)."""
import torch

from diffusionkit_b200.config import MMDiTConfig, PositionalEncoding
from oracle.mmdit_ref import RefMMDiTConfig


def ref_config(cfg: MMDiTConfig) -> RefMMDiTConfig:
    return RefMMDiTConfig(
        num_heads=cfg.num_heads, depth_multimodal=cfg.depth_multimodal, depth_unified=cfg.depth_unified,
        hidden_size=cfg.hidden_size, mlp_ratio=cfg.mlp_ratio, vae_latent_dim=cfg.vae_latent_dim,
        layer_norm_eps=cfg.layer_norm_eps, use_rope=cfg.pos_embed_type == PositionalEncoding.PreSDPARope,
        rope_axes_dim=tuple(cfg.rope_axes_dim or ()), use_qk_norm=cfg.use_qk_norm,
        max_latent_resolution=cfg.max_latent_resolution, patch_size=cfg.patch_size,
        patchify_via_reshape=cfg.patchify_via_reshape, pooled_text_embed_dim=cfg.pooled_text_embed_dim,
        token_level_text_embed_dim=cfg.token_level_text_embed_dim, frequency_embed_dim=cfg.frequency_embed_dim,
        max_period=cfg.max_period, dtype=cfg.dtype,
        parallel_mlp_for_unified_blocks=cfg.parallel_mlp_for_unified_blocks)


def to_f32(params):
    return {k: v.detach().to("cpu", torch.float32) for k, v in params.items()}
