"""ORACLE (test infrastructure — never imported by the product path).

CPU PyTorch fp32 restatement of the reference MMDiT forward (argmaxinc/DiffusionKit @ 498e5dba,
python/src/diffusionkit/mlx/mmdit.py) for SD3 (dual-stream only) and FLUX (dual + single stream).

PARITY STATUS: MLX 0.17.3 (setup.py:32) cannot run here (Metal-only, no wheel, no network) and the repository holds no
golden tensors for this path, but this restatement is pinned against the reference's own source code:
  * FLUX and SD3 forward + modulation cache: the reference's MLX source (python/src/diffusionkit/mlx/mmdit.py) imported
    from /root/reference and executed unmodified on a torch-backed stand-in for the MLX primitives it calls
    (tests/golden/mlx_standin.py); fixtures tests/golden/reference_mlxsrc_{flux,sd3}_mmdit.npz, reproduced to 3e-4 by
    tests/test_reference_mlxsrc_pin_cpu.py.  Pins the wiring, not MLX's kernel numerics.
  * SD3 forward: additionally the reference's PyTorch twin (python/src/diffusionkit/torch/mmdit.py) —
    tests/golden/reference_torch_mmdit.npz, tests/test_reference_pin_cpu.py (tanh GELU switched by `gelu_tanh`).
  * schedule / noise known-answer values derived from the reference formulas (tests/golden/schedule_kats.json).

Conventions: parameters are a flat dict name -> tensor using the reference's module-tree names (SURVEY.md App. C,
e.g. "multimodal_transformer_blocks.3.image_transformer_block.attn.q_proj.weight").  Linear weights are (out, in)
(mlx nn.Linear), the SD3 patch conv weight is (O, kh, kw, I) (mlx nn.Conv2d).

`act_dtype` (optional) emulates the reference's 16-bit activations by rounding at module boundaries
(Linear outputs, norms, residual adds); None = pure fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class RefMMDiTConfig:
    """Subset of MMDiTConfig the forward reads (reference mlx/config.py:19-71)."""

    num_heads: int = 24
    depth_multimodal: int = 24
    depth_unified: int = 0
    hidden_size: int = 1536
    mlp_ratio: int = 4
    vae_latent_dim: int = 16
    layer_norm_eps: float = 1e-6
    use_rope: bool = False                      # pos_embed_type == PreSDPARope
    rope_axes_dim: Tuple[int, ...] = (16, 56, 56)
    use_qk_norm: bool = False
    max_latent_resolution: int = 192
    patch_size: int = 2
    patchify_via_reshape: bool = False
    pooled_text_embed_dim: int = 2048
    token_level_text_embed_dim: int = 4096
    frequency_embed_dim: int = 256
    max_period: int = 10000
    dtype: torch.dtype = torch.float16          # config.dtype: sinusoid arithmetic dtype (quirk Q5)
    parallel_mlp_for_unified_blocks: bool = True
    # False = the MLX path's nn.GELU() (erf, mmdit.py:421).  True = GELU(approximate="tanh"), what the reference's
    # PyTorch twin uses (python/src/diffusionkit/torch/mmdit.py:242); only set by the test that pins this oracle
    # against that module (tests/test_reference_pin_cpu.py).
    gelu_tanh: bool = False


def _r(x: torch.Tensor, dt: Optional[torch.dtype]) -> torch.Tensor:
    """round-trip through the emulated activation dtype"""
    return x if dt is None else x.to(dt).to(torch.float32)


def linear(x, w, b=None, dt=None):
    """mlx nn.Linear: y = x W^T + b (App. A.3)"""
    y = x @ w.float().t()
    if b is not None:
        y = y + b.float()
    return _r(y, dt)


def layer_norm_noaffine(x, eps):
    """mx.fast.layer_norm(x, None, None, eps): last axis, biased variance, fp32 (mmdit.py:838-849)"""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps)


def affine_transform(x, shift, residual_scale, eps, dt=None):
    """mmdit.py:958-972 — LN(x) * (1 + scale) + shift (the B==1 fused single-rounding form)"""
    return _r(layer_norm_noaffine(x, eps) * (1.0 + residual_scale) + shift, dt)


def rms_norm(x, w, eps=1e-6, dt=None):
    """mlx nn.RMSNorm(d, eps=1e-6): x * rsqrt(mean(x^2) + eps) * w (mmdit.py:754-764)"""
    return _r(x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * w.float(), dt)


def gelu_erf(x):
    """mlx nn.GELU() default approx='none' (mmdit.py:421; quirk Q4)"""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def timestep_embedding(t: torch.Tensor, cfg: RefMMDiTConfig) -> torch.Tensor:
    """mmdit.py:379-389 — sinusoid computed in config.dtype arithmetic (quirk Q5); cos first."""
    half = cfg.frequency_embed_dim // 2
    cd = cfg.dtype
    ar = torch.arange(0, half, dtype=torch.float32).to(cd)
    # -log(max_period) is an fp32 scalar array in MLX; times a config.dtype array promotes to fp32, the result is
    # cast back with .astype(config.dtype) (mmdit.py:382-386)
    freqs = torch.exp(-math.log(cfg.max_period) * ar.float() / half).to(cd)
    args = (t.reshape(-1, 1).to(cd).float() * freqs.float()[None]).to(cd)   # product of two config.dtype arrays
    emb = torch.cat([torch.cos(args.float()).to(cd), torch.sin(args.float()).to(cd)], dim=-1)
    return emb.float()


def rope_table(text_len: int, hp: int, wp: int, axes_dim=(16, 56, 56), theta: float = 10000.0) -> torch.Tensor:
    """mmdit.py:865-911 — returns (S, D/2, 2) = (cos, sin) per token and rotation pair.
    positions: text tokens (0,0,0), image token (row r, col c) -> (0, r, c)."""
    S = text_len + hp * wp
    pos = torch.zeros((S, 3), dtype=torch.float32)
    rr = torch.arange(hp, dtype=torch.float32)[:, None].expand(hp, wp).reshape(-1)
    cc = torch.arange(wp, dtype=torch.float32)[None, :].expand(hp, wp).reshape(-1)
    pos[text_len:, 1] = rr
    pos[text_len:, 2] = cc
    outs = []
    for a, dim in enumerate(axes_dim):
        scale = torch.arange(0, dim, 2, dtype=torch.float32) / dim
        omega = 1.0 / (theta ** scale)
        outs.append(pos[:, a:a + 1] * omega[None, :])
    ang = torch.cat(outs, dim=-1)  # (S, D/2)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)


def rope_apply(x: torch.Tensor, table: torch.Tensor, dt=None) -> torch.Tensor:
    """mmdit.py:934-942 — x (B, H, S, D) viewed as adjacent pairs; out = (x0 c - x1 s, x0 s + x1 c) in fp32."""
    B, H, S, D = x.shape
    xp = x.reshape(B, H, S, D // 2, 2)
    c = table[None, None, :, :, 0]
    s = table[None, None, :, :, 1]
    out = torch.stack([xp[..., 0] * c - xp[..., 1] * s, xp[..., 0] * s + xp[..., 1] * c], dim=-1)
    return _r(out.reshape(B, H, S, D), dt)


def sdpa(q, k, v, scale):
    """mx.fast.scaled_dot_product_attention: softmax(scale q k^T) v, fp32 softmax, no mask (App. A.3)"""
    return torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1) @ v


class MMDiTRef:
    """Mirror of reference class MMDiT (mmdit.py:22-266)."""

    def __init__(self, cfg: RefMMDiTConfig, params: Dict[str, torch.Tensor], act_dtype: Optional[torch.dtype] = None):
        self.cfg = cfg
        self.p = params
        self.dt = act_dtype
        self._mod: Dict[float, Dict[str, torch.Tensor]] = {}
        self._rope = None

    # ------------------------------------------------------------------ helpers
    def _gelu(self, x):
        if self.cfg.gelu_tanh:
            return torch.nn.functional.gelu(x, approximate="tanh")
        return gelu_erf(x)

    def W(self, name):
        return self.p[name]

    def _lin(self, x, prefix, bias=True):
        b = self.p.get(prefix + ".bias") if bias else None
        return linear(x, self.p[prefix + ".weight"], b, self.dt)

    def _mlp_embedder(self, x, prefix):
        """Sequential(Linear, SiLU, Linear) (mmdit.py:352-364, 367-392)"""
        h = self._lin(x, prefix + ".mlp.layers.0")
        h = _r(F.silu(h), self.dt)
        return self._lin(h, prefix + ".mlp.layers.2")

    # ------------------------------------------------------------------ modulation cache (mmdit.py:77-180)
    def block_names(self):
        c = self.cfg
        names = []
        for i in range(c.depth_multimodal):
            names.append(f"multimodal_transformer_blocks.{i}.image_transformer_block")
            names.append(f"multimodal_transformer_blocks.{i}.text_transformer_block")
        for i in range(c.depth_unified):
            names.append(f"unified_transformer_blocks.{i}.transformer_block")
        names.append("final_layer")
        return names

    def cache_modulation_params(self, pooled: torch.Tensor, timesteps: torch.Tensor):
        """For every timestep value: c = y_embedder(pooled) + t_embedder(t); per block Linear(SiLU(c))."""
        y = self._mlp_embedder(pooled.float(), "y_embedder")                      # (B, h)  mmdit.py:85
        B = pooled.shape[0]
        for t in timesteps.reshape(-1):
            key = float(t)
            temb = timestep_embedding(t.reshape(1).repeat(B), self.cfg)           # mmdit.py:94-96
            tvec = self._mlp_embedder(temb, "t_embedder")
            c = _r(y + tvec, self.dt)
            sc = _r(F.silu(c), self.dt)
            mods = {}
            for bn in self.block_names():
                mods[bn] = self._lin(sc, bn + ".adaLN_modulation.layers.1")      # mmdit.py:430-435
            self._mod[key] = mods

    # ------------------------------------------------------------------ transformer block pieces
    def _pre_sdpa(self, x, bn, key, n_mod):
        """TransformerBlock.pre_sdpa (mmdit.py:440-519): x (B, S, h) -> q, k, v (B, heads, S, d), m, mods"""
        c = self.cfg
        mod = self._mod[key][bn]
        mods = mod.chunk(n_mod, dim=-1)  # order: shift1, scale1, gate1, shift2, scale2, gate2 (mmdit.py:449-455)
        m = affine_transform(x, mods[0][:, None, :], mods[1][:, None, :], c.layer_norm_eps, self.dt)
        q = self._lin(m, bn + ".attn.q_proj")
        k = self._lin(m, bn + ".attn.k_proj", bias=False)   # k has no bias (mmdit.py:821; quirk Q3)
        v = self._lin(m, bn + ".attn.v_proj")
        B, S, _ = x.shape
        d = c.hidden_size // c.num_heads

        def heads(t):
            return t.reshape(B, S, c.num_heads, d).permute(0, 2, 1, 3)

        q, k, v = heads(q), heads(k), heads(v)
        if c.use_qk_norm:                                    # mmdit.py:487-488
            q = rms_norm(q, self.p[bn + ".qk_norm.q_norm.weight"], 1e-6, self.dt)
            k = rms_norm(k, self.p[bn + ".qk_norm.k_norm.weight"], 1e-6, self.dt)
        return q, k, v, m, mods

    def _post_sdpa(self, x, o, m, mods, bn, parallel_mlp):
        """TransformerBlock.post_sdpa (mmdit.py:521-548)"""
        c = self.cfg
        attn_out = self._lin(o, bn + ".attn.o_proj")
        if parallel_mlp:
            h1 = _r(self._gelu(self._lin(m, bn + ".mlp.fc1")), self.dt)
            mlp_out = linear(h1, self.p[bn + ".mlp.fc2.weight"], None, self.dt)   # fc2.bias zeroed (mmdit.py:742)
            return _r(x + mods[2][:, None, :] * _r(attn_out + mlp_out, self.dt), self.dt)
        x = _r(x + attn_out * mods[2][:, None, :], self.dt)
        m2 = affine_transform(x, mods[3][:, None, :], mods[4][:, None, :], c.layer_norm_eps, self.dt)
        h1 = _r(self._gelu(self._lin(m2, bn + ".mlp.fc1")), self.dt)
        mlp_out = self._lin(h1, bn + ".mlp.fc2")
        return _r(x + mods[5][:, None, :] * mlp_out, self.dt)

    # ------------------------------------------------------------------ forward (mmdit.py:188-266)
    def __call__(self, latent: torch.Tensor, text: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        """latent (B, H, W, 16) NHWC, text (B, T, 4096) [the reference's extra singleton axis dropped],
        timestep (B,) (all equal) -> (B, H, W, 16)"""
        c = self.cfg
        dt = self.dt
        B, H, Wd, Cl = latent.shape
        key = float(timestep.reshape(-1)[0])
        hp, wp = H // c.patch_size, Wd // c.patch_size
        d = c.hidden_size // c.num_heads

        txt = self._lin(_r(text.float(), dt), "context_embedder")                 # mmdit.py:195

        x = _r(latent.float(), dt)
        if c.patchify_via_reshape:                                                # FLUX (mmdit.py:292-302)
            p = c.patch_size
            rows = x.reshape(B, hp, p, wp, p, Cl).permute(0, 1, 3, 5, 2, 4).reshape(B, hp * wp, Cl * p * p)
            wconv = self.p["x_embedder.proj.weight"].float().reshape(c.hidden_size, -1)   # (h, 1, 1, 64)
            img = linear(rows, wconv, self.p["x_embedder.proj.bias"], dt)
        else:                                                                     # SD3 conv k2 s2 (mmdit.py:285-290)
            p = c.patch_size
            rows = x.reshape(B, hp, p, wp, p, Cl).permute(0, 1, 3, 2, 4, 5).reshape(B, hp * wp, p * p * Cl)
            wconv = self.p["x_embedder.proj.weight"].float().reshape(c.hidden_size, -1)   # (O, kh, kw, I) flattened
            img = linear(rows, wconv, self.p["x_embedder.proj.bias"], dt)
            # learned position table, centre crop (mmdit.py:334-349)
            mh = c.max_latent_resolution
            y0, x0 = (mh - hp) // 2, (mh - wp) // 2
            pos = self.p["x_pos_embedder.pos_embed.weight"].float().reshape(mh, mh, c.hidden_size)
            pos = pos[y0:y0 + hp, x0:x0 + wp].reshape(1, hp * wp, c.hidden_size)
            img = _r(img + pos, dt)

        T = txt.shape[1]
        rope = None
        if c.use_rope:                                                            # mmdit.py:208-215
            rope = rope_table(T, hp, wp, c.rope_axes_dim)
        scale = 1.0 / math.sqrt(d)

        for i in range(c.depth_multimodal):                                       # mmdit.py:568-675
            ibn = f"multimodal_transformer_blocks.{i}.image_transformer_block"
            tbn = f"multimodal_transformer_blocks.{i}.text_transformer_block"
            skip_text = (i == c.depth_multimodal - 1) and (c.depth_unified < 1)   # mmdit.py:62-66
            qi, ki, vi, mi, modi = self._pre_sdpa(img, ibn, key, 6)
            qt, kt, vt, mt, modt = self._pre_sdpa(txt, tbn, key, 2 if skip_text else 6)
            if c.depth_unified > 0:                                               # FLUX order [text, image] (:594-606)
                q, k, v = torch.cat([qt, qi], 2), torch.cat([kt, ki], 2), torch.cat([vt, vi], 2)
            else:                                                                 # SD3 order [image, text] (:608-625)
                q, k, v = torch.cat([qi, qt], 2), torch.cat([ki, kt], 2), torch.cat([vi, vt], 2)
            if rope is not None:
                q, k = rope_apply(q, rope, dt), rope_apply(k, rope, dt)
            o = _r(sdpa(q, k, v, scale), dt).permute(0, 2, 1, 3).reshape(B, -1, c.hidden_size)
            n_img = img.shape[1]
            if c.depth_unified > 0:
                ot, oi = o[:, :T], o[:, T:]
            else:
                oi, ot = o[:, :n_img], o[:, n_img:]
            img = self._post_sdpa(img, oi, mi, modi, ibn, False)
            if not skip_text:
                txt = self._post_sdpa(txt, ot, mt, modt, tbn, False)

        if c.depth_unified > 0:                                                   # mmdit.py:233-247, 693-751
            u = torch.cat([txt, img], dim=1)
            for i in range(c.depth_unified):
                bn = f"unified_transformer_blocks.{i}.transformer_block"
                par = c.parallel_mlp_for_unified_blocks
                q, k, v, m, mods = self._pre_sdpa(u, bn, key, 3 if par else 6)
                if rope is not None:
                    q, k = rope_apply(q, rope, dt), rope_apply(k, rope, dt)
                o = _r(sdpa(q, k, v, scale), dt).permute(0, 2, 1, 3).reshape(B, -1, c.hidden_size)
                u = self._post_sdpa(u, o, m, mods, bn, par)
            img = u[:, T:]

        # FinalLayer (mmdit.py:780-796)
        mod = self._mod[key]["final_layer"]
        shift, sc = mod.chunk(2, dim=-1)
        y = affine_transform(img, shift[:, None, :], sc[:, None, :], c.layer_norm_eps, dt)
        y = self._lin(y, "final_layer.linear")                                    # (B, N, p*p*C)

        p = c.patch_size
        if c.patchify_via_reshape:                                                # unpack (c, ph, pw) (mmdit.py:304-321)
            out = y.reshape(B, hp, wp, Cl, p, p).permute(0, 1, 4, 2, 5, 3).reshape(B, hp * p, wp * p, Cl)
        else:                                                                     # unpatchify (p, q, c) (mmdit.py:975-988)
            out = y.reshape(B, hp, wp, p, p, Cl).permute(0, 5, 1, 3, 2, 4).reshape(B, Cl, H, Wd).permute(0, 2, 3, 1)
        return out.contiguous()
