"""Generates the committed fixtures under tests/golden/.

  schedule_kats.json : known-answer values of the reference's closed-form schedule / noise code
                       (mlx/sampler.py:28-35, mlx/__init__.py:553-571), evaluated here in float64 straight from
                       the formulas — independent of oracle/ — so they pin oracle/sampler_ref.py and the product's
                       get_sigmas/get_noise.
  tiny_*.npz         : outputs of the fp32 oracle (oracle/*.py) on tiny configs with the deterministic synthetic
                       weights (diffusionkit_b200/weights.py).  The reference itself cannot run here (MLX is
                       Apple-only, not installed), so these are oracle-generated regression vectors, not reference
                       outputs ("parity unpinned", see oracle/mmdit_ref.py).

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def sigma_closed_form(t, shift):
    t = t / 1000.0
    return t if shift == 1.0 else shift * t / (1 + (shift - 1) * t)


def schedule_kats():
    kats = {}
    # FLUX: table t=0..1000 -> sigma_min = 0, sigma_max = 1; num_steps + 1 points, no extra 0
    for n, shift in [(4, 1.0), (50, 1.0), (4, 3.0)]:
        ts = np.linspace(1000.0, 0.0, n + 1)
        kats[f"flux_n{n}_shift{shift}"] = {"sigmas": [sigma_closed_form(t, shift) for t in ts],
                                           "timesteps": [1000 * sigma_closed_form(t, shift) for t in ts]}
    # SD3: table t=1..1000 -> start = 1000*sigma(1000), end = 1000*sigma(1); linspace(num_steps) then append 0
    for n, shift in [(3, 3.0), (50, 3.0), (2, 1.0)]:
        start = 1000 * sigma_closed_form(1000.0, shift)
        end = 1000 * sigma_closed_form(1.0, shift)
        ts = np.linspace(start, end, n)
        sig = [sigma_closed_form(t, shift) for t in ts] + [0.0]
        kats[f"sd3_n{n}_shift{shift}"] = {"sigmas": sig, "timesteps": [1000 * s for s in sig]}
    # get_noise: numpy global RNG, randn(1, 16, H, W), NCHW draw order (mlx/__init__.py:553-557)
    np.random.seed(0)
    nz = np.random.randn(1, 16, 4, 4)
    kats["noise_seed0_4x4_first_nchw"] = nz.reshape(-1)[:8].tolist()
    kats["noise_seed0_4x4_nhwc_0_0_0_c"] = nz[0, :, 0, 0].tolist()
    return kats


def tiny_vectors():
    from diffusionkit_b200.config import tiny_flux_config, tiny_sd3_config, VAEDecoderConfig
    from diffusionkit_b200.weights import init_params, mmdit_param_specs, vae_decoder_param_specs
    from tests.oracle_bridge import ref_config
    from oracle.mmdit_ref import MMDiTRef
    from oracle.vae_ref import VAEDecoderRef, decode_latents_to_image

    out = {}
    for name, cfg, T, lat in [("flux", tiny_flux_config(), 16, (8, 12)), ("sd3", tiny_sd3_config(), 24, (8, 8))]:
        params = init_params(mmdit_param_specs(cfg), seed=7, dtype=torch.float32)
        ref = MMDiTRef(ref_config(cfg), params, act_dtype=None)
        g = torch.Generator().manual_seed(3)
        B = 2
        latent = torch.randn((B, lat[0], lat[1], 16), generator=g)
        text = torch.randn((B, T, cfg.token_level_text_embed_dim), generator=g)
        pooled = torch.randn((B, cfg.pooled_text_embed_dim), generator=g)
        t = torch.tensor([752.0])
        ref.cache_modulation_params(pooled, t)
        y = ref(latent, text, t.repeat(B))
        np.savez_compressed(os.path.join(HERE, f"tiny_{name}_mmdit.npz"), latent=latent.numpy(), text=text.numpy(),
                            pooled=pooled.numpy(), timestep=t.numpy(), out=y.numpy())
        out[name] = float(y.abs().mean())
    vcfg = VAEDecoderConfig()
    vp = init_params(vae_decoder_param_specs(vcfg), seed=8, dtype=torch.float32)
    g = torch.Generator().manual_seed(4)
    z = torch.randn((1, 8, 8, 16), generator=g)
    img = decode_latents_to_image(VAEDecoderRef(vp), z)
    np.savez_compressed(os.path.join(HERE, "tiny_vae_decode.npz"), latent=z.numpy(), image=img.numpy())
    out["vae"] = float(img.mean())
    # VAE encoder + img2img posterior sample, reduced widths to keep the fixture small
    from diffusionkit_b200.config import VAEEncoderConfig
    from diffusionkit_b200.weights import vae_encoder_param_specs
    from oracle.vae_ref import VAEEncoderRef, encode_image_to_latents, read_image_array
    from oracle.sampler_ref import get_noise

    ecfg = VAEEncoderConfig(block_out_channels=(32, 64, 64, 64), layers_per_block=1)
    ep = init_params(vae_encoder_param_specs(ecfg), seed=9, dtype=torch.float32)
    rng = np.random.RandomState(5)
    img_u8 = rng.randint(0, 256, (32, 48, 3), dtype=np.uint8)
    enc = VAEEncoderRef(ep, None, ecfg.block_out_channels, ecfg.layers_per_block)
    x = read_image_array(torch.from_numpy(img_u8))
    hidden = enc(x)
    z = encode_image_to_latents(enc, x, get_noise(3, 4, 6))
    np.savez_compressed(os.path.join(HERE, "tiny_vae_encode.npz"), image_u8=img_u8, hidden=hidden.numpy(),
                        latent_seed3=z.numpy())
    out["vae_encode"] = float(hidden.abs().mean())
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "schedule_kats.json"), "w") as f:
        json.dump(schedule_kats(), f, indent=1)
    print(tiny_vectors())
