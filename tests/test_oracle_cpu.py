"""CPU: the oracle against the committed golden vectors / known-answer values."""
import json
import math
import os

import numpy as np
import torch

from diffusionkit_b200.config import VAEDecoderConfig, tiny_flux_config, tiny_sd3_config
from diffusionkit_b200.weights import init_params, mmdit_param_specs, vae_decoder_param_specs
from oracle import sampler_ref as sr
from oracle.mmdit_ref import MMDiTRef, rope_apply, rope_table, timestep_embedding
from oracle.vae_ref import VAEDecoderRef, decode_latents_to_image, to_uint8
from tests.oracle_bridge import ref_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "schedule_kats.json")))


def test_schedule_kats_survey_values():
    """the values SURVEY.md §8c (i) derives from the reference formulas"""
    assert KATS["flux_n4_shift1.0"]["sigmas"] == [1.0, 0.75, 0.5, 0.25, 0.0]
    s = KATS["sd3_n3_shift3.0"]["sigmas"]
    assert np.allclose(s, [1.0, 0.7511211, 0.00892857, 0.0], atol=1e-7)
    s50 = KATS["sd3_n50_shift3.0"]["sigmas"]
    assert np.allclose(s50[1:4], [0.9931244, 0.98605704, 0.97878975], atol=1e-7) and abs(s50[-2] - 0.00892857) < 1e-8
    assert np.allclose(KATS["noise_seed0_4x4_first_nchw"][:3], [1.76405235, 0.40015721, 0.97873798], atol=1e-8)


def test_oracle_sigmas_match_kats():
    for key, val in KATS.items():
        if not key.startswith(("flux_n", "sd3_n")):
            continue
        fam, n, shift = key.split("_")
        n, shift = int(n[1:]), float(shift[5:])
        sampler = sr.FluxSamplerRef(shift) if fam == "flux" else sr.ModelSamplingDiscreteFlowRef(shift)
        sig = sr.get_sigmas(sampler, n)
        assert len(sig) == n + 1
        assert np.allclose(sig.numpy(), np.array(val["sigmas"]), rtol=2e-6, atol=1e-7), key
        assert np.allclose((sig * 1000).numpy(), np.array(val["timesteps"]), rtol=2e-6, atol=1e-4), key


def test_oracle_noise_matches_kat():
    nz = sr.get_noise(0, 4, 4)
    assert nz.shape == (1, 4, 4, 16) and nz.dtype == torch.float32
    # NHWC[0,0,0,c] == NCHW draw c*16 (mlx/__init__.py:553-557)
    assert np.allclose(nz[0, 0, 0, :].numpy(), np.array(KATS["noise_seed0_4x4_nhwc_0_0_0_c"]), atol=1e-6)
    assert abs(float(nz[0, 0, 1, 0]) - KATS["noise_seed0_4x4_first_nchw"][1]) < 1e-6
    assert torch.allclose(sr.get_empty_latent(2, 2), torch.full((1, 2, 2, 16), 0.0609))


def test_timestep_rounding_quirk_q5():
    # bf16 has 8 significant bits: 750 -> 752
    t = (torch.tensor([750.0]).to(torch.bfloat16)).float()
    assert float(t) == 752.0
    cfg = ref_config(tiny_flux_config())
    emb = timestep_embedding(torch.tensor([752.0]), cfg)
    assert emb.shape == (1, 256)
    assert abs(float(emb[0, 0]) - math.cos(752.0)) < 2e-2   # cos first; bf16 argument/rounding
    assert torch.equal(emb, emb.to(torch.bfloat16).float())  # values live on the bf16 grid


def test_rope_properties():
    tab = rope_table(5, 3, 4, (16, 56, 56))
    assert tab.shape == (5 + 12, 64, 2)
    # text tokens and axis 0 are unrotated (position 0)
    assert torch.allclose(tab[:5, :, 0], torch.ones(5, 64)) and torch.allclose(tab[:5, :, 1], torch.zeros(5, 64))
    assert torch.allclose(tab[:, :8, 0], torch.ones(17, 8))
    # image token (r=2, c=3) is index 5 + 2*4 + 3; first pair of axis 1 rotates by r * theta^0 = 2 rad
    i = 5 + 2 * 4 + 3
    assert abs(float(tab[i, 8, 0]) - math.cos(2.0)) < 1e-6 and abs(float(tab[i, 8 + 28, 1]) - math.sin(3.0)) < 1e-6
    x = torch.randn(1, 2, 17, 128)
    y = rope_apply(x, tab)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-4)          # rotations preserve norm
    assert torch.allclose(y[:, :, :5], x[:, :, :5])


def _load(name):
    return np.load(os.path.join(GOLD, name))


def test_oracle_mmdit_matches_golden():
    for name, cfg in [("flux", tiny_flux_config()), ("sd3", tiny_sd3_config())]:
        g = _load(f"tiny_{name}_mmdit.npz")
        params = init_params(mmdit_param_specs(cfg), seed=7, dtype=torch.float32)
        ref = MMDiTRef(ref_config(cfg), params)
        t = torch.from_numpy(g["timestep"])
        ref.cache_modulation_params(torch.from_numpy(g["pooled"]), t)
        y = ref(torch.from_numpy(g["latent"]), torch.from_numpy(g["text"]), t.repeat(g["latent"].shape[0]))
        assert y.shape == g["latent"].shape
        err = float((y - torch.from_numpy(g["out"])).norm() / torch.from_numpy(g["out"]).norm())
        assert err < 1e-5, (name, err)


def test_oracle_batch_independence():
    """samples are independent (no cross-sample op): batch of 2 == two batch-1 calls — the property batch sharding
    across GPUs relies on"""
    cfg = tiny_flux_config()
    g = _load("tiny_flux_mmdit.npz")
    params = init_params(mmdit_param_specs(cfg), seed=7, dtype=torch.float32)
    t = torch.from_numpy(g["timestep"])
    outs = []
    for b in range(2):
        ref = MMDiTRef(ref_config(cfg), params)
        ref.cache_modulation_params(torch.from_numpy(g["pooled"][b:b + 1]), t)
        outs.append(ref(torch.from_numpy(g["latent"][b:b + 1]), torch.from_numpy(g["text"][b:b + 1]), t))
    y = torch.cat(outs)
    assert float((y - torch.from_numpy(g["out"])).abs().max()) < 1e-4


def test_oracle_vae_matches_golden():
    g = _load("tiny_vae_decode.npz")
    vp = init_params(vae_decoder_param_specs(VAEDecoderConfig()), seed=8, dtype=torch.float32)
    img = decode_latents_to_image(VAEDecoderRef(vp), torch.from_numpy(g["latent"]))
    assert img.shape == (1, 64, 64, 3) and float(img.min()) >= 0 and float(img.max()) <= 1
    assert float((img - torch.from_numpy(g["image"])).abs().max()) < 1e-4
    u8 = to_uint8(torch.tensor([0.5, 0.999, 1.0, 0.0]))
    assert u8.tolist() == [127, 254, 255, 0]                                   # truncation (mlx/__init__.py:526)


def test_sample_euler_cfg_and_update():
    """Euler/CFG algebra on a linear stand-in denoiser"""
    calls = []

    def fake_mmdit(xin, cond, t):
        calls.append((xin.shape[0], float(t[0])))
        scale = torch.arange(1, xin.shape[0] + 1, dtype=torch.float32).reshape(-1, 1, 1, 1)
        return xin * 0.1 * scale

    x0 = torch.randn(1, 2, 2, 16)
    sig = torch.tensor([1.0, 0.5, 0.0])
    out = sr.sample_euler(fake_mmdit, lambda p, ts: None, x0.clone(), sig, torch.zeros(2, 3, 4), torch.zeros(2, 5), 2.0,
                          None)
    # by hand: step i: den_k = x - 0.1*k*x*s ; den = den_2 + 2 (den_1 - den_2) = x - 0.1 x s (2 + 2(1-2)) = x
    #          => d = (x - x)/s... den = x (1 - 0.1 s (2 - 2)) = x  -> x unchanged
    assert torch.allclose(out, x0, atol=1e-6)
    assert calls == [(2, 1000.0), (2, 500.0)]
    out2 = sr.sample_euler(lambda xin, c, t: xin * 0.0 + 1.0, lambda p, ts: None, x0.clone(), sig, torch.zeros(1, 3, 4),
                           torch.zeros(1, 5), 0.0, None)
    # model output 1: den = x - s; d = 1; x += (s_next - s) -> x0 - 1
    assert torch.allclose(out2, x0 - 1.0, atol=1e-6)


def test_compute_psnr_definition():
    a = np.array([0.0, 1.0, 2.0, 4.0])
    b = a + 0.1
    # utils.py:70-82: 20 log10((peak + 1e-5) / (rmse + 1e-10))
    assert abs(sr.compute_psnr(a, b) - 20 * np.log10((4 + 1e-5) / (0.1 + 1e-10))) < 1e-9
    assert sr.process_out(torch.tensor([0.0]), "flux").item() == np.float32(0.1159)


def test_q4_oracle_layout_and_roundtrip():
    """MLX 4-bit layout: element j of a word sits in bits [4j, 4j+4); dequantised value = scale * q + bias"""
    import numpy as np

    from oracle import quant_ref as qr

    assert qr.unpack_q4(np.array([[0x76543210]], dtype=np.uint32)).tolist() == [[0, 1, 2, 3, 4, 5, 6, 7]]
    rng = np.random.RandomState(0)
    w = rng.randn(16, 256).astype(np.float32)
    wq, sc, bi = qr.quantize_q4(w)
    assert wq.shape == (16, 32) and sc.shape == bi.shape == (16, 4) and wq.dtype == np.uint32
    back = qr.dequantize_q4(wq, sc, bi)
    step = np.repeat(sc, 64, axis=1)
    assert np.all(np.abs(back - w) <= 0.5 * step + 1e-6)           # round-to-nearest inside [min, max]
    # group extremes are represented exactly (q = 0 and q = 15)
    g = w.reshape(16, 4, 64)
    assert np.allclose(back.reshape(16, 4, 64).min(-1), g.min(-1), atol=1e-6)
    assert np.allclose(back.reshape(16, 4, 64).max(-1), g.max(-1), atol=1e-5)
    # KAT: scales 0.5, bias -1, nibbles 0..7 -> -1, -0.5, ..., 2.5
    kat = qr.dequantize_q4(np.array([[0x76543210]], dtype=np.uint32), np.array([[0.5]], np.float32),
                           np.array([[-1.0]], np.float32), group_size=8)
    assert kat.tolist() == [[-1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0, 2.5]]


def test_oracle_vae_encoder_matches_golden():
    """encoder stack (stride-2 downsample with bottom/right pad), read_image scaling and the seeded posterior sample"""
    from diffusionkit_b200.config import VAEEncoderConfig
    from diffusionkit_b200.weights import init_params, vae_encoder_param_specs
    from oracle.sampler_ref import get_noise
    from oracle.vae_ref import VAEEncoderRef, conv3x3_s2, encode_image_to_latents, read_image_array

    g = _load("tiny_vae_encode.npz")
    ecfg = VAEEncoderConfig(block_out_channels=(32, 64, 64, 64), layers_per_block=1)
    ep = init_params(vae_encoder_param_specs(ecfg), seed=9, dtype=torch.float32)
    enc = VAEEncoderRef(ep, None, ecfg.block_out_channels, ecfg.layers_per_block)
    x = read_image_array(torch.from_numpy(g["image_u8"]))
    assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0 and x.shape == (1, 32, 48, 3)
    hidden = enc(x)
    assert hidden.shape == (1, 4, 6, 32)
    assert torch.allclose(hidden, torch.from_numpy(g["hidden"]), atol=1e-5, rtol=1e-4)
    z = encode_image_to_latents(enc, x, get_noise(3, 4, 6))
    assert torch.allclose(z, torch.from_numpy(g["latent_seed3"]), atol=1e-5, rtol=1e-4)
    # the downsample really pads bottom/right only: output (i, j) sees input rows 2i..2i+2, the last one zero-padded
    xs = torch.arange(16.0).reshape(1, 4, 4, 1)
    w = torch.zeros(1, 3, 3, 1)
    w[0, 2, 2, 0] = 1.0                                   # picks input (2i + 2, 2j + 2)
    y = conv3x3_s2(xs, w, torch.zeros(1))
    assert y.reshape(2, 2).tolist() == [[10.0, 0.0], [0.0, 0.0]]
