"""CPU: checkpoint key remapping — upstream BFL / Stability / LDM layouts -> engine parameter tree.

The upstream layouts are built here from their published structure (fused qkv, linear1 = [q|k|v|fc1], linear2 =
[o|fc2] with one bias, OIHW convs, LDM decoder names), independently of the converter, then converted back."""
import os

import pytest
import torch

from diffusionkit_b200 import model_io
from diffusionkit_b200.config import VAEDecoderConfig, tiny_flux_config, tiny_sd3_config
from diffusionkit_b200.weights import (init_params, mmdit_param_specs, vae_decoder_param_specs,
                                       vae_encoder_param_specs)


def _flux_upstream(params, cfg):
    """engine tree -> BFL FLUX key layout"""
    h = cfg.hidden_size
    sd = {}
    for i in range(cfg.depth_multimodal):
        for s_up, s_us in (("img", "image_transformer_block"), ("txt", "text_transformer_block")):
            b = f"multimodal_transformer_blocks.{i}.{s_us}"
            u = f"double_blocks.{i}.{s_up}"
            sd[f"{u}_attn.qkv.weight"] = torch.cat([params[f"{b}.attn.{n}_proj.weight"] for n in "qkv"], 0)
            kb = torch.full((h,), 0.123)            # upstream carries a k bias; the reference drops it (quirk Q3)
            sd[f"{u}_attn.qkv.bias"] = torch.cat([params[f"{b}.attn.q_proj.bias"], kb, params[f"{b}.attn.v_proj.bias"]])
            sd[f"{u}_attn.proj.weight"] = params[f"{b}.attn.o_proj.weight"]
            sd[f"{u}_attn.proj.bias"] = params[f"{b}.attn.o_proj.bias"]
            sd[f"{u}_attn.norm.query_norm.scale"] = params[f"{b}.qk_norm.q_norm.weight"]
            sd[f"{u}_attn.norm.key_norm.scale"] = params[f"{b}.qk_norm.k_norm.weight"]
            for idx, fc in (("0", "fc1"), ("2", "fc2")):
                sd[f"{u}_mlp.{idx}.weight"] = params[f"{b}.mlp.{fc}.weight"]
                sd[f"{u}_mlp.{idx}.bias"] = params[f"{b}.mlp.{fc}.bias"]
            sd[f"{u}_mod.lin.weight"] = params[f"{b}.adaLN_modulation.layers.1.weight"]
            sd[f"{u}_mod.lin.bias"] = params[f"{b}.adaLN_modulation.layers.1.bias"]
    for i in range(cfg.depth_unified):
        b = f"unified_transformer_blocks.{i}.transformer_block"
        u = f"single_blocks.{i}"
        sd[f"{u}.linear1.weight"] = torch.cat([params[f"{b}.attn.q_proj.weight"], params[f"{b}.attn.k_proj.weight"],
                                               params[f"{b}.attn.v_proj.weight"], params[f"{b}.mlp.fc1.weight"]], 0)
        sd[f"{u}.linear1.bias"] = torch.cat([params[f"{b}.attn.q_proj.bias"], torch.full((h,), -0.5),
                                             params[f"{b}.attn.v_proj.bias"], params[f"{b}.mlp.fc1.bias"]])
        sd[f"{u}.linear2.weight"] = torch.cat([params[f"{b}.attn.o_proj.weight"], params[f"{b}.mlp.fc2.weight"]], 1)
        sd[f"{u}.linear2.bias"] = params[f"{b}.attn.o_proj.bias"]
        sd[f"{u}.modulation.lin.weight"] = params[f"{b}.adaLN_modulation.layers.1.weight"]
        sd[f"{u}.modulation.lin.bias"] = params[f"{b}.adaLN_modulation.layers.1.bias"]
        sd[f"{u}.norm.query_norm.scale"] = params[f"{b}.qk_norm.q_norm.weight"]
        sd[f"{u}.norm.key_norm.scale"] = params[f"{b}.qk_norm.k_norm.weight"]
    sd["img_in.weight"] = params["x_embedder.proj.weight"].reshape(h, -1)
    sd["img_in.bias"] = params["x_embedder.proj.bias"]
    sd["txt_in.weight"], sd["txt_in.bias"] = params["context_embedder.weight"], params["context_embedder.bias"]
    for up, us in (("time_in", "t_embedder"), ("vector_in", "y_embedder")):
        for layer, idx in (("in_layer", 0), ("out_layer", 2)):
            for leaf in ("weight", "bias"):
                sd[f"{up}.{layer}.{leaf}"] = params[f"{us}.mlp.layers.{idx}.{leaf}"]
    for leaf in ("weight", "bias"):
        sd[f"final_layer.linear.{leaf}"] = params[f"final_layer.linear.{leaf}"]
        sd[f"final_layer.adaLN_modulation.1.{leaf}"] = params[f"final_layer.adaLN_modulation.layers.1.{leaf}"]
    sd["guidance_in.in_layer.weight"] = torch.zeros(4, 4)      # FLUX.1-dev only; ignored (quirk Q1)
    return sd


def test_flux_checkpoint_roundtrip():
    cfg = tiny_flux_config()
    specs = mmdit_param_specs(cfg)
    params = init_params(specs, seed=5, dtype=torch.float32)
    got = model_io.flux_checkpoint_to_params(_flux_upstream(params, cfg), cfg.hidden_size, cfg.mlp_ratio)
    model_io.check_against_specs(got, specs)
    assert not any(k.endswith("k_proj.bias") for k in got)
    for name, t in params.items():
        if name.endswith("transformer_block.mlp.fc2.bias") and name.startswith("unified"):
            # linear2 has ONE bias: the reference gives it to o_proj and to fc2 (zeroed at run time, mmdit.py:742)
            assert torch.equal(got[name], params[name.replace("mlp.fc2.bias", "attn.o_proj.bias")])
            continue
        assert torch.equal(got[name], t), name
    assert got["x_embedder.proj.weight"].shape == (cfg.hidden_size, 1, 1, 64)


def _sd3_upstream(params, cfg, prefix="model.diffusion_model."):
    sd = {}
    for i in range(cfg.depth_multimodal):
        for b_up, s_us in (("x_block", "image_transformer_block"), ("context_block", "text_transformer_block")):
            b = f"multimodal_transformer_blocks.{i}.{s_us}"
            u = f"{prefix}joint_blocks.{i}.{b_up}"
            sd[f"{u}.attn.qkv.weight"] = torch.cat([params[f"{b}.attn.{n}_proj.weight"] for n in "qkv"], 0)
            sd[f"{u}.attn.qkv.bias"] = torch.cat([params[f"{b}.attn.q_proj.bias"], torch.zeros(cfg.hidden_size),
                                                  params[f"{b}.attn.v_proj.bias"]])
            for leaf in ("weight", "bias"):
                sd[f"{u}.adaLN_modulation.1.{leaf}"] = params[f"{b}.adaLN_modulation.layers.1.{leaf}"]
                if f"{b}.attn.o_proj.{leaf}" in params:
                    sd[f"{u}.attn.proj.{leaf}"] = params[f"{b}.attn.o_proj.{leaf}"]
                    sd[f"{u}.mlp.fc1.{leaf}"] = params[f"{b}.mlp.fc1.{leaf}"]
                    sd[f"{u}.mlp.fc2.{leaf}"] = params[f"{b}.mlp.fc2.{leaf}"]
    sd[prefix + "pos_embed"] = params["x_pos_embedder.pos_embed.weight"][None]
    sd[prefix + "x_embedder.proj.weight"] = params["x_embedder.proj.weight"].permute(0, 3, 1, 2).contiguous()   # OIHW
    sd[prefix + "x_embedder.proj.bias"] = params["x_embedder.proj.bias"]
    for leaf in ("weight", "bias"):
        sd[prefix + f"context_embedder.{leaf}"] = params[f"context_embedder.{leaf}"]
        sd[prefix + f"final_layer.linear.{leaf}"] = params[f"final_layer.linear.{leaf}"]
        sd[prefix + f"final_layer.adaLN_modulation.1.{leaf}"] = params[f"final_layer.adaLN_modulation.layers.1.{leaf}"]
        for emb in ("y_embedder", "t_embedder"):
            for idx in (0, 2):
                sd[prefix + f"{emb}.mlp.{idx}.{leaf}"] = params[f"{emb}.mlp.layers.{idx}.{leaf}"]
    sd["first_stage_model.decoder.conv_in.bias"] = torch.zeros(3)          # VAE tensors in the same file are skipped
    return sd


def test_sd3_checkpoint_roundtrip():
    cfg = tiny_sd3_config()
    specs = mmdit_param_specs(cfg)
    params = init_params(specs, seed=6, dtype=torch.float32)
    got = model_io.sd3_checkpoint_to_params(_sd3_upstream(params, cfg))
    model_io.check_against_specs(got, specs)
    for name, t in params.items():
        assert torch.equal(got[name], t), name
    assert got["x_embedder.proj.weight"].shape == (cfg.hidden_size, 2, 2, 16)      # (O, kh, kw, I)


def _vae_upstream(params, prefix="first_stage_model.decoder."):
    def oihw(w):
        return w.permute(0, 3, 1, 2).contiguous()

    sd = {}
    for name, t in params.items():
        leaf = name.rsplit(".", 1)[1]
        k = name
        k = k.replace("conv_norm_out.", "norm_out.")
        k = k.replace("mid_blocks.0.", "mid.block_1.").replace("mid_blocks.2.", "mid.block_2.")
        if k.startswith("mid_blocks.1."):
            k = k.replace("mid_blocks.1.", "mid.attn_1.").replace("group_norm", "norm").replace("query_proj", "q")
            k = k.replace("key_proj", "k").replace("value_proj", "v").replace("out_proj", "proj_out")
            v = t[:, :, None, None] if (leaf == "weight" and t.dim() == 2) else t
        else:
            k = k.replace("up_blocks.", "up.").replace(".resnets.", ".block.").replace(".conv_shortcut.", ".nin_shortcut.")
            k = k.replace(".upsample.", ".upsample.conv.")
            k = k.replace("down_blocks.", "down.").replace(".downsample.", ".downsample.conv.")
            if leaf == "weight" and t.dim() == 4:
                v = oihw(t)
            elif leaf == "weight" and "nin_shortcut" in k:
                v = t[:, :, None, None]
            else:
                v = t
        sd[prefix + k] = v
    return sd


def test_vae_decoder_checkpoint_roundtrip(tmp_path):
    specs = vae_decoder_param_specs(VAEDecoderConfig())
    params = init_params(specs, seed=7, dtype=torch.float32)
    up = _vae_upstream(params)
    assert "first_stage_model.decoder.up.3.upsample.conv.weight" in up and up[
        "first_stage_model.decoder.mid.attn_1.q.weight"].dim() == 4
    # through an actual .safetensors file
    from safetensors.torch import save_file

    path = os.path.join(tmp_path, "vae.safetensors")
    save_file({k: v.contiguous() for k, v in up.items()}, path)
    got = model_io.vae_decoder_checkpoint_to_params(model_io.load_safetensors(path))
    model_io.check_against_specs(got, specs)
    for name, t in params.items():
        assert torch.equal(got[name], t), name


def test_vae_encoder_checkpoint_roundtrip():
    from diffusionkit_b200.config import VAEEncoderConfig

    specs = vae_encoder_param_specs(VAEEncoderConfig())
    params = init_params(specs, seed=8, dtype=torch.float32)
    up = _vae_upstream(params, prefix="first_stage_model.encoder.")
    assert "first_stage_model.encoder.down.2.downsample.conv.weight" in up
    assert "first_stage_model.encoder.down.3.downsample.conv.weight" not in up
    assert up["first_stage_model.encoder.down.1.block.0.nin_shortcut.weight"].shape == (256, 128, 1, 1)
    up["first_stage_model.decoder.conv_in.bias"] = torch.zeros(3)          # the decoder half of the file is skipped
    got = model_io.vae_encoder_checkpoint_to_params(up)
    model_io.check_against_specs(got, specs)
    for name, t in params.items():
        assert torch.equal(got[name], t), name
    assert got["conv_out.weight"].shape == (32, 3, 3, 512)


def test_unknown_keys_and_shape_mismatch_are_reported():
    with pytest.raises(KeyError):
        model_io.flux_checkpoint_to_params({"double_blocks.0.img_attn.mystery.weight": torch.zeros(1)})
    cfg = tiny_sd3_config()
    specs = mmdit_param_specs(cfg)
    params = init_params(specs, seed=6, dtype=torch.float32)
    params.pop("context_embedder.bias")
    with pytest.raises(ValueError):
        model_io.check_against_specs(params, specs)
