"""CPU: the text-path oracle (oracle/text_ref.py) pinned against transformers' own CLIP / T5 encoder implementations on
random small models (an independent implementation of the same published architectures the reference mirrors), the
checkpoint key mappings, the relative-position buckets, and the CLIP byte-pair tokenizer against transformers'
CLIPTokenizer on a synthetic vocabulary."""
import json
import os

import numpy as np
import pytest
import torch

from diffusionkit_b200 import model_io
from diffusionkit_b200.config import tiny_clip_config, tiny_t5_config
from diffusionkit_b200.text_encoders import clip_param_specs, relative_position_bucket, t5_param_specs
from diffusionkit_b200.tokenizer import Tokenizer, load_tokenizer
from oracle import text_ref as tr


def _hf_clip(cfg, act):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection

    torch.manual_seed(0)
    hc = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.model_dims, intermediate_size=4 * cfg.model_dims,
                        projection_dim=cfg.projection_dim, num_hidden_layers=cfg.num_layers,
                        num_attention_heads=cfg.num_heads, max_position_embeddings=cfg.max_length, hidden_act=act,
                        eos_token_id=2)        # legacy eos id -> pooled = argmax token, like the reference (clip.py:94)
    return CLIPTextModelWithProjection(hc).eval()


@pytest.mark.parametrize("act", ["quick_gelu", "gelu"])
def test_clip_oracle_matches_transformers(act):
    cfg = tiny_clip_config(projection=True, act=act)
    hf = _hf_clip(cfg, act)
    params = model_io.clip_checkpoint_to_params(dict(hf.state_dict()))
    model_io.check_against_specs(params, clip_param_specs(cfg))
    tokens = torch.randint(1, cfg.vocab_size - 1, (2, 20))
    tokens[0, 7] = cfg.vocab_size - 1           # "EOS" = largest id, at different places
    tokens[1, 19] = cfg.vocab_size - 1
    with torch.no_grad():
        want = hf(input_ids=tokens, output_hidden_states=True)
    pooled, last, hidden = tr.CLIPTextModelRef(params, cfg.num_layers, cfg.num_heads, act)(tokens)
    assert torch.allclose(last, want.last_hidden_state, atol=2e-5, rtol=1e-4)
    assert torch.allclose(pooled, want.text_embeds, atol=2e-5, rtol=1e-4)
    # hidden_states[-2] of the reference = output of the second-to-last layer = HF hidden_states[-2]
    assert torch.allclose(hidden[-2], want.hidden_states[-2], atol=2e-5, rtol=1e-4)


def test_t5_oracle_matches_transformers():
    from transformers import T5Config, T5EncoderModel

    cfg = tiny_t5_config()
    torch.manual_seed(1)
    hc = T5Config(vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff,
                  num_layers=cfg.num_layers, num_heads=cfg.num_heads, feed_forward_proj="gated-gelu",
                  relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0,
                  layer_norm_epsilon=1e-6)
    hf = T5EncoderModel(hc).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():       # HF initialises the norms to 1 and the bias table tiny: randomise
            if "layer_norm" in n:
                p.copy_(1 + 0.1 * torch.randn_like(p))
            if "relative_attention_bias" in n:
                p.copy_(torch.randn_like(p))
    params = model_io.t5_checkpoint_to_params(dict(hf.state_dict()))
    model_io.check_against_specs(params, t5_param_specs(cfg))
    tokens = torch.randint(0, cfg.vocab_size, (2, 150))     # > 128 apart: exercises the saturated buckets
    with torch.no_grad():
        want = hf(input_ids=tokens).last_hidden_state
    ref = tr.T5EncoderRef(params, cfg.num_layers, cfg.num_heads)
    got = ref(tokens)
    # HF's T5 v1.1 uses the tanh GELU ("gelu_new"); the reference uses nn.gelu (erf) — t5.py:187-188.  Compare with the
    # activation swapped to erf so the rest of the block is pinned exactly.
    import torch.nn as nn

    for blk in hf.encoder.block:
        blk.layer[1].DenseReluDense.act = nn.GELU()
    with torch.no_grad():
        want_erf = hf(input_ids=tokens).last_hidden_state
    assert torch.allclose(got, want_erf, atol=5e-5, rtol=1e-4)
    assert not torch.allclose(want, want_erf, atol=1e-6)     # the two activations really differ


def test_relative_position_buckets():
    rel = np.arange(-600, 601)
    got = relative_position_bucket(rel)
    want = tr.relative_position_bucket(torch.from_numpy(rel)).numpy()
    assert np.array_equal(got, want)
    from transformers.models.t5.modeling_t5 import T5Attention

    hf = T5Attention._relative_position_bucket(torch.from_numpy(rel), bidirectional=True, num_buckets=32,
                                               max_distance=128).numpy()
    # the reference's op order (log(n / 8) * scale in fp32) can differ from HF's at exact bucket edges only
    assert np.mean(got != hf) < 0.01 and np.all(np.abs(got - hf) <= 1)
    assert got[600] == 0 and got[601] == 17 and got[599] == 1 and got[0] == 15 and got[-1] == 31
    assert got.min() == 0 and got.max() == 31


_WORDS = ["a", "photo", "of", "cat", "cats", "the", "astronaut", "riding", "horse", "on", "mars", "!", "!!", ",", "42"]


def _synthetic_clip_vocab(tmp_path):
    """A small but real BPE vocabulary: characters, characters with </w>, and merges learnt greedily from _WORDS."""
    chars = sorted({c for w in _WORDS for c in w})
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    merges = []
    corpus = [list(w[:-1]) + [w[-1] + "</w>"] for w in _WORDS]
    for _ in range(60):
        counts = {}
        for sym in corpus:
            for p in zip(sym, sym[1:]):
                counts[p] = counts.get(p, 0) + 1
        if not counts:
            break
        best = max(sorted(counts), key=lambda p: counts[p])
        merges.append(best)
        if best[0] + best[1] not in vocab:
            vocab[best[0] + best[1]] = len(vocab)
        new = []
        for sym in corpus:
            out, i = [], 0
            while i < len(sym):
                if i + 1 < len(sym) and (sym[i], sym[i + 1]) == best:
                    out.append(sym[i] + sym[i + 1])
                    i += 2
                else:
                    out.append(sym[i])
                    i += 1
            new.append(out)
        corpus = new
    vocab["<|startoftext|>"] = len(vocab)
    vocab["<|endoftext|>"] = len(vocab)
    vf, mf = str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt")
    json.dump(vocab, open(vf, "w"))
    with open(mf, "w") as f:
        f.write("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    return vf, mf, vocab


def test_clip_tokenizer_matches_transformers(tmp_path):
    vf, mf, vocab = _synthetic_clip_vocab(tmp_path)
    tok = load_tokenizer(vf, mf, pad_with_eos=True)
    assert isinstance(tok, Tokenizer) and tok.eos_token == vocab["<|endoftext|>"]
    from transformers import CLIPTokenizer

    hf = CLIPTokenizer(vf, mf)
    for text in ["a photo of a cat", "The  astronaut riding a horse on Mars!!", "cats, cats , 42 cats!", "a",
                 "photo of the horse!"]:
        assert tok.tokenize(text) == hf(text)["input_ids"], text
    long = " ".join(["cat"] * 200)
    ids = tok.tokenize(long)
    assert len(ids) == 77 and ids[0] == tok.bos_token and ids[-1] == tok.eos_token
    assert ids == hf(long, truncation=True, max_length=77)["input_ids"]
    # batching rule of the pipeline (reference _tokenize): first row padded to 77, EOS padding for tokenizer_l
    pair = tr.tokenize_pair(tok, "a cat", None)
    assert pair.shape == (2, 77) and int(pair[0, -1]) == tok.eos_token and int(pair[1, 1]) == tok.eos_token
