"""Tokenizers of the text path: the reference's CLIP byte-pair tokenizer
(python/src/diffusionkit/mlx/tokenizer.py:14-122, itself a reduced CLIPTokenizer) and its T5 wrapper (:125-160).

No vocabulary ships with this repository (no network): `load_tokenizer` / `load_t5_tokenizer` take local files — the
`vocab.json` + `merges.txt` of openai/clip-vit-large-patch14 (tokenizer_l) / laion CLIP-ViT-bigG (tokenizer_g), and a
local google/t5-v1_1-xxl tokenizer directory (or `spiece.model`).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import regex


class Tokenizer:
    """Lower-cases, collapses whitespace, splits with CLIP's pattern and applies the byte-pair merges in rank order.
    Like the reference it skips CLIPTokenizer's ftfy / html clean-up and byte-level fallback (tokenizer.py:99-101)."""

    _SPLIT = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""

    def __init__(self, bpe_ranks: Dict[Tuple[str, str], int], vocab: Dict[str, int], pad_with_eos: bool = False):
        self.bpe_ranks = bpe_ranks
        self.vocab = vocab
        self.pat = regex.compile(self._SPLIT, regex.IGNORECASE)
        self.pad_to_max_length = True
        self.max_length = 77
        self.pad_with_eos = pad_with_eos
        self._cache: Dict[str, List[str]] = {self.bos: [self.bos], self.eos: [self.eos]}

    bos = "<|startoftext|>"
    eos = "<|endoftext|>"

    @property
    def bos_token(self) -> int:
        return self.vocab[self.bos]

    @property
    def eos_token(self) -> int:
        return self.vocab[self.eos]

    def bpe(self, word: str) -> List[str]:
        """Greedy lowest-rank-first merging of one pre-token; the last symbol carries the `</w>` end-of-word mark."""
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        symbols = list(word[:-1]) + [word[-1] + "</w>"]
        while len(symbols) > 1:
            best_rank, best_pair = None, None
            for pair in zip(symbols, symbols[1:]):
                rank = self.bpe_ranks.get(pair)
                if rank is not None and (best_rank is None or rank < best_rank):
                    best_rank, best_pair = rank, pair
            if best_pair is None:
                break
            merged, i = [], 0
            while i < len(symbols):
                if i + 1 < len(symbols) and (symbols[i], symbols[i + 1]) == best_pair:
                    merged.append(symbols[i] + symbols[i + 1])
                    i += 2
                else:
                    merged.append(symbols[i])
                    i += 1
            symbols = merged
        self._cache[word] = symbols
        return symbols

    def tokenize(self, text, prepend_bos: bool = True, append_eos: bool = True):
        if isinstance(text, list):
            return [self.tokenize(t, prepend_bos, append_eos) for t in text]
        clean = regex.sub(r"\s+", " ", text.lower())
        pieces = [sym for word in regex.findall(self.pat, clean) for sym in self.bpe(word)]
        tokens = [self.vocab[s] for s in pieces]
        room = self.max_length - int(prepend_bos) - int(append_eos)
        if len(tokens) > room:
            tokens = tokens[:room]                                      # truncation (tokenizer.py:111-116)
        if prepend_bos:
            tokens = [self.bos_token] + tokens
        if append_eos:
            tokens.append(self.eos_token)
        return tokens


def load_tokenizer(vocab_file: str, merges_file: str, pad_with_eos: bool = False) -> Tokenizer:
    """reference model_io.py:941-959, with local files instead of hf_hub_download.  Like the reference it keeps the
    first 49152 - 256 - 2 merges after the header line."""
    with open(vocab_file, encoding="utf-8") as f:
        vocab = json.load(f)
    with open(merges_file, encoding="utf-8") as f:
        lines = f.read().strip().split("\n")[1: 49152 - 256 - 2 + 1]
    merges = [tuple(m.split()) for m in lines]
    return Tokenizer({pair: rank for rank, pair in enumerate(merges)}, vocab, pad_with_eos)


class T5Tokenizer:
    """reference tokenizer.py:125-160: a transformers T5 tokenizer (legacy=False) truncating to max_context_length.
    `source` is a local directory holding the google/t5-v1_1-xxl tokenizer files, or the path of its spiece.model."""

    def __init__(self, source: str, max_context_length: int, decoder_start_token_id: int = 0):
        self.max_length = max_context_length
        self._decoder_start_id = decoder_start_token_id
        if os.path.isdir(source) and not any(os.path.exists(os.path.join(source, f))
                                            for f in ("tokenizer.json", "tokenizer_config.json")):
            source = os.path.join(source, "spiece.model")          # a bare sentencepiece model in a directory
        if os.path.isdir(source):
            from transformers import AutoTokenizer

            self._tokenizer = AutoTokenizer.from_pretrained(source, legacy=False, model_max_length=self.max_length,
                                                            local_files_only=True)
        else:
            import sentencepiece as spm
            from transformers import T5Tokenizer as HFT5Tokenizer

            sp = spm.SentencePieceProcessor(model_file=source)
            vocab = [(sp.id_to_piece(i), sp.get_score(i)) for i in range(sp.get_piece_size())]
            self._tokenizer = HFT5Tokenizer(vocab=vocab, legacy=False, model_max_length=self.max_length)
        self.pad_to_max_length = True
        self.pad_with_eos = False

    @property
    def eos_id(self) -> int:
        return self._tokenizer.eos_token_id

    @property
    def decoder_start_id(self) -> int:
        return self._decoder_start_id

    def encode(self, s: str) -> np.ndarray:
        return np.asarray(self._tokenizer(s, return_tensors="np", return_attention_mask=False,
                                          max_length=self.max_length, truncation=True)["input_ids"])

    def decode(self, t: List[int], with_sep: bool = True) -> str:
        tokens = self._tokenizer.convert_ids_to_tokens(t)
        return "".join(tok.replace("▁", " " if with_sep else "") for tok in tokens)

    def tokenize(self, s: str) -> List[int]:
        return [int(t) for t in self.encode(s)[0]]


def load_t5_tokenizer(source: str, max_context_length: int = 256) -> T5Tokenizer:
    """reference model_io.py:962-964"""
    return T5Tokenizer(source, max_context_length)
