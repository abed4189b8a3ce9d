"""Seeded synthetic inputs at the boundary of the hot path: the
outputs of G2P / HuBERT / SV / RoBERTa, which are out of scope and unavailable."""
from __future__ import annotations

import numpy as np


def make_prompt_inputs(seed: int, Lr: int = 60, Ts: int = 264, n_audio: int = 169600, bert: bool = False,
                       v2pp: bool = False) -> dict:
    rng = np.random.default_rng(seed)
    d = {
        "ref_seq": rng.integers(0, 732, (1, Lr)).astype(np.int64),
        "ref_bert": (rng.standard_normal((Lr, 1024)).astype(np.float32) if bert
                     else np.zeros((Lr, 1024), np.float32)),
        "ssl_content": rng.standard_normal((1, 768, Ts)).astype(np.float32),
        "ref_audio": (0.1 * rng.standard_normal((1, n_audio))).astype(np.float32),
    }
    if v2pp:
        d["sv_emb"] = rng.standard_normal((1, 20480)).astype(np.float32)
    return d


def make_text_inputs(seed: int, Lt: int = 50, bert: bool = False) -> dict:
    rng = np.random.default_rng(seed)
    seq = rng.integers(0, 732, (1, Lt)).astype(np.int64)
    seq[0, 0] = 3          # '.' — the '。' prefix (reference Inference.py:27)
    return {
        "text_seq": seq,
        "text_bert": (rng.standard_normal((Lt, 1024)).astype(np.float32) if bert
                      else np.zeros((Lt, 1024), np.float32)),
    }


def make_zp_noise(seed: int, n_tokens: int) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((1, 192, 2 * n_tokens)).astype(np.float32)
