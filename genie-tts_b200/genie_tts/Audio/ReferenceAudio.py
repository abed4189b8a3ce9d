"""Prompt feature container (reference: src/genie_tts/Audio/ReferenceAudio.py:13-76).

Same fields and LRU (env ``Max_Cached_Reference_Audio``, default 10).  The HuBERT
and speaker-verification networks are external assets outside this build's scope
(SURVEY.md §2 row 16): their outputs enter through ``set_feature_extractors`` or
``ReferenceAudio.from_features``.  What is new: ``device_prompt(model)`` turns the
features into a device-resident ``B200Prompt`` ONCE (prompt VQ, ref_enc /
prompt_encoder, ge-only conditioning), instead of the reference recomputing
them inside every encoder / vocoder call."""
from __future__ import annotations

import os
import threading
from typing import Callable, Dict, Optional

import numpy as np

from ..GetPhonesAndBert import get_phones_and_bert
from ..Utils.Utils import LRUCacheDict
from .Audio import load_audio, resample

_hubert: Optional[Callable[[np.ndarray], np.ndarray]] = None     # audio_16k [1,N] -> ssl_content [1,768,T]
_sv: Optional[Callable[[np.ndarray], np.ndarray]] = None         # audio_16k [1,N] -> sv_emb [1,20480]


def set_feature_extractors(hubert=None, sv=None) -> None:
    global _hubert, _sv
    _hubert, _sv = hubert, sv


class ReferenceAudio:
    _prompt_cache: Dict[str, "ReferenceAudio"] = LRUCacheDict(
        capacity=int(os.getenv("Max_Cached_Reference_Audio", "10")),
        on_evict=lambda k, inst: inst._drop_device_prompts())        # an evicted reference frees its HBM
    _device_lock = threading.RLock()      # device prompts are built from request threads (server) and the worker

    def __new__(cls, prompt_wav: str, prompt_text: str, language: str):
        if prompt_wav in cls._prompt_cache:
            inst = cls._prompt_cache[prompt_wav]
            if inst.text != prompt_text:
                inst.set_text(prompt_text, language=language)
            return inst
        inst = super().__new__(cls)
        cls._prompt_cache[prompt_wav] = inst
        return inst

    def __init__(self, prompt_wav: str, prompt_text: str, language: str):
        if getattr(self, "_initialized", False):
            return
        self._device_prompts: Dict[int, object] = {}
        self.text = prompt_text
        self.phonemes_seq: Optional[np.ndarray] = None
        self.text_bert: Optional[np.ndarray] = None
        self.set_text(prompt_text, language=language)
        a32 = load_audio(prompt_wav, target_sampling_rate=32000)
        if a32 is None:
            raise FileNotFoundError(f"cannot read reference audio {prompt_wav}")
        a16 = resample(a32, 32000, 16000)
        self.audio_32k = a32[None].astype(np.float32)
        self.audio_16k = a16[None].astype(np.float32)
        if _hubert is None:
            raise RuntimeError("no HuBERT feature extractor registered (external asset chinese-hubert-base; "
                               "call genie_tts.Audio.ReferenceAudio.set_feature_extractors)")
        self.ssl_content = np.asarray(_hubert(self.audio_16k), dtype=np.float32)
        self.sv_emb: Optional[np.ndarray] = None
        self.global_emb: Optional[np.ndarray] = None
        self.global_emb_advanced: Optional[np.ndarray] = None
        self._initialized = True

    @classmethod
    def from_features(cls, key: str, phonemes_seq, text_bert, ssl_content, audio_32k, sv_emb=None,
                      text: str = "") -> "ReferenceAudio":
        """Build from boundary features directly (synthetic inputs, precomputed HuBERT/SV)."""
        inst = object.__new__(cls)
        inst._device_prompts = {}
        inst.text = text
        inst.phonemes_seq = np.asarray(phonemes_seq, dtype=np.int64).reshape(1, -1)
        inst.text_bert = np.asarray(text_bert, dtype=np.float32)
        inst.ssl_content = np.asarray(ssl_content, dtype=np.float32)
        inst.audio_32k = np.asarray(audio_32k, dtype=np.float32).reshape(1, -1)
        inst.audio_16k = None
        inst.sv_emb = None if sv_emb is None else np.asarray(sv_emb, dtype=np.float32)
        inst.global_emb = inst.global_emb_advanced = None
        inst._initialized = True
        cls._prompt_cache[key] = inst
        return inst

    def set_text(self, prompt_text: str, language: str) -> None:
        self.text = prompt_text
        self.phonemes_seq, self.text_bert = get_phones_and_bert(prompt_text, language=language)
        self._drop_device_prompts()

    def _drop_device_prompts(self) -> None:
        with self._device_lock:
            for p in getattr(self, "_device_prompts", {}).values():
                p.close()
            self._device_prompts = {}

    @classmethod
    def clear_cache(cls) -> None:
        for inst in list(cls._prompt_cache.values()):
            inst._drop_device_prompts()
        cls._prompt_cache.clear()

    # -- device side -----------------------------------------------------------
    def device_prompt(self, model):
        """B200Prompt for ``model`` (built once per (reference audio, character))."""
        with self._device_lock:
            return self._device_prompt_locked(model)

    def _device_prompt_locked(self, model):
        key = id(model)
        for k in [k for k, q in self._device_prompts.items() if q.closed]:   # models closed / evicted since
            del self._device_prompts[k]
        p = self._device_prompts.get(key)
        if p is None or p.closed or p.model is not model:
            sv = None
            if model.is_v2pp:
                if self.sv_emb is None:
                    if _sv is None:
                        raise RuntimeError("V2ProPlus needs a speaker-verification embedding: register an SV "
                                           "extractor or pass sv_emb to ReferenceAudio.from_features")
                    self.sv_emb = np.asarray(_sv(self.audio_16k), dtype=np.float32)
                sv = self.sv_emb
            p = model.make_prompt(self.phonemes_seq, self.text_bert, self.ssl_content, self.audio_32k, sv)
            self._device_prompts[key] = p
        return p

    def update_global_emb(self, prompt_encoder) -> None:
        """Reference API (:68-76): fills global_emb / global_emb_advanced once (V2ProPlus)."""
        if self.global_emb is not None or prompt_encoder is None:
            return
        _, ge, gea = self.device_prompt(prompt_encoder.model).read()
        self.global_emb = ge.reshape(1, -1, 1)
        self.global_emb_advanced = gea.reshape(1, -1, 1)
