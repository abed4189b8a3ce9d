"""Per-request synthesis service behind the REST surface (BASELINE config 5).

The reference's server drives ONE global ``tts_player`` / ``context`` (src/genie_tts/Server.py:87-143): a second
concurrent ``/tts`` clears the first one's queues (``Core/TTSPlayer.py:185-186``).  Here every request is its own
object: its sentences are split and phonemised on the host (unchanged reference front end), each sentence becomes one
utterance on the least-loaded GPU replica's ``ContinuousBatcher`` (slot admission while other requests decode), and
the request's chunks come back in text order as int16 PCM @32 kHz — one chunk per sentence, as the reference streams.

Replicas: one per GPU, one full copy of the character's weights each (utterances are independent: no collective,
§8e).  The replica on the process-wide ``model_manager``'s device shares that manager, so characters loaded through
``genie.load_character`` before ``start_server`` are served without a second copy.
"""
from __future__ import annotations

import logging
import os
import threading
import wave
from concurrent.futures import Future
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np

from .Audio.ReferenceAudio import ReferenceAudio
from .GetPhonesAndBert import get_phones_and_bert
from .ModelManager import ModelManager, model_manager
from .Scheduler import ContinuousBatcher, ReplicaPool
from .Utils.TextSplitter import TextSplitter

logger = logging.getLogger(__name__)


def pcm16(audio_float: np.ndarray) -> bytes:
    """Reference chunk format (Core/TTSPlayer.py:50-53): clip, scale to int16, raw bytes."""
    return (np.clip(np.asarray(audio_float).squeeze(), -1.0, 1.0) * 32767).astype(np.int16).tobytes()


class RequestStream:
    """One /tts request: futures of its sentences in text order."""

    def __init__(self, futures: List[Future], save_path: Optional[str] = None, sample_rate: int = 32000):
        self.futures, self.save_path, self.sample_rate = futures, save_path, sample_rate
        self._saved: List[bytes] = []

    def chunks(self, timeout: Optional[float] = None) -> Iterator[bytes]:
        """PCM chunks in sentence order; a failed sentence is logged and skipped (reference: the worker logs and
        keeps the stream alive, TTSPlayer.py:109-114)."""
        for f in self.futures:
            try:
                audio = f.result(timeout=timeout)
            except Exception as e:
                logger.error(f"TTS request: sentence failed: {e}")
                continue
            if audio is None or len(audio) == 0:
                continue
            b = pcm16(audio)
            if self.save_path:
                self._saved.append(b)
            yield b
        self.finish()

    def finish(self) -> None:
        if self.save_path and self._saved:
            try:
                with wave.open(self.save_path, "wb") as w:      # reference: 32 kHz mono s16 (TTSPlayer.py:149-162)
                    w.setnchannels(1)
                    w.setsampwidth(2)
                    w.setframerate(self.sample_rate)
                    w.writeframes(b"".join(self._saved))
            except Exception as e:
                logger.error(f"failed to save audio to {self.save_path}: {e}")
            self._saved = []


class SynthesisService:
    def __init__(self, devices: Optional[Sequence[int]] = None, n_slots: int = 256, kv_capacity: int = 1024,
                 max_prompt_tokens: int = 512, max_steps: int = 500, sampling=None,
                 contexts_per_gpu: Optional[int] = None):
        from . import _native as N
        N.require_gpu()
        if devices is None:
            env = os.getenv("GENIE_DEVICES")
            devices = [int(x) for x in env.split(",")] if env else list(range(N.lib().genie_device_count()))
        self.devices = list(devices)
        # schedulers per GPU, each on its own execution context (own streams, slot pool and graphs; the weights are
        # shared).  Measured on one B200, 256 closed-loop clients: 1714 / 1743 / 1556 audio-s/s for 1 / 2 / 3 - no gain, default 1
        self.contexts_per_gpu = int(contexts_per_gpu or os.getenv("GENIE_CONTEXTS_PER_GPU", "1"))
        self.pool_cfg = dict(n_slots=n_slots, kv_capacity=kv_capacity, max_prompt_tokens=max_prompt_tokens,
                             max_steps=max_steps, sampling=sampling)
        self.managers: List[ModelManager] = [model_manager if d == model_manager.device else ModelManager(device=d)
                                             for d in self.devices]
        self._batchers: Dict[str, List[ContinuousBatcher]] = {}    # character -> one batcher per replica
        self._pools: Dict[str, ReplicaPool] = {}
        self._refs: Dict[str, ReferenceAudio] = {}
        self._paths: Dict[str, tuple] = {}
        self._lock = threading.RLock()
        self._splitter = TextSplitter()

    # ---- characters ------------------------------------------------------------------------
    def load_character(self, name: str, model_dir: str, language: str) -> None:
        with self._lock:
            self._drop_batchers(name)
            for mgr in self.managers:
                if not mgr.load_character(name, model_dir, language):
                    raise RuntimeError(f"failed to load '{model_dir}' on cuda:{mgr.device}")
            self._paths[name.lower()] = (model_dir, language)

    def unload_character(self, name: str) -> None:
        with self._lock:
            self._drop_batchers(name)                 # scheduler threads finish their work and stop first
            for mgr in self.managers:
                mgr.remove_character(name)
            self._paths.pop(name.lower(), None)

    def _drop_batchers(self, name: str) -> None:
        for b in self._batchers.pop(name.lower(), []):
            b.close()
        self._pools.pop(name.lower(), None)

    def set_reference(self, name: str, reference: ReferenceAudio) -> None:
        with self._lock:
            self._refs[name.lower()] = reference

    def has_reference(self, name: str) -> bool:
        return name.lower() in self._refs

    def is_loaded(self, name: str) -> bool:
        return all(m.has_character(name) for m in self.managers)

    def _pool(self, name: str) -> ReplicaPool:
        key = name.lower()
        with self._lock:
            if key not in self._pools:
                bs = []
                for mgr in self.managers:
                    gsv = mgr.get(name)
                    if gsv is None and model_manager.has_character(name):     # loaded through genie.load_character
                        path = model_manager.character_model_paths[key]
                        mgr.load_character(name, path, model_manager.character_to_language.get(key, "Japanese"))
                        gsv = mgr.get(name)
                    if gsv is None:
                        raise KeyError(f"character '{name}' is not loaded")
                    # the pool lives on its own context: the character's main handle stays free for the
                    # reference-facing batch-1 path (genie.tts) next to the server
                    for k in range(self.contexts_per_gpu):
                        ctx = gsv.engine.create_context()
                        bs.append(ContinuousBatcher(ctx, name=f"{key}@cuda{mgr.device}.{k}", **self.pool_cfg))
                self._batchers[key] = bs
                self._pools[key] = ReplicaPool(bs)
            return self._pools[key]

    # ---- requests ---------------------------------------------------------------------------
    def language_of(self, name: str) -> str:
        gsv = self.managers[0].get(name)
        if gsv is None:
            raise KeyError(f"character '{name}' is not loaded")
        return gsv.LANGUAGE

    def submit(self, name: str, text: str, split_sentence: bool = False, save_path: Optional[str] = None,
               sampling=None) -> RequestStream:
        """Queue one request; returns at once.  Nothing global is touched: any number of requests may be in
        flight, each gets exactly its own sentences back."""
        ref = self._refs.get(name.lower())
        if ref is None:
            raise KeyError("reference audio not set")
        pool = self._pool(name)
        language = self.language_of(name)
        sentences = self._splitter.split(text) if split_sentence else [text]
        futures: List[Future] = []
        for s in sentences:
            if not s:
                continue
            seq, bert = get_phones_and_bert("。" + s, language=language)          # Inference.py:27-28
            i = pool.pick()
            prompt = ref.device_prompt(self._batchers[name.lower()][i].model)
            bert_arg = bert if (bert is not None and np.any(bert)) else None
            futures.append(pool.schedulers[i].submit(prompt, seq, bert_arg, sampling=sampling))
        return RequestStream(futures, save_path)

    def stop_all(self) -> None:
        """The reference's /stop: drop everything queued or decoding."""
        with self._lock:
            for bs in self._batchers.values():
                for b in bs:
                    b.cancel_all()

    def stats(self) -> dict:
        out = {}
        for name, bs in self._batchers.items():
            out[name] = [dict(b.stats.summary(), device=b.model.device) for b in bs]
        return out

    def close(self) -> None:
        with self._lock:
            for name in list(self._batchers):
                self._drop_batchers(name)
