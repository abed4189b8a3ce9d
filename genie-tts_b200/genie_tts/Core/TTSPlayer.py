"""Streaming session object (reference: src/genie_tts/Core/TTSPlayer.py:24-241).

Same public methods — start_session / feed / end_session / stop /
wait_for_tts_completion / wait_for_playback_done — and the same outputs (one
int16 PCM chunk per sentence to ``chunk_callback``, optional 32 kHz mono s16 wav,
optional playback when ``sounddevice`` is present).  What changes is inside the
worker: it drains every sentence that is already queued and synthesises them as
ONE ragged batch on the GPU (first sentence alone, to keep first-audio latency
at batch-1), emitting chunks in text order."""
from __future__ import annotations

import logging
import queue
import threading
import wave
from typing import Callable, List, Optional

import numpy as np

from ..ModelManager import model_manager
from ..Utils.Shared import context
from ..Utils.TextSplitter import TextSplitter
from ..Utils.Utils import clear_queue
from .Inference import tts_client

logger = logging.getLogger(__name__)
STREAM_END = object()
MAX_BATCH_SENTENCES = 64


class TTSPlayer:
    def __init__(self, sample_rate: int = 32000):
        self.sample_rate = sample_rate
        self._text_queue: "queue.Queue" = queue.Queue()
        self._audio_queue: "queue.Queue" = queue.Queue()
        self._api_lock = threading.Lock()
        self._tts_done_event = threading.Event()
        self._playback_done_event = threading.Event()
        self._tts_done_event.set()
        self._playback_done_event.set()
        self._worker: Optional[threading.Thread] = None
        self._playback_worker: Optional[threading.Thread] = None
        self._splitter = TextSplitter()
        self._play = False
        self._split = False
        self._save_path: Optional[str] = None
        self._chunk_callback: Optional[Callable[[Optional[bytes]], None]] = None
        self._session_audio_chunks: List[np.ndarray] = []
        self._first_of_session = True

    @staticmethod
    def _preprocess_for_playback(audio_float: np.ndarray) -> bytes:
        return (np.clip(audio_float.squeeze(), -1.0, 1.0) * 32767).astype(np.int16).tobytes()

    # -- worker -----------------------------------------------------------------
    def _emit(self, audio: Optional[np.ndarray]) -> None:
        if audio is None or len(audio) == 0:
            return
        if self._play:
            self._playback_done_event.clear()
            self._audio_queue.put(audio)
        if self._save_path:
            self._session_audio_chunks.append(audio)
        if self._chunk_callback:
            self._chunk_callback(self._preprocess_for_playback(audio))

    def _finish_session(self) -> None:
        if self._save_path and self._session_audio_chunks:
            self._save_session_audio()
        if self._chunk_callback:
            self._chunk_callback(None)
        self._tts_done_event.set()

    def _tts_worker_loop(self) -> None:
        while True:
            item = self._text_queue.get()
            if item is None:
                return
            if item is STREAM_END:
                self._finish_session()
                continue
            sentences = [item]
            saw_end = False
            # drain what is already queued into one batch (not across a session end)
            while not self._first_of_session and len(sentences) < MAX_BATCH_SENTENCES:
                try:
                    nxt = self._text_queue.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._text_queue.put(None)
                    break
                if nxt is STREAM_END:
                    saw_end = True
                    break
                sentences.append(nxt)
            self._first_of_session = False
            try:
                gsv = model_manager.get(context.current_speaker)
                if gsv is None or context.current_prompt_audio is None:
                    raise RuntimeError("no character / reference audio set")
                if len(sentences) == 1:
                    audio = tts_client.tts(
                        text=sentences[0], prompt_audio=context.current_prompt_audio, encoder=gsv.T2S_ENCODER,
                        first_stage_decoder=gsv.T2S_FIRST_STAGE_DECODER, stage_decoder=gsv.T2S_STAGE_DECODER,
                        vocoder=gsv.VITS, prompt_encoder=gsv.PROMPT_ENCODER, language=gsv.LANGUAGE)
                    self._emit(audio)
                else:
                    from ..GetPhonesAndBert import get_phones_and_bert
                    model = gsv.engine
                    prompt = context.current_prompt_audio.device_prompt(model)
                    feats = [get_phones_and_bert("。" + s, language=gsv.LANGUAGE) for s in sentences]
                    auds = tts_client.tts_batch(model, [prompt] * len(feats), [f[0] for f in feats],
                                                [f[1] for f in feats])
                    for a in auds:
                        self._emit(a)
            except Exception as e:                      # reference: log, keep the stream alive (:109-114)
                logger.error(f"TTS worker error: {e}", exc_info=True)
            if saw_end:
                self._finish_session()

    def _playback_worker_loop(self) -> None:
        try:
            import sounddevice as sd
        except Exception:
            sd = None
        while True:
            audio = self._audio_queue.get()
            if audio is None:
                return
            try:
                if sd is not None:
                    sd.play(audio, self.sample_rate, blocking=True)
            except Exception as e:
                logger.error(f"playback error: {e}")
            if self._audio_queue.empty():
                self._playback_done_event.set()

    def _save_session_audio(self) -> None:
        pcm = self._preprocess_for_playback(np.concatenate(self._session_audio_chunks))
        try:
            with wave.open(self._save_path, "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(self.sample_rate)
                w.writeframes(pcm)
        except Exception as e:
            logger.error(f"failed to save audio to {self._save_path}: {e}")
        self._session_audio_chunks = []

    # -- public API --------------------------------------------------------------
    def start_session(self, play: bool = False, split: bool = False, save_path: Optional[str] = None,
                      chunk_callback: Optional[Callable[[Optional[bytes]], None]] = None, sampling=None) -> None:
        """``sampling`` (extension): SamplingParams of this session; None = the reference's graph constants."""
        with self._api_lock:
            if self._worker is None or not self._worker.is_alive():
                self._worker = threading.Thread(target=self._tts_worker_loop, daemon=True)
                self._worker.start()
            if play and (self._playback_worker is None or not self._playback_worker.is_alive()):
                self._playback_worker = threading.Thread(target=self._playback_worker_loop, daemon=True)
                self._playback_worker.start()
            tts_client.stop_event.clear()
            if sampling is not None:
                tts_client.sampling = sampling
            else:
                from ..engine import SamplingParams
                tts_client.sampling = SamplingParams()
            clear_queue(self._text_queue)
            clear_queue(self._audio_queue)
            self._tts_done_event.clear()
            self._play, self._split, self._save_path = play, split, save_path
            self._chunk_callback = chunk_callback
            self._session_audio_chunks = []
            self._first_of_session = True

    def feed(self, text_chunk: str) -> None:
        with self._api_lock:
            if not text_chunk:
                return
            parts = self._splitter.split(text_chunk) if self._split else [text_chunk]
            for s in parts:
                self._text_queue.put(s)

    def end_session(self) -> None:
        with self._api_lock:
            self._text_queue.put(STREAM_END)

    def stop(self) -> None:
        with self._api_lock:
            tts_client.stop_event.set()
            clear_queue(self._text_queue)
            clear_queue(self._audio_queue)
            try:
                import sounddevice as sd
                sd.stop()
            except Exception:
                pass
            self._text_queue.put(STREAM_END)
            self._playback_done_event.set()

    def wait_for_tts_completion(self) -> None:
        self._tts_done_event.wait()

    def wait_for_playback_done(self) -> None:
        self._tts_done_event.wait()
        self._playback_done_event.wait()


tts_player: TTSPlayer = TTSPlayer()
