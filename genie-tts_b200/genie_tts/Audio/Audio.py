"""Reference-audio loading (reference: src/genie_tts/Audio/Audio.py:19-51): mono,
resample to the target rate, append 0.3 s of silence.  soundfile/soxr are optional
here; the stdlib ``wave`` module plus scipy's polyphase resampler stand in."""
import logging
import os
import wave
from typing import Optional

import numpy as np

logger = logging.getLogger(__name__)
MIN_DURATION_S, MAX_DURATION_S = 3, 10
SILENCE_TO_APPEND_S = 0.3
TARGET_SAMPLING_RATE = 16000


def _read(path: str):
    try:
        import soundfile as sf
        return sf.read(path, dtype="float32")
    except ImportError:
        with wave.open(path, "rb") as w:
            n, sr, ch, sw = w.getnframes(), w.getframerate(), w.getnchannels(), w.getsampwidth()
            raw = w.readframes(n)
        dt = {1: np.uint8, 2: np.int16, 4: np.int32}[sw]
        a = np.frombuffer(raw, dtype=dt).astype(np.float32)
        a = (a - 128.0) / 128.0 if sw == 1 else a / float(2 ** (8 * sw - 1))
        return a.reshape(-1, ch) if ch > 1 else a, sr


def resample(wav: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    if sr_in == sr_out:
        return wav
    try:
        import soxr
        return soxr.resample(wav, sr_in, sr_out, quality="hq")
    except ImportError:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(sr_in, sr_out)
        return resample_poly(wav, sr_out // g, sr_in // g).astype(np.float32)


def load_audio(audio_path: str, target_sampling_rate: int = TARGET_SAMPLING_RATE) -> Optional[np.ndarray]:
    try:
        wav, sr = _read(audio_path)
        if wav.ndim > 1:
            wav = wav.mean(axis=1)
        wav = resample(np.asarray(wav, dtype=np.float32), sr, target_sampling_rate)
    except Exception as e:
        logger.error(f"Failed to load reference audio: {audio_path}. Error: {e}")
        return None
    dur = len(wav) / target_sampling_rate
    if not (MIN_DURATION_S <= dur <= MAX_DURATION_S):
        logger.warning(f"The reference audio '{os.path.basename(audio_path)}' has a duration of {dur:.2f} seconds, "
                       f"which is outside the recommended range of {MIN_DURATION_S} to {MAX_DURATION_S} seconds!")
    pad = np.zeros(int(SILENCE_TO_APPEND_S * target_sampling_rate), dtype=np.float32)
    return np.concatenate([wav.astype(np.float32), pad])
