"""Public API (reference: src/genie_tts/Internal.py:41-398) — signatures frozen."""
from __future__ import annotations

import asyncio
import logging
import os
from os import PathLike
from pathlib import Path
from typing import AsyncIterator, Dict, Optional, Union

from .Audio.ReferenceAudio import ReferenceAudio
from .Core.TTSPlayer import tts_player
from .ModelManager import model_manager
from .Utils.Language import normalize_language
from .Utils.Shared import context

logging.basicConfig(level=logging.INFO, format="%(message)s")
logger = logging.getLogger(__name__)

SUPPORTED_AUDIO_EXTS = {".wav", ".flac", ".ogg", ".aiff", ".aif"}
_REQUIRED = ("t2s_encoder_fp32.bin", "t2s_encoder_fp32.onnx", "t2s_first_stage_decoder_fp32.onnx",
             "t2s_shared_fp16.bin", "t2s_stage_decoder_fp32.onnx", "vits_fp16.bin", "vits_fp32.onnx")
_reference_audios: Dict[str, Dict[str, str]] = {}


def check_onnx_model_dir(onnx_model_dir: Union[str, PathLike]) -> None:
    """FileNotFoundError unless the directory holds the seven base files (reference :41-91);
    prompt_encoder_fp32.onnx / prompt_encoder_fp16.bin additionally mark a V2ProPlus model."""
    d = Path(onnx_model_dir)
    if not d.is_dir():
        raise FileNotFoundError(f"The model directory '{onnx_model_dir}' does not exist or is not a directory.")
    missing = [f for f in _REQUIRED if not (d / f).is_file()]
    if missing:
        raise FileNotFoundError(
            f"\n\n[Genie Error] Invalid ONNX model directory: '{d}'\nMissing base files: {', '.join(missing)}\n"
            "A valid model folder holds: " + ", ".join(_REQUIRED) +
            " (+ prompt_encoder_fp16.bin, prompt_encoder_fp32.onnx for v2ProPlus)\n")


def load_character(character_name: str, onnx_model_dir: Union[str, PathLike], language: str) -> None:
    check_onnx_model_dir(onnx_model_dir)
    language = normalize_language(language)
    if language not in ("Japanese", "English", "Chinese", "Hybrid-Chinese-English"):
        raise ValueError("Unknown language")
    model_manager.load_character(character_name=character_name, model_dir=os.fspath(onnx_model_dir),
                                 language=language)


def unload_character(character_name: str) -> None:
    model_manager.remove_character(character_name=character_name)


def set_reference_audio(character_name: str, audio_path: Union[str, PathLike], audio_text: str,
                        language: Optional[str] = None) -> None:
    audio_path = os.fspath(audio_path)
    ext = os.path.splitext(audio_path)[1].lower()
    if ext not in SUPPORTED_AUDIO_EXTS:
        logger.error(f"Audio format '{ext}' is not supported. Only the following formats are supported: "
                     f"{SUPPORTED_AUDIO_EXTS}")
        return
    if language is None:
        gsv = model_manager.get(character_name)
        if gsv is None:
            raise ValueError("No language specified")
        language = gsv.LANGUAGE
    language = normalize_language(language)
    if language not in ("Japanese", "English", "Chinese"):
        raise ValueError("Unknown language")
    _reference_audios[character_name] = {"audio_path": audio_path, "audio_text": audio_text, "language": language}
    context.current_prompt_audio = ReferenceAudio(prompt_wav=audio_path, prompt_text=audio_text, language=language)


def set_reference_features(character_name: str, reference: ReferenceAudio) -> None:
    """Extension: register a ReferenceAudio built with ``ReferenceAudio.from_features`` (precomputed
    HuBERT / SV features), bypassing wav loading.  Not part of the reference API."""
    _reference_audios[character_name] = {"reference": reference}     # type: ignore[dict-item]
    context.current_prompt_audio = reference


def _activate(character_name: str) -> None:
    entry = _reference_audios[character_name]
    context.current_speaker = character_name
    if "reference" in entry:
        context.current_prompt_audio = entry["reference"]
    else:
        context.current_prompt_audio = ReferenceAudio(prompt_wav=entry["audio_path"], prompt_text=entry["audio_text"],
                                                      language=entry["language"])


def _prep_save_path(save_path) -> Optional[str]:
    if not save_path:
        return None
    save_path = os.fspath(save_path)
    parent = os.path.dirname(save_path)
    if parent:
        os.makedirs(parent, exist_ok=True)
    return save_path


def _sampling(top_k, top_p, temperature, repetition_penalty, seed):
    """Sampling knobs surfaced through the public API.  The reference bakes them into its
    graphs as constants (top_k 15, temperature 1.0, repetition_penalty 1.35, no top-p: SURVEY K7) — ``None``
    keeps exactly those; ``seed=None`` draws fresh noise per sentence like the unseeded graphs."""
    from .engine import SamplingParams
    return SamplingParams(top_k=int(top_k or 0), top_p=float(top_p) if top_p else 1.0,
                          temperature=float(temperature or 0.0), repetition_penalty=float(repetition_penalty or 0.0),
                          seed=seed)


async def tts_async(character_name: str, text: str, play: bool = False, split_sentence: bool = False,
                    save_path: Union[str, PathLike, None] = None, *, top_k: Optional[int] = None,
                    top_p: Optional[float] = None, temperature: Optional[float] = None,
                    repetition_penalty: Optional[float] = None, seed: Optional[int] = None) -> AsyncIterator[bytes]:
    if character_name not in _reference_audios:
        raise ValueError("Please call 'set_reference_audio' first to set the reference audio.")
    save_path = _prep_save_path(save_path)
    q: "asyncio.Queue[Optional[bytes]]" = asyncio.Queue()
    loop = asyncio.get_running_loop()
    _activate(character_name)
    tts_player.start_session(play=play, split=split_sentence, save_path=save_path,
                             chunk_callback=lambda c: loop.call_soon_threadsafe(q.put_nowait, c),
                             sampling=_sampling(top_k, top_p, temperature, repetition_penalty, seed))
    tts_player.feed(text)
    tts_player.end_session()
    while True:
        chunk = await q.get()
        if chunk is None:
            break
        yield chunk


def tts(character_name: str, text: str, play: bool = False, split_sentence: bool = True,
        save_path: Union[str, PathLike, None] = None, *, top_k: Optional[int] = None, top_p: Optional[float] = None,
        temperature: Optional[float] = None, repetition_penalty: Optional[float] = None,
        seed: Optional[int] = None) -> None:
    """Reference signature (Internal.py:265-309) plus keyword-only sampling knobs, all defaulting to the
    reference's graph constants."""
    if character_name not in _reference_audios:
        logger.error("Please call 'set_reference_audio' first to set the reference audio.")
        return
    save_path = _prep_save_path(save_path)
    _activate(character_name)
    tts_player.start_session(play=play, split=split_sentence, save_path=save_path,
                             sampling=_sampling(top_k, top_p, temperature, repetition_penalty, seed))
    tts_player.feed(text)
    tts_player.end_session()
    tts_player.wait_for_tts_completion()
    if play:
        tts_player.wait_for_playback_done()


def wait_for_playback_done() -> None:
    tts_player.wait_for_playback_done()


def stop() -> None:
    tts_player.stop()


def clear_reference_audio_cache() -> None:
    ReferenceAudio.clear_cache()


def convert_to_onnx(torch_ckpt_path, torch_pth_path, output_dir) -> None:
    """Offline checkpoint converter (reference Converter/**): out of scope.
    The model directory it emits is this build's weight format and is read as is."""
    raise NotImplementedError("convert_to_onnx is an offline tool of the reference and is not part of the "
                              "B200 hot-path build; convert with the reference package, load the directory here")


def load_predefined_character(character_name: str) -> None:
    raise NotImplementedError("predefined characters are downloaded assets of the reference (no network here)")


def download_genie_data() -> None:
    raise NotImplementedError("GenieData (HuBERT/SV/RoBERTa/G2P assets) is an external download of the reference")
