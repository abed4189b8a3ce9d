"""FastAPI front end: the reference's REST surface (src/genie_tts/Server.py:25-165 — same six routes, same JSON
payloads, /tts streams raw s16 PCM @32 kHz, one chunk per sentence) on top of ``SynthesisService``.

What is different from the reference is behind the routes: there is no global ``tts_player`` / ``context`` on the
request path.  Every /tts is a ``RequestStream`` of per-sentence futures served by continuous batching on every GPU
of the box, so concurrent requests neither wait for each other's sentences nor clear each other's queues (the
reference's ``start_session`` does, Core/TTSPlayer.py:185-186)."""
from __future__ import annotations

import asyncio
import logging
import os
import threading
from typing import AsyncIterator, Optional

import uvicorn
from fastapi import FastAPI, HTTPException
from fastapi.responses import StreamingResponse
from pydantic import BaseModel

from .Audio.ReferenceAudio import ReferenceAudio
from .Internal import (_prep_save_path, _reference_audios, _sampling, check_onnx_model_dir, set_reference_audio)
from .Service import RequestStream, SynthesisService
from .Utils.Language import normalize_language

logger = logging.getLogger(__name__)
app = FastAPI()
_service: Optional[SynthesisService] = None
_service_lock = threading.Lock()


def get_service() -> SynthesisService:
    """One service per server process, created on first use (needs a GPU)."""
    global _service
    with _service_lock:
        if _service is None:
            _service = SynthesisService(n_slots=int(os.getenv("GENIE_SLOTS", "256")),
                                        kv_capacity=int(os.getenv("GENIE_KV_CAPACITY", "1024")))
        return _service


def set_service(service: Optional[SynthesisService]) -> None:
    global _service
    with _service_lock:
        _service = service


class CharacterPayload(BaseModel):
    character_name: str
    onnx_model_dir: str
    language: str


class UnloadCharacterPayload(BaseModel):
    character_name: str


class ReferenceAudioPayload(BaseModel):
    character_name: str
    audio_path: str
    audio_text: str
    language: str


class TTSPayload(BaseModel):
    character_name: str
    text: str
    split_sentence: bool = False
    save_path: Optional[str] = None
    # extension: sampling knobs, None = the reference's graph constants
    top_k: Optional[int] = None
    top_p: Optional[float] = None
    temperature: Optional[float] = None
    repetition_penalty: Optional[float] = None
    seed: Optional[int] = None


def _reference_of(name: str) -> Optional[ReferenceAudio]:
    entry = _reference_audios.get(name)
    if entry is None:
        return None
    if "reference" in entry:
        return entry["reference"]           # type: ignore[return-value]
    return ReferenceAudio(prompt_wav=entry["audio_path"], prompt_text=entry["audio_text"], language=entry["language"])


@app.post("/load_character")
def load_character_endpoint(payload: CharacterPayload):
    try:
        check_onnx_model_dir(payload.onnx_model_dir)
        language = normalize_language(payload.language)
        if language not in ("Japanese", "English", "Chinese", "Hybrid-Chinese-English"):
            raise ValueError("Unknown language")
        get_service().load_character(payload.character_name, payload.onnx_model_dir, language)
    except Exception as e:
        raise HTTPException(status_code=500, detail=str(e))
    return {"status": "success", "message": f"Character '{payload.character_name}' loaded."}


@app.post("/unload_character")
def unload_character_endpoint(payload: UnloadCharacterPayload):
    try:
        get_service().unload_character(payload.character_name)
    except Exception as e:
        raise HTTPException(status_code=500, detail=str(e))
    return {"status": "success", "message": f"Character '{payload.character_name}' unloaded."}


@app.post("/set_reference_audio")
def set_reference_audio_endpoint(payload: ReferenceAudioPayload):
    try:
        set_reference_audio(payload.character_name, payload.audio_path, payload.audio_text, payload.language)
    except Exception as e:
        raise HTTPException(status_code=500, detail=str(e))
    return {"status": "success", "message": "Reference audio set."}


async def audio_stream_generator(stream: RequestStream) -> AsyncIterator[bytes]:
    """Sentence futures -> PCM chunks, in text order, without blocking the event loop."""
    from .Service import pcm16
    saved = []
    for fut in stream.futures:
        try:
            audio = await asyncio.wrap_future(fut)
        except Exception as e:                      # reference: log, keep the stream alive (TTSPlayer.py:109-114)
            logger.error(f"/tts: sentence failed: {e}")
            continue
        if audio is None or len(audio) == 0:
            continue
        chunk = pcm16(audio)
        if stream.save_path:
            saved.append(chunk)
        yield chunk
    if stream.save_path and saved:
        stream._saved = saved
        stream.finish()


@app.post("/tts")
async def tts_endpoint(payload: TTSPayload):
    svc = get_service()
    ref = _reference_of(payload.character_name)
    if ref is None:
        raise HTTPException(status_code=404, detail="Character not found or reference audio not set.")
    svc.set_reference(payload.character_name, ref)
    try:
        sp = None
        if any(v is not None for v in (payload.top_k, payload.top_p, payload.temperature,
                                       payload.repetition_penalty, payload.seed)):
            sp = _sampling(payload.top_k, payload.top_p, payload.temperature, payload.repetition_penalty, payload.seed)
        loop = asyncio.get_running_loop()
        # phonemisation + dispatch happen off the event loop (G2P is host work of unbounded cost)
        stream = await loop.run_in_executor(None, lambda: svc.submit(
            payload.character_name, payload.text, payload.split_sentence, _prep_save_path(payload.save_path), sp))
    except KeyError:
        raise HTTPException(status_code=404, detail="Character not loaded.")
    return StreamingResponse(audio_stream_generator(stream), media_type="audio/wav")


@app.post("/stop")
def stop_endpoint():
    get_service().stop_all()
    return {"status": "success", "message": "TTS stopped."}


@app.post("/clear_reference_audio_cache")
def clear_reference_audio_cache_endpoint():
    ReferenceAudio.clear_cache()
    return {"status": "success", "message": "Reference audio cache cleared."}


@app.get("/stats")
def stats_endpoint():
    """Extension: per-replica scheduler counters (batches, requests, latency percentiles)."""
    return get_service().stats()


def start_server(host: str = "127.0.0.1", port: int = 8000, workers: int = 1):
    # one process drives every GPU of the box (one scheduler thread per GPU and character); `workers` > 1 would
    # duplicate the weights per worker process, exactly as the reference's uvicorn workers do (Server.py:165)
    uvicorn.run(app, host=host, port=port, workers=workers)


def _serve_devices(host: str, port: int, devices) -> None:
    os.environ["GENIE_DEVICES"] = ",".join(str(d) for d in devices)
    uvicorn.run(app, host=host, port=port)


def start_server_per_gpu(host: str = "127.0.0.1", base_port: int = 8000, devices=None, block: bool = True):
    """One server PROCESS per GPU on ports base_port, base_port + 1, ... (put any HTTP balancer in front).  One Python
    process tops out near 1500 requests/s on its interpreter lock - 8 GPUs behind one process reach 4.9 k audio-s/s,
    behind 8 processes 15.2 k (profiles/r02/config5_load_8gpu*.json).  Characters and reference audio are
    per process: send /load_character and /set_reference_audio to every port."""
    import multiprocessing as mp
    from . import _native as N
    N.require_gpu()
    if devices is None:
        devices = list(range(N.lib().genie_device_count()))
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_serve_devices, args=(host, base_port + i, [d]), daemon=False)
             for i, d in enumerate(devices)]
    for p in procs:
        p.start()
    if block:
        for p in procs:
            p.join()
    return procs

