"""FastAPI front end (reference: src/genie_tts/Server.py:25-165): same routes and payloads;
/tts streams raw s16 PCM @32 kHz, one chunk per sentence."""
from __future__ import annotations

import asyncio
from typing import AsyncIterator, Optional

import uvicorn
from fastapi import FastAPI, HTTPException
from fastapi.responses import StreamingResponse
from pydantic import BaseModel

from .Audio.ReferenceAudio import ReferenceAudio
from .Core.TTSPlayer import tts_player
from .Internal import _activate, _reference_audios, load_character, set_reference_audio, unload_character
from .ModelManager import model_manager

app = FastAPI()


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


@app.post("/load_character")
def load_character_endpoint(payload: CharacterPayload):
    try:
        load_character(payload.character_name, payload.onnx_model_dir, payload.language)
    except Exception as e:
        raise HTTPException(status_code=500, detail=str(e))
    return {"status": "success", "message": f"Character '{payload.character_name}' loaded."}


@app.post("/unload_character")
def unload_character_endpoint(payload: UnloadCharacterPayload):
    try:
        unload_character(payload.character_name)
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


def run_tts_in_background(character_name: str, text: str, split_sentence: bool, save_path: Optional[str],
                          chunk_callback) -> None:
    try:
        _activate(character_name)
        tts_player.start_session(play=False, split=split_sentence, save_path=save_path, chunk_callback=chunk_callback)
        tts_player.feed(text)
        tts_player.end_session()
        tts_player.wait_for_tts_completion()
    except Exception:
        chunk_callback(None)
        raise


async def audio_stream_generator(queue: "asyncio.Queue") -> AsyncIterator[bytes]:
    while True:
        chunk = await queue.get()
        if chunk is None:
            return
        yield chunk


@app.post("/tts")
async def tts_endpoint(payload: TTSPayload):
    if payload.character_name not in _reference_audios:
        raise HTTPException(status_code=404, detail="Character not found or reference audio not set.")
    if model_manager.get(payload.character_name) is None:
        raise HTTPException(status_code=404, detail="Character not loaded.")
    loop = asyncio.get_running_loop()
    q: "asyncio.Queue" = asyncio.Queue()
    loop.run_in_executor(None, run_tts_in_background, payload.character_name, payload.text, payload.split_sentence,
                         payload.save_path, lambda c: loop.call_soon_threadsafe(q.put_nowait, c))
    return StreamingResponse(audio_stream_generator(q), media_type="audio/wav")


@app.post("/stop")
def stop_endpoint():
    tts_player.stop()
    return {"status": "success", "message": "TTS stopped."}


@app.post("/clear_reference_audio_cache")
def clear_reference_audio_cache_endpoint():
    ReferenceAudio.clear_cache()
    return {"status": "success", "message": "Reference audio cache cleared."}


def start_server(host: str = "127.0.0.1", port: int = 8000, workers: int = 1):
    uvicorn.run(app, host=host, port=port, workers=workers)
