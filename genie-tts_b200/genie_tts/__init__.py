"""genie_tts — B200-native build of the Genie-TTS GPT-SoVITS synthesis hot path.

Public API identical to the reference package (src/genie_tts/__init__.py:16-29).
Importing this package needs neither onnxruntime nor ./GenieData nor a GPU; every
compute entry point fails loudly when libgenie_b200.so or a B200 is missing."""
from .Internal import (clear_reference_audio_cache, convert_to_onnx, download_genie_data, load_character,
                       load_predefined_character, set_reference_audio, set_reference_features, stop, tts,
                       tts_async, unload_character, wait_for_playback_done)


def start_server(host: str = "127.0.0.1", port: int = 8000, workers: int = 1):
    from .Server import start_server as _s
    return _s(host=host, port=port, workers=workers)


def start_server_per_gpu(host: str = "127.0.0.1", base_port: int = 8000, devices=None, block: bool = True):
    """Extension: one server process per GPU on consecutive ports (Server.start_server_per_gpu)."""
    from .Server import start_server_per_gpu as _s
    return _s(host=host, base_port=base_port, devices=devices, block=block)


__all__ = ["load_character", "unload_character", "set_reference_audio", "tts_async", "tts", "stop",
           "convert_to_onnx", "clear_reference_audio_cache", "start_server", "start_server_per_gpu", "wait_for_playback_done",
           "load_predefined_character", "download_genie_data", "set_reference_features"]
