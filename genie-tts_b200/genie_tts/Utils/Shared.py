"""Process-wide "current speaker / current prompt" (reference: src/genie_tts/Utils/Shared.py:7-13).
Kept for API compatibility; the batched server path carries these per request instead."""
from typing import Any, Optional


class Context:
    def __init__(self) -> None:
        self.current_speaker: str = ""
        self.current_prompt_audio: Optional[Any] = None


context = Context()
