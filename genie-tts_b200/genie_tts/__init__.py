"""genie_tts on B200 — placeholder package init (public API wired in Internal.py)."""
