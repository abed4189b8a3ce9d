"""Language aliases (reference: src/genie_tts/Utils/Language.py:1-31)."""
_ALIASES = {
    "Chinese": ("chinese", "zh", "zh-cn", "zh-tw", "zh-hans", "zh-hant"),
    "English": ("english", "en", "en-us", "en-gb", "eng"),
    "Japanese": ("japanese", "jp", "ja", "nihongo"),
    "Hybrid-Chinese-English": ("hybrid", "hybrid-zh-en", "hybrid-en-zh"),
}
language_map = {alias: canon for canon, aliases in _ALIASES.items() for alias in aliases}


def normalize_language(lang: str) -> str:
    """Canonical language name; unknown strings pass through unchanged."""
    return language_map.get(lang.lower(), lang)
