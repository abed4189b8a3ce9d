"""Sentence splitter with the reference's semantics (src/genie_tts/Utils/TextSplitter.py:5-123):
punctuation runs are separators; a run containing a sentence-final mark closes the
sentence once its content width >= min_len; any other run closes it once the width
>= max_len.  Width counts ASCII as 1 and everything else as 2, punctuation as 0."""
import re
from typing import List

_FINAL = "。！？…!?."
_OTHER = ["，", "、", "；", "：", "——", ",", ";", ":", "“", "”", "‘", "’", '"', "'"]


class TextSplitter:
    def __init__(self, max_len: int = 40, min_len: int = 5):
        self.max_len, self.min_len = max_len, min_len
        self.end_chars = set(_FINAL)
        marks = sorted(list(_FINAL) + _OTHER, key=len, reverse=True)
        self.all_puncts_chars = set(_FINAL) | set(_OTHER)
        self.pattern = re.compile("((?:" + "|".join(re.escape(m) for m in marks) + ")+)")

    def get_effective_len(self, text: str) -> int:
        return sum((1 if ord(ch) < 128 else 2) for ch in text if ch not in self.all_puncts_chars)

    def split(self, text: str) -> List[str]:
        if not text:
            return []
        out: List[str] = []
        buf = ""
        for piece in self.pattern.split(text.replace("\n", "")):
            if not piece:
                continue
            buf += piece
            if piece[0] not in self.all_puncts_chars:
                continue                                    # plain text: keep accumulating
            width = self.get_effective_len(buf)
            final = any(ch in self.end_chars for ch in piece)
            if (final and width >= self.min_len) or (not final and width >= self.max_len):
                out.append(buf.strip())
                buf = ""
        tail = buf.strip()
        if tail:
            if self.get_effective_len(tail) > 0:
                out.append(tail)
            elif out:
                out[-1] += tail                             # trailing punctuation joins the last sentence
        return out
