"""Text front end boundary: text -> (phones_seq int64[1,L], text_bert f32[L,1024]).

The reference's G2P (src/genie_tts/G2P/**, 2.5 k lines over pyopenjtalk / g2pM /
nltk) and RoBERTa features (src/genie_tts/GetPhonesAndBert.py:33-83) stay on the
host, untimed, and are OUT OF SCOPE of this build: this
module keeps the call signature and lets the deployment plug the reference's own
implementation in with ``set_text_frontend``."""
from typing import Callable, Optional, Tuple

import numpy as np

from .Utils.Constants import BERT_FEATURE_DIM

Frontend = Callable[[str, str], Tuple[np.ndarray, np.ndarray]]
_frontend: Optional[Frontend] = None


def set_text_frontend(fn: Optional[Frontend]) -> None:
    """fn(text, language) -> (int64[1,L] phoneme ids in SymbolsV2 order, f32[L,1024])."""
    global _frontend
    _frontend = fn


def get_phones_and_bert(prompt_text: str, language: str = "japanese") -> Tuple[np.ndarray, np.ndarray]:
    if _frontend is None:
        try:   # the reference's own front end, when its G2P package and dictionaries are installed
            from genie_tts_g2p import get_phones_and_bert as ref_fn   # type: ignore
        except Exception as e:
            raise RuntimeError(
                "no text front end registered: G2P/BERT are host-side components of the reference "
                "(src/genie_tts/GetPhonesAndBert.py) outside this build's scope; call "
                "genie_tts.GetPhonesAndBert.set_text_frontend(fn)") from e
        return ref_fn(prompt_text, language)
    seq, bert = _frontend(prompt_text, language)
    seq = np.asarray(seq, dtype=np.int64).reshape(1, -1)
    if bert is None:
        bert = np.zeros((seq.shape[1], BERT_FEATURE_DIM), dtype=np.float32)
    return seq, np.asarray(bert, dtype=np.float32)
