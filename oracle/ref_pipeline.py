"""TEST INFRASTRUCTURE — CPU oracle, never on the product path.

Restatement of the reference's per-utterance host loop
(src/genie_tts/Core/Inference.py:27,41-61,63-109) over "session" objects with
the ``InferenceSession.run`` duck type; with ``oracle.onnx_interp.OnnxProgram``
sessions this is the reference path end to end minus G2P (host text front end,
out of scope).  Only tests/, ``__graft_entry__.smoke()`` and bench.py's
cpu_baseline / ``--impl reference`` legs may import this.

Switches the reference lacks:
  * ``greedy``      sampler noise (RandomNormalLike in both T2S decoders) := 1,
                    so token = argmax(softmax(top-k(penalised logits)))
  * ``zp_noise``    the vocoder's RandomNormalLike (vits#[6490]) returns this
                    tensor (or zeros) instead of unseeded N(0,1)
  * ``max_steps``   the decode loop bound (500 in the reference, :95)
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from .onnx_interp import OnnxProgram


@dataclass
class RefSessions:
    encoder: OnnxProgram
    first_stage: OnnxProgram
    stage: OnnxProgram
    vits: OnnxProgram
    prompt_encoder: Optional[OnnxProgram] = None


def load_sessions(model_dir: str) -> RefSessions:
    """Mirror of ModelManager.load_character (src/genie_tts/ModelManager.py:231-310):
    fp16 .bin mapping at :248-253, fp32 encoder via its own .bin at :282-286."""
    j = lambda n: os.path.join(model_dir, n)  # noqa: E731
    pe = None
    if os.path.exists(j("prompt_encoder_fp32.onnx")):
        pe = OnnxProgram(j("prompt_encoder_fp32.onnx"), j("prompt_encoder_fp16.bin"))
    return RefSessions(
        encoder=OnnxProgram(j("t2s_encoder_fp32.onnx")),
        first_stage=OnnxProgram(j("t2s_first_stage_decoder_fp32.onnx"), j("t2s_shared_fp16.bin")),
        stage=OnnxProgram(j("t2s_stage_decoder_fp32.onnx"), j("t2s_shared_fp16.bin")),
        vits=OnnxProgram(j("vits_fp32.onnx"), j("vits_fp16.bin")),
        prompt_encoder=pe,
    )


def set_sampler_mode(s: RefSessions, greedy: bool, zp_noise: Optional[np.ndarray] = None,
                     zp_zero: bool = False) -> None:
    ones = (lambda name, a: torch.ones_like(a)) if greedy else None
    s.first_stage.rng_hook = ones
    s.stage.rng_hook = ones
    if zp_noise is not None:
        zp = torch.from_numpy(np.asarray(zp_noise, dtype=np.float32))

        def hook(name, a):
            return zp[..., : a.shape[-1]].reshape(a.shape)
        s.vits.rng_hook = hook
    elif zp_zero:
        s.vits.rng_hook = lambda name, a: torch.zeros_like(a)
    else:
        s.vits.rng_hook = None


def t2s_cpu(s: RefSessions, ref_seq, ref_bert, text_seq, text_bert, ssl_content,
            max_steps: int = 500, stop_event=None, collect: Optional[Dict] = None, on_step=None):
    """src/genie_tts/Core/Inference.py:63-109, line for line.  ``on_step(idx, y)`` (tests only) is called after
    the first-stage run (idx == -1) and after every stage-decoder run, e.g. to read traced logits."""
    x, prompts = s.encoder.run(None, {                                   # :76-85
        "ref_seq": ref_seq, "text_seq": text_seq, "ref_bert": ref_bert,
        "text_bert": text_bert, "ssl_content": ssl_content})
    y, y_emb, *present_key_values = s.first_stage.run(None, {"x": x, "prompts": prompts})   # :88-90
    if collect is not None:
        collect.update(x=x, prompts=prompts, y0=y.copy(), kv0=[k.copy() for k in present_key_values[:2]])
    if on_step is not None:
        on_step(-1, y)
    input_names: List[str] = [inp.name for inp in s.stage.get_inputs()]  # :93
    idx = 0
    for idx in range(0, max_steps):                                      # :95
        if stop_event is not None and stop_event.is_set():               # :96-97
            return None
        feed = {name: data for name, data in zip(input_names, [y, y_emb, *present_key_values])}
        outputs = s.stage.run(None, feed)                                # :102
        y, y_emb, stop_condition_tensor, *present_key_values = outputs
        if on_step is not None:
            on_step(idx, y)
        if stop_condition_tensor:                                        # :105-106
            break
    if collect is not None:
        collect.update(y_full=y.copy(), idx=idx)
    y[0, -1] = 0                                                          # :108
    return np.expand_dims(y[:, -idx:], axis=0)                            # :109 (idx==0 -> whole y)


def strip_eos(semantic_tokens: np.ndarray) -> np.ndarray:
    """src/genie_tts/Core/Inference.py:41-44."""
    eos_indices = np.where(semantic_tokens >= 1024)
    if len(eos_indices[0]) > 0:
        first_eos_index = eos_indices[-1][0]
        semantic_tokens = semantic_tokens[..., :first_eos_index]
    return semantic_tokens


def vocode(s: RefSessions, text_seq, semantic_tokens, ref_audio_32k=None, ge=None, ge_advanced=None):
    """src/genie_tts/Core/Inference.py:46-61."""
    if s.prompt_encoder is None:
        return s.vits.run(None, {"text_seq": text_seq, "pred_semantic": semantic_tokens,
                                 "ref_audio": ref_audio_32k})[0]
    return s.vits.run(None, {"text_seq": text_seq, "pred_semantic": semantic_tokens,
                             "ge": ge, "ge_advanced": ge_advanced})[0]


def prompt_global_emb(s: RefSessions, ref_audio_32k, sv_emb):
    """src/genie_tts/Audio/ReferenceAudio.py:68-76 (V2ProPlus only)."""
    return s.prompt_encoder.run(None, {"ref_audio": ref_audio_32k, "sv_emb": sv_emb})


def tts(s: RefSessions, prompt: Dict[str, np.ndarray], text_seq, text_bert, max_steps: int = 500):
    """src/genie_tts/Core/Inference.py:16-61 after g2p (``text_seq``/``text_bert``
    are get_phones_and_bert's outputs for '。'+text)."""
    toks = t2s_cpu(s, prompt["ref_seq"], prompt["ref_bert"], text_seq, text_bert,
                   prompt["ssl_content"], max_steps=max_steps)
    toks = strip_eos(toks)
    if s.prompt_encoder is None:
        return toks, vocode(s, text_seq, toks, ref_audio_32k=prompt["ref_audio"])
    if "ge" not in prompt:
        prompt["ge"], prompt["ge_advanced"] = prompt_global_emb(s, prompt["ref_audio"], prompt["sv_emb"])
    return toks, vocode(s, text_seq, toks, ge=prompt["ge"], ge_advanced=prompt["ge_advanced"])
