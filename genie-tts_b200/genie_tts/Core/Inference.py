"""Per-utterance pipeline on the GPU (reference: src/genie_tts/Core/Inference.py:12-112).

``GENIE.tts`` / ``GENIE.t2s_cpu`` keep the reference's argument lists — the five
"session" arguments are ``StageSession`` handles onto one ``B200Model`` — so
``TTSPlayer._tts_worker_loop`` calls them unchanged.  The encoder + first-stage +
<=500 stage-decoder ``run`` calls collapse into ONE ``genie_t2s_generate`` C-ABI
call; the vocoder ``run`` into ONE ``genie_vits_decode``.  ``tts_batch`` is the
batched form of the same path (configs 2-4)."""
from __future__ import annotations

import ctypes
import threading
from typing import List, Optional, Sequence

import numpy as np

from ..GetPhonesAndBert import get_phones_and_bert
from ..engine import SamplingParams

import os as _os
# one vocoder call per wave instead of one per batch: +1.5 % device-resident, but -4 % through host buffers (measured,
# DESIGN.md section 5), so the default keeps one call per batch
MERGE_WAVE_VOCODER = _os.environ.get("GENIE_WAVE_VOCODER", "split") == "merged"
MAX_T2S_LEN = 1000
MAX_DECODE_STEPS = 500      # reference Inference.py:95


def finish_t2s(y_full: np.ndarray, idx: int) -> np.ndarray:
    """The reference's post-loop slicing, quirks included (Inference.py:108-109):
    ``y[0,-1] = 0`` then ``y[:, -idx:]`` — idx == 0 returns the whole sequence
    (prompt tokens included); the first two generated tokens are otherwise dropped."""
    y = np.array(y_full, dtype=np.int64).reshape(1, -1)
    y[0, -1] = 0
    return np.expand_dims(y[:, -idx:], axis=0)


def strip_eos(semantic_tokens: np.ndarray) -> np.ndarray:
    """Inference.py:41-44: cut at the first id >= 1024 along the last axis."""
    hits = np.where(semantic_tokens >= 1024)
    if len(hits[0]) > 0:
        semantic_tokens = semantic_tokens[..., :hits[-1][0]]
    return semantic_tokens


class _CancelFlag:
    """threading.Event whose state is also visible to the C-ABI decode loop
    (reference: stop_event polled once per decode step, Inference.py:96-97)."""

    def __init__(self):
        self._ev = threading.Event()
        self.c_flag = ctypes.c_int(0)

    def set(self):
        self.c_flag.value = 1
        self._ev.set()

    def clear(self):
        self.c_flag.value = 0
        self._ev.clear()

    def is_set(self) -> bool:
        return self._ev.is_set()

    def wait(self, timeout=None):
        return self._ev.wait(timeout)


class GENIE:
    def __init__(self):
        self.stop_event = _CancelFlag()
        self.sampling = SamplingParams()          # graph-constant defaults; tests flip .greedy
        self.zp_noise_hook = None                 # tests: callable(n_tokens) -> f32[1,192,2n]

    # -- reference signature ---------------------------------------------------
    def tts(self, text: str, prompt_audio, encoder, first_stage_decoder, stage_decoder, vocoder,
            prompt_encoder=None, language: str = "japanese") -> Optional[np.ndarray]:
        text = "。" + text                                                   # Inference.py:27
        text_seq, text_bert = get_phones_and_bert(text, language=language)
        model = vocoder.model
        prompt = prompt_audio.device_prompt(model)
        semantic_tokens = self._t2s(model, prompt, text_seq, text_bert)
        if semantic_tokens is None:
            return None
        semantic_tokens = strip_eos(semantic_tokens)
        if semantic_tokens.shape[-1] == 0:
            return np.zeros(0, dtype=np.float32)
        if prompt_encoder is not None:
            prompt_audio.update_global_emb(prompt_encoder=prompt_encoder)    # Inference.py:54
        sem = semantic_tokens.reshape(-1)
        noise = [self.zp_noise_hook(len(sem))] if self.zp_noise_hook else None
        return model.vits_decode([prompt], [text_seq], [sem], noise, seed=self.sampling.seed)[0]

    def t2s_cpu(self, ref_seq, ref_bert, text_seq, text_bert, ssl_content, encoder, first_stage_decoder,
                stage_decoder) -> Optional[np.ndarray]:
        """Name kept from the reference; runs on the GPU.  Returns int64 [1,1,N] or None when stopped."""
        model = stage_decoder.model
        ge = np.zeros(1024 if model.is_v2pp else 512, dtype=np.float32)     # T2S does not read ge
        prompt = model.make_prompt(ref_seq, ref_bert, ssl_content, ge=ge,
                                   ge_advanced=np.zeros(512, np.float32) if model.is_v2pp else None)
        try:
            return self._t2s(model, prompt, text_seq, text_bert)
        finally:
            prompt.close()

    def _t2s(self, model, prompt, text_seq, text_bert) -> Optional[np.ndarray]:
        if self.stop_event.is_set():
            return None
        sp = self.sampling
        if sp.max_steps <= 0:
            sp.max_steps = MAX_DECODE_STEPS
        ys, idx = model.t2s_generate([prompt], [text_seq], [text_bert], sp, cancel_flag=self.stop_event.c_flag)
        if not ys:
            return None
        return finish_t2s(ys[0], idx[0])

    # -- batched form of the same path -------------------------------------------
    def tts_batch(self, model, prompts: Sequence, text_seqs: Sequence[np.ndarray],
                  text_berts: Optional[Sequence[np.ndarray]] = None,
                  sampling: Optional[SamplingParams] = None, zp_noise: Optional[Sequence[np.ndarray]] = None
                  ) -> List[np.ndarray]:
        """B independent utterances: one T2S call + one vocoder call.  Per-utterance semantics are
        exactly those of ``tts`` (slicing quirks and EOS strip included)."""
        sp = sampling or self.sampling
        ys, idx = model.t2s_generate(prompts, text_seqs, text_berts, sp, cancel_flag=self.stop_event.c_flag)
        if not ys:
            return []
        return self._vocode_batch(model, prompts, text_seqs, ys, idx, sp, zp_noise)

    def _vocode_batch(self, model, prompts, text_seqs, ys, idx, sp, zp_noise, noise_ids=None) -> List[np.ndarray]:
        """Host glue of Inference.py:41-44,108-109 per utterance, then ONE vocoder call for the batch.
        ``noise_ids[b]``: the utterance's index within its own batch when several batches share the call."""
        sems = [strip_eos(finish_t2s(y, i)).reshape(-1) for y, i in zip(ys, idx)]
        keep = [b for b, s in enumerate(sems) if len(s) > 0]
        out: List[np.ndarray] = [np.zeros(0, np.float32) for _ in sems]
        if keep:
            # the vocoder numbers the utterances it receives 0..n-1: keep every utterance on the noise stream
            # of its position among its own batch's non-empty utterances
            if noise_ids is not None:
                ids = [noise_ids[b] for b in keep]
            else:
                ids = None
            auds = model.vits_decode([prompts[b] for b in keep], [text_seqs[b] for b in keep],
                                     [sems[b] for b in keep],
                                     [zp_noise[b] for b in keep] if zp_noise is not None else None, seed=sp.seed,
                                     noise_ids=ids)
            for b, a in zip(keep, auds):
                out[b] = a
        return out


    def tts_batch_stream(self, model, batches, sampling: Optional[SamplingParams] = None, depth: int = 3):
        """Throughput form of ``tts_batch`` for a stream of batches: ``depth`` batches are in flight together on
        ``depth`` execution contexts of ``model`` (same weights, own streams / workspaces), and each batch's
        waveforms are yielded in order.  The schedule is stage-aligned, per wave of ``depth`` batches:

            prefill of every batch (one after the other: throughput-bound, nothing to gain from overlap)
            decode of ALL batches at the same time (a decode step is a chain of latency-bound kernels; measured:
                two 100-sentence decodes side by side take 1.4x the time of one — DESIGN.md section 5)
            SoVITS of every batch (one after the other; GENIE_WAVE_VOCODER=merged: one call for the wave)

        Per-batch results are exactly those of ``tts_batch``.  ``batches``: iterable of
        (prompts, text_seqs, text_berts or None)."""
        from concurrent.futures import ThreadPoolExecutor
        sp = sampling or self.sampling
        ctxs = model.pipeline_contexts(max(1, depth))
        it = iter(batches)
        with ThreadPoolExecutor(max_workers=len(ctxs)) as ex:
            while True:
                wave = []
                for _ in ctxs:
                    try:
                        wave.append(next(it))
                    except StopIteration:
                        break
                if not wave:
                    return
                for ctx, b in zip(ctxs, wave):
                    ctx.t2s_prefill(b[0], b[1], b[2] if len(b) > 2 else None, sp)

                def decode(ctx):
                    steps = sp.fixed_steps if sp.fixed_steps > 0 else (sp.max_steps if sp.max_steps > 0 else MAX_DECODE_STEPS)
                    _, _, cancelled = ctx.t2s_decode_steps(steps, cancel_flag=self.stop_event.c_flag)
                    return None if cancelled else ctx.t2s_read()
                toks = list(ex.map(decode, ctxs[:len(wave)]))
                # ONE vocoder call for the whole wave (throughput-bound: a 200-utterance pass costs ~10 % less than
                # two 100-utterance passes); every utterance keeps the Philox stream it has in a call of its own
                # batch: its position among that batch's non-empty utterances
                prm, seqs, ys, idx, ids, spans = [], [], [], [], [], []
                for b, tk in zip(wave, toks):
                    if tk is None:
                        spans.append(None)
                        continue
                    spans.append((len(ys), len(tk[0])))
                    prm += list(b[0]); seqs += list(b[1]); ys += tk[0]; idx += tk[1]
                    nonempty = 0
                    for y, i in zip(tk[0], tk[1]):
                        ids.append(nonempty)
                        if strip_eos(finish_t2s(y, i)).shape[-1] > 0:
                            nonempty += 1
                if MERGE_WAVE_VOCODER:
                    auds = self._vocode_batch(ctxs[0], prm, seqs, ys, idx, sp, None, noise_ids=ids) if ys else []
                    for sp_ in spans:
                        yield [] if sp_ is None else auds[sp_[0]:sp_[0] + sp_[1]]
                else:
                    for ctx, b, tk in zip(ctxs, wave, toks):
                        yield [] if tk is None else self._vocode_batch(ctx, b[0], b[1], tk[0], tk[1], sp, None)


tts_client: GENIE = GENIE()
