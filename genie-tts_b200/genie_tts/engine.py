"""Device-resident model / prompt objects over the C-ABI.

``B200Model`` replaces the five onnxruntime sessions of one character
(reference: GSVModel, src/genie_tts/ModelManager.py:48-56); ``B200Prompt``
replaces the per-reference-audio work (src/genie_tts/Audio/ReferenceAudio.py:
28-76 features + the VQ / ref_enc parts of the graphs).  Batched entry points
take lists of utterances; the batch-1 ``GENIE.tts`` path is a batch of one.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .weights import read_model_dir

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "Data")


def load_constants() -> dict:
    with open(os.path.join(_DATA, "t2s_constants.json")) as f:
        c = json.load(f)
    c["pe_div_term"] = np.frombuffer(bytes.fromhex(c["pe_div_term_f32_hex"]), dtype="<f4").copy()
    return c


def _ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a.data_ptr()))   # torch tensor (device-resident leg)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).reshape(-1), dtype=np.int64)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


@dataclass
class SamplingParams:
    """Defaults (<=0) fall back to the graph constants: top_k 15, temperature 1.0,
    repetition_penalty 1.35 (t2s_stage_decoder#[1780-1790]).  ``top_p`` is an extension (the reference graphs have
    no top-p node); 1.0 = off.  ``greedy`` is the test-only switch that makes token sequences comparable with the
    oracle.  ``seed=None`` draws a fresh 64-bit seed per call — the reference's RandomNormalLike nodes are unseeded,
    so successive sentences must not share a noise stream; pass an int for reproducible output."""
    top_k: int = 0
    temperature: float = 0.0
    repetition_penalty: float = 0.0
    greedy: bool = False
    seed: Optional[int] = None
    max_steps: int = 500
    fixed_steps: int = 0
    top_p: float = 1.0

    def resolved_seed(self) -> int:
        if self.seed is None:
            return int.from_bytes(os.urandom(8), "little")
        return self.seed & 0xFFFFFFFFFFFFFFFF

    def to_c(self, seed: Optional[int] = None) -> N.Sampling:
        return N.Sampling(self.top_k, self.temperature, self.repetition_penalty, int(self.greedy),
                          self.resolved_seed() if seed is None else seed, self.max_steps, self.fixed_steps,
                          float(self.top_p))


class _PinnedPool:
    """Page-locked result buffers, recycled when every numpy view of a buffer has been dropped.

    ``cudaHostAlloc`` costs milliseconds, so buffers are kept: ``take(n)`` returns a float32 array of n elements
    backed by pinned memory; a ``weakref.finalize`` on that array returns the block to the pool (slices handed to
    callers keep the array alive through ``.base``)."""

    MAX_KEEP = 8

    def __init__(self):
        import threading
        self._free: List[Tuple[int, int]] = []          # (bytes, address)
        self._lock = threading.Lock()

    def take(self, n_floats: int) -> np.ndarray:
        import weakref
        nbytes = max(4, int(n_floats) * 4)
        with self._lock:
            fit = [i for i, (b, _) in enumerate(self._free) if b >= nbytes]
            blk = self._free.pop(min(fit, key=lambda i: self._free[i][0])) if fit else None
        if blk is None:
            cap = max(1 << 20, 1 << (nbytes - 1).bit_length())        # power-of-two blocks: steady sizes get reused
            ptr = C.c_void_p()
            N.check(N.lib().genie_host_alloc(C.c_size_t(cap), C.byref(ptr)))
            blk = (cap, ptr.value)
        cap, addr = blk
        arr = np.frombuffer((C.c_char * cap).from_address(addr), dtype=np.float32, count=int(n_floats))
        weakref.finalize(arr, self._release, cap, addr)
        return arr

    def _release(self, cap: int, addr: int) -> None:
        with self._lock:
            if len(self._free) < self.MAX_KEEP:
                self._free.append((cap, addr))
                return
        try:
            N.lib().genie_host_free(C.c_void_p(addr))
        except Exception:
            pass


_pinned = _PinnedPool()


class B200Model:
    def __init__(self, model_dir: str, device: int = 0):
        N.require_gpu()
        tabs = read_model_dir(model_dir)        # raises FileNotFoundError like the reference loader
        self.model_dir = model_dir
        self.device = device
        self.is_v2pp = tabs.is_v2pp
        self._h = C.c_void_p(0)
        self._prompts: "weakref.WeakSet[B200Prompt]" = weakref.WeakSet()   # closed with the model
        self._contexts: "weakref.WeakSet[B200Model]" = weakref.WeakSet()
        self._pool = None
        self.lock = threading.RLock()           # callers that share a handle between threads (server workers)
        L = N.lib()
        N.check(L.genie_model_create(device, C.byref(self._h)))
        try:
            for graph, tab in ((N.GRAPH_T2S_ENCODER, tabs.encoder), (N.GRAPH_T2S, tabs.t2s),
                               (N.GRAPH_VITS, tabs.vits), (N.GRAPH_PROMPT_ENCODER, tabs.prompt_encoder)):
                if tab is None:
                    continue
                for name, arr in tab.tensors.items():
                    a = np.ascontiguousarray(arr)     # memmap view -> bytes as stored (fp16 stays fp16)
                    dt = N.F16 if a.dtype == np.float16 else N.F32
                    if dt == N.F32 and a.dtype != np.float32:
                        a = a.astype(np.float32)
                    dims = (C.c_int64 * max(a.ndim, 1))(*a.shape)
                    N.check(L.genie_model_add_tensor(self._h, graph, name.encode(), _ptr(a), dt, dims, a.ndim))
            c = load_constants()
            self.constants = c
            div = _f32(c["pe_div_term"])
            N.check(L.genie_model_set_constants(self._h, _ptr(div), c["top_k"], c["repetition_penalty"],
                                                c["temperature"], c["vits_noise_scale"]))
            N.check(L.genie_model_finalize(self._h))
        except Exception:
            L.genie_model_destroy(self._h)
            self._h = C.c_void_p(0)
            raise

    # -- lifecycle -----------------------------------------------------------
    @property
    def closed(self) -> bool:
        return not self._h

    def close(self) -> None:
        """Releases the handle and everything built on it: contexts and device prompts are closed FIRST, so that no
        live ``B200Prompt`` keeps HBM of an evicted character and none outlives its model by accident (the C side
        tolerates either order; this is about releasing memory when the LRU evicts a character)."""
        with self.lock:
            for c in list(self._contexts):
                c.close()
            for p in list(self._prompts):
                p.close()
            if self._h:
                N.lib().genie_model_destroy(self._h)
                self._h = C.c_void_p(0)

    def create_context(self, cuda_stream: Optional[int] = None) -> "B200Model":
        """A second independent execution context on the same weights (own stream / workspace / slot pool /
        graphs): e.g. a continuous-batching scheduler plus a latency-critical batch-1 caller on one GPU.
        ``cuda_stream``: a caller-owned ``cudaStream_t`` (as int, e.g. ``torch.cuda.Stream().cuda_stream``)."""
        ctx = object.__new__(B200Model)
        ctx.model_dir, ctx.device, ctx.is_v2pp, ctx.constants = self.model_dir, self.device, self.is_v2pp, self.constants
        ctx._h = C.c_void_p(0)
        ctx._prompts = self._prompts              # prompts are valid on every handle of the model
        ctx._contexts = weakref.WeakSet()
        ctx._pool = None
        ctx.lock = threading.RLock()
        N.check(N.lib().genie_context_create(self._h, C.c_void_p(cuda_stream or 0), C.byref(ctx._h)))
        self._contexts.add(ctx)
        return ctx

    def pipeline_contexts(self, depth: int) -> List["B200Model"]:
        """``depth`` handles onto these weights (this one first): batches synthesised concurrently on them overlap
        the latency-bound T2S decode of one batch with the throughput-bound SoVITS decode of another."""
        ctxs = self.__dict__.setdefault("_pipe_ctxs", [])
        ctxs[:] = [c for c in ctxs if not c.closed]
        while len(ctxs) < depth - 1:
            ctxs.append(self.create_context())
        return [self] + ctxs[:depth - 1]

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        N.check(N.lib().genie_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        v, wb, sb = C.c_int(0), C.c_longlong(0), C.c_longlong(0)
        N.check(N.lib().genie_model_info(self._h, C.byref(v), C.byref(wb), C.byref(sb)))
        return {"is_v2pp": bool(v.value), "weight_bytes": wb.value, "workspace_bytes": sb.value}

    def set_option(self, key: str, value: int) -> None:
        N.check(N.lib().genie_set_option(self._h, key.encode(), int(value)))
        self.__dict__.setdefault("options", {})[key] = int(value)

    @property
    def kv_bytes_per_element(self) -> int:
        """2: KV cache rows stored as fp16 (default), 4: fp32 rows (option kv_fp16 = 0)."""
        return 2 if self.__dict__.get("options", {}).get("kv_fp16", 1) else 4

    # -- prompt ----------------------------------------------------------------
    def make_prompt(self, ref_seq, ref_bert, ssl_content, ref_audio_32k=None, sv_emb=None,
                    ge=None, ge_advanced=None) -> "B200Prompt":
        return B200Prompt(self, ref_seq, ref_bert, ssl_content, ref_audio_32k, sv_emb, ge, ge_advanced)

    # -- T2S --------------------------------------------------------------------
    def t2s_generate(self, prompts: Sequence["B200Prompt"], text_seqs: Sequence[np.ndarray],
                     text_berts: Optional[Sequence[Optional[np.ndarray]]] = None,
                     sampling: Optional[SamplingParams] = None, cancel_flag: Optional[C.c_int] = None
                     ) -> Tuple[List[np.ndarray], List[int]]:
        """Returns (y_full per utterance: prompt tokens + every generated token, idx per utterance)."""
        sp = sampling or SamplingParams()
        B = len(prompts)
        seqs = [_i64(t) for t in text_seqs]
        lens = np.asarray([len(t) for t in seqs], dtype=np.int32)
        cat = np.concatenate(seqs)
        bert = None
        if text_berts is not None and any(b is not None and np.any(b) for b in text_berts):
            bert = np.concatenate([_f32(b) if b is not None else np.zeros((len(s), 1024), np.float32)
                                   for b, s in zip(text_berts, seqs)], axis=0)
        steps = sp.fixed_steps if sp.fixed_steps > 0 else (sp.max_steps if sp.max_steps > 0 else 500)
        y_ld = max(p.n_prompt_tokens for p in prompts) + steps + 2
        y = np.zeros((B, y_ld), dtype=np.int64)
        y_len = np.zeros(B, dtype=np.int32)
        idx = np.zeros(B, dtype=np.int32)
        hs = (C.c_void_p * B)(*[p._h for p in prompts])
        csp = sp.to_c()
        rc = N.lib().genie_t2s_generate(self._h, hs, B, _ptr(cat), _ptr(lens), _ptr(bert), C.byref(csp),
                                        C.cast(C.pointer(cancel_flag), C.c_void_p) if cancel_flag is not None else None,
                                        0, _ptr(y), y_ld, _ptr(y_len), _ptr(idx))
        if rc == N.CANCELLED:
            return [], []
        N.check(rc)
        return [y[b, :y_len[b]].copy() for b in range(B)], [int(i) for i in idx]

    # -- T2S in pieces: prefill / decode_steps / read (streaming, polling between chunks of steps) --------
    def t2s_prefill(self, prompts: Sequence["B200Prompt"], text_seqs: Sequence[np.ndarray],
                    text_berts: Optional[Sequence[Optional[np.ndarray]]] = None,
                    sampling: Optional[SamplingParams] = None) -> None:
        """Encoder + first-stage pass for the batch; it stays in flight inside the model until the next prefill."""
        sp = sampling or SamplingParams()
        B = len(prompts)
        seqs = [_i64(t) for t in text_seqs]
        lens = np.asarray([len(t) for t in seqs], dtype=np.int32)
        cat = np.concatenate(seqs)
        bert = None
        if text_berts is not None and any(b is not None and np.any(b) for b in text_berts):
            bert = np.concatenate([_f32(b) if b is not None else np.zeros((len(s), 1024), np.float32)
                                   for b, s in zip(text_berts, seqs)], axis=0)
        steps = sp.fixed_steps if sp.fixed_steps > 0 else (sp.max_steps if sp.max_steps > 0 else 500)
        self._inflight = (B, max(p.n_prompt_tokens for p in prompts) + steps + 2)
        hs = (C.c_void_p * B)(*[p._h for p in prompts])
        csp = sp.to_c()
        N.check(N.lib().genie_t2s_prefill(self._h, hs, B, _ptr(cat), _ptr(lens), _ptr(bert), C.byref(csp), 0))

    def t2s_decode_steps(self, n_steps: int, cancel_flag: Optional[C.c_int] = None) -> Tuple[int, int, bool]:
        """Up to ``n_steps`` more decode steps; returns (utterances still decoding, steps so far, cancelled)."""
        n_active, done = C.c_int(0), C.c_int(0)
        rc = N.lib().genie_t2s_decode_steps(self._h, int(n_steps),
                                            C.cast(C.pointer(cancel_flag), C.c_void_p) if cancel_flag is not None else None,
                                            C.byref(n_active), C.byref(done))
        if rc != N.CANCELLED:
            N.check(rc)
        return n_active.value, done.value, rc == N.CANCELLED

    def t2s_read(self) -> Tuple[List[np.ndarray], List[int]]:
        """(y_full per utterance, idx per utterance) for what the batch in flight has generated so far."""
        B, y_ld = self._inflight
        y = np.zeros((B, y_ld), dtype=np.int64)
        y_len = np.zeros(B, dtype=np.int32)
        idx = np.zeros(B, dtype=np.int32)
        N.check(N.lib().genie_t2s_read(self._h, 0, _ptr(y), y_ld, _ptr(y_len), _ptr(idx)))
        return [y[b, :y_len[b]].copy() for b in range(B)], [int(i) for i in idx]

    # -- continuous batching: the T2S stage as a pool of decode slots ---------------------------------
    def pool_create(self, n_slots: int, kv_capacity: int, max_prompt_tokens: int, max_steps: int = 500) -> None:
        N.check(N.lib().genie_t2s_pool_create(self._h, n_slots, kv_capacity, max_prompt_tokens, max_steps))
        self._pool = (n_slots, max_prompt_tokens + max_steps + 2)

    def pool_admit(self, slots: Sequence[int], prompts: Sequence["B200Prompt"], text_seqs: Sequence[np.ndarray],
                   text_berts: Optional[Sequence[Optional[np.ndarray]]] = None,
                   samplings: Optional[Sequence[SamplingParams]] = None) -> None:
        """Prefill ``len(slots)`` requests into free slots while the other slots keep their state."""
        n = len(slots)
        seqs = [_i64(t) for t in text_seqs]
        lens = np.asarray([len(t) for t in seqs], dtype=np.int32)
        cat = np.concatenate(seqs)
        bert = None
        if text_berts is not None and any(b is not None and np.any(b) for b in text_berts):
            bert = np.concatenate([_f32(b) if b is not None else np.zeros((len(s), 1024), np.float32)
                                   for b, s in zip(text_berts, seqs)], axis=0)
        sl = np.asarray(slots, dtype=np.int32)
        hs = (C.c_void_p * n)(*[p._h for p in prompts])
        sps = None
        if samplings is not None:
            sps = (N.Sampling * n)(*[(sp or SamplingParams()).to_c() for sp in samplings])
        N.check(N.lib().genie_t2s_admit(self._h, n, _ptr(sl), hs, _ptr(cat), _ptr(lens), _ptr(bert),
                                        C.cast(sps, C.c_void_p) if sps is not None else None))

    def pool_step(self, n_steps: int) -> int:
        """Up to ``n_steps`` decode steps over all active slots; returns the number still decoding."""
        n_active = C.c_int(0)
        N.check(N.lib().genie_t2s_pool_step(self._h, int(n_steps), C.byref(n_active)))
        return n_active.value

    def pool_poll(self) -> Tuple[np.ndarray, np.ndarray]:
        """(state[n_slots]: 0 free / 1 decoding / 2 finished, tokens generated so far per slot)."""
        n = self._pool[0]
        state = np.zeros(n, dtype=np.int32)
        gen = np.zeros(n, dtype=np.int32)
        N.check(N.lib().genie_t2s_pool_poll(self._h, _ptr(state), _ptr(gen), n))
        return state, gen

    def pool_read(self, slot: int) -> Tuple[np.ndarray, int]:
        y = np.zeros(self._pool[1], dtype=np.int64)
        y_len, idx = C.c_int(0), C.c_int(0)
        N.check(N.lib().genie_t2s_pool_read(self._h, int(slot), _ptr(y), len(y), C.byref(y_len), C.byref(idx)))
        return y[:y_len.value].copy(), idx.value

    def pool_release(self, slot: int) -> None:
        N.check(N.lib().genie_t2s_release(self._h, int(slot)))

    # -- SoVITS -----------------------------------------------------------------
    def vits_decode(self, prompts: Sequence["B200Prompt"], text_seqs: Sequence[np.ndarray],
                    semantic: Sequence[np.ndarray], zp_noise: Optional[Sequence[np.ndarray]] = None,
                    seed: Optional[int] = None, noise_scale: float = -1.0,
                    noise_ids: Optional[Sequence[int]] = None) -> List[np.ndarray]:
        """``seed=None``: fresh z_p noise per call, as the reference's unseeded RandomNormalLike (vits#[6490])."""
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        B = len(prompts)
        seqs = [_i64(t) for t in text_seqs]
        sems = [_i64(t) for t in semantic]
        tl = np.asarray([len(t) for t in seqs], dtype=np.int32)
        sl = np.asarray([len(t) for t in sems], dtype=np.int32)
        for s_ in sems:
            if len(s_) == 0 or s_.max() >= 1024 or s_.min() < 0:
                raise ValueError("semantic tokens must be non-empty ids in [0, 1024)")
        noise = None
        if zp_noise is not None:
            noise = np.concatenate([_f32(np.asarray(z).reshape(192, -1)[:, :2 * len(s_)]).reshape(-1)
                                    for z, s_ in zip(zp_noise, sems)])
        # one fresh buffer per call (every sample is written by the library); the per-utterance results are
        # views into it, so the 1280 samples per token cross the host memory bus once
        audio = _pinned.take(int(sl.sum()) * 1280)       # page-locked: the device-to-host copy runs at DMA speed
        alen = np.zeros(B, dtype=np.int32)
        hs = (C.c_void_p * B)(*[p._h for p in prompts])
        seq_cat, sem_cat = np.concatenate(seqs), np.concatenate(sems)   # keep alive across the call
        ids = np.ascontiguousarray(noise_ids, dtype=np.int32) if noise_ids is not None else None
        N.check(N.lib().genie_vits_decode(self._h, hs, B, _ptr(seq_cat), _ptr(tl),
                                          _ptr(sem_cat), _ptr(sl), _ptr(noise), seed & (2 ** 64 - 1),
                                          noise_scale, 0, _ptr(audio), _ptr(alen), _ptr(ids)))
        out, o = [], 0
        for b in range(B):
            out.append(audio[o:o + alen[b]])
            o += alen[b]
        return out

    # -- device-resident payload variants (bench `value` leg: inputs/outputs stay in HBM) ------
    def t2s_generate_device(self, prompts, text_seq_cat, text_lens: np.ndarray, sampling: SamplingParams, y_out,
                            text_bert_cat=None):
        """text_seq_cat: int64 CUDA tensor (concat); text_bert_cat: optional f32 CUDA tensor [sum len, 1024];
        y_out: int64 CUDA tensor [B, y_ld].  Lengths and the small per-utterance results (y_len, idx) are host
        metadata."""
        B = len(prompts)
        lens = np.ascontiguousarray(text_lens, dtype=np.int32)
        y_len = np.zeros(B, dtype=np.int32)
        idx = np.zeros(B, dtype=np.int32)
        hs = (C.c_void_p * B)(*[p._h for p in prompts])
        csp = sampling.to_c()
        N.check(N.lib().genie_t2s_generate(self._h, hs, B, _ptr(text_seq_cat), _ptr(lens), _ptr(text_bert_cat),
                                           C.byref(csp), None, 1, _ptr(y_out), int(y_out.shape[1]), _ptr(y_len),
                                           _ptr(idx)))
        return y_len, idx

    def t2s_prefill_device(self, prompts, text_seq_cat, text_lens: np.ndarray, sampling: SamplingParams,
                           text_bert_cat=None) -> None:
        """Device-resident form of ``t2s_prefill`` (inputs are CUDA tensors)."""
        B = len(prompts)
        lens = np.ascontiguousarray(text_lens, dtype=np.int32)
        hs = (C.c_void_p * B)(*[p._h for p in prompts])
        csp = sampling.to_c()
        steps = sampling.fixed_steps if sampling.fixed_steps > 0 else (sampling.max_steps if sampling.max_steps > 0 else 500)
        self._inflight = (B, max(p.n_prompt_tokens for p in prompts) + steps + 2)
        N.check(N.lib().genie_t2s_prefill(self._h, hs, B, _ptr(text_seq_cat), _ptr(lens), _ptr(text_bert_cat),
                                          C.byref(csp), 1))

    def t2s_read_device(self, y_out) -> Tuple[np.ndarray, np.ndarray]:
        """Device-resident form of ``t2s_read``: tokens into the CUDA tensor ``y_out`` [B, y_ld]; (y_len, idx) on the host."""
        B = self._inflight[0]
        y_len = np.zeros(B, dtype=np.int32)
        idx = np.zeros(B, dtype=np.int32)
        N.check(N.lib().genie_t2s_read(self._h, 1, _ptr(y_out), int(y_out.shape[1]), _ptr(y_len), _ptr(idx)))
        return y_len, idx

    def vits_decode_device(self, prompts, text_seq_cat, text_lens: np.ndarray, sem_cat, sem_lens: np.ndarray,
                           audio_out, seed: int = 0, noise_ids=None) -> np.ndarray:
        B = len(prompts)
        tl = np.ascontiguousarray(text_lens, dtype=np.int32)
        sl = np.ascontiguousarray(sem_lens, dtype=np.int32)
        alen = np.zeros(B, dtype=np.int32)
        hs = (C.c_void_p * B)(*[p._h for p in prompts])
        ids = np.ascontiguousarray(noise_ids, dtype=np.int32) if noise_ids is not None else None
        N.check(N.lib().genie_vits_decode(self._h, hs, B, _ptr(text_seq_cat), _ptr(tl), _ptr(sem_cat), _ptr(sl), None,
                                          seed & (2 ** 64 - 1), -1.0, 1, _ptr(audio_out), _ptr(alen), _ptr(ids)))
        return alen

    # -- debug ------------------------------------------------------------------
    def record_logits(self, enable: bool) -> None:
        N.check(N.lib().genie_debug_record_logits(self._h, int(enable)))

    def read_logits(self) -> np.ndarray:
        n = C.c_int(0)
        N.check(N.lib().genie_debug_read_logits(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.float32)
        N.check(N.lib().genie_debug_read_logits(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def keep(self, enable: bool) -> None:
        N.check(N.lib().genie_debug_keep(self._h, int(enable)))

    def debug_sample(self, logits: np.ndarray, hist: np.ndarray, hist_len: np.ndarray, sampling: SamplingParams,
                     noise: Optional[np.ndarray] = None, n_draws: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """The sampler kernel on host-supplied logits [rows,1025] and histories [rows,ld]; returns
        (tokens [rows,n_draws], stop flags [rows,n_draws]).  ``noise`` [n_draws,rows,1025] replaces Philox."""
        lg = _f32(logits).reshape(-1, 1025)
        rows = lg.shape[0]
        h = np.ascontiguousarray(hist, dtype=np.int64).reshape(rows, -1)
        hl = np.ascontiguousarray(hist_len, dtype=np.int32).reshape(rows)
        nz = _f32(noise).reshape(n_draws, rows, 1025) if noise is not None else None
        tok = np.zeros((rows, n_draws), dtype=np.int64)
        stop = np.zeros((rows, n_draws), dtype=np.int32)
        csp = sampling.to_c()
        N.check(N.lib().genie_debug_sample(self._h, _ptr(lg), rows, _ptr(h), h.shape[1], _ptr(hl), C.byref(csp),
                                           _ptr(nz), n_draws, _ptr(tok), _ptr(stop)))
        return tok, stop

    def read_kept(self, what: str) -> np.ndarray:
        n = C.c_longlong(0)
        N.check(N.lib().genie_debug_read(self._h, what.encode(), None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.float32)
        N.check(N.lib().genie_debug_read(self._h, what.encode(), _ptr(out), n.value, C.byref(n)))
        return out

    def last_timing(self) -> dict:
        t = np.zeros(12, dtype=np.float32)
        N.check(N.lib().genie_last_timing(self._h, _ptr(t), 12))
        return {"prefill_ms": float(t[0]), "decode_ms": float(t[1]), "t2s_ms": float(t[2]), "steps": int(t[3]),
                "vits_ms": float(t[4]), "generator_ms": float(t[5]), "generator_launches": int(t[6]),
                "latent_rows": int(t[7]), "decode_attention_us": float(t[8]), "decode_attention_kv_mb": float(t[9]),
                "narrow_conv_ms": float(t[10]), "narrow_conv_mb": float(t[11])}


class B200Prompt:
    def __init__(self, model: B200Model, ref_seq, ref_bert, ssl_content, ref_audio_32k=None, sv_emb=None,
                 ge=None, ge_advanced=None):
        self.model = model
        self._h = C.c_void_p(0)
        seq = _i64(ref_seq)
        bert = None
        if ref_bert is not None and np.any(ref_bert):
            bert = _f32(ref_bert).reshape(len(seq), 1024)
        ssl = _f32(ssl_content).reshape(768, -1)
        L = N.lib()
        if ge is not None:
            g = _f32(ge).reshape(-1)
            ga = _f32(ge_advanced).reshape(-1) if ge_advanced is not None else None
            N.check(L.genie_prompt_create_with_ge(model._h, _ptr(seq), len(seq), _ptr(bert), _ptr(ssl), ssl.shape[1],
                                                  _ptr(g), len(g), _ptr(ga), C.byref(self._h)))
        else:
            au = _f32(ref_audio_32k).reshape(-1)
            sv = _f32(sv_emb).reshape(-1) if sv_emb is not None else None
            N.check(L.genie_prompt_create(model._h, _ptr(seq), len(seq), _ptr(bert), _ptr(ssl), ssl.shape[1],
                                          _ptr(au), len(au), _ptr(sv), C.byref(self._h)))
        n, gd, lr = C.c_int(0), C.c_int(0), C.c_int(0)
        N.check(L.genie_prompt_info(self._h, C.byref(n), C.byref(gd), C.byref(lr)))
        self.n_prompt_tokens, self.ge_dim, self.ref_len = n.value, gd.value, lr.value
        model._prompts.add(self)                 # the model closes its prompts before it closes itself

    @property
    def closed(self) -> bool:
        return not self._h

    def read(self):
        pr = np.zeros(self.n_prompt_tokens, dtype=np.int64)
        ge = np.zeros(self.ge_dim, dtype=np.float32)
        gea = np.zeros(512, dtype=np.float32)
        N.check(N.lib().genie_prompt_read(self._h, _ptr(pr), _ptr(ge), _ptr(gea)))
        return pr, ge, gea

    def close(self) -> None:
        if self._h:
            N.lib().genie_prompt_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
