"""Dynamic batching in front of one model replica (first step).

The reference serves one sentence at a time from a single worker thread
(``Core/TTSPlayer.py:74-120``: ``tts_client.tts`` per sentence).  On a B200 a decode step costs
the same ~1.3 ms for 1 or 100 utterances (it is latency-bound), so the server-side
win is to run whatever is waiting as ONE ``GENIE.tts_batch`` call.  ``BatchScheduler`` owns the
replica's worker thread: requests are queued, the worker takes everything that arrives within
``max_wait_ms`` of the first waiting request (up to ``max_batch``) and resolves one future per
request.  ``ReplicaPool`` is the per-box front: one scheduler per GPU replica, least-loaded
dispatch by queued phoneme count — utterances are independent, no collective (§8e).

Joining requests into a batch that is already decoding (true continuous batching) needs a
resumable ``genie_t2s_generate`` and is the next step; this module fixes the interface it will
sit behind.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np


@dataclass
class _Request:
    prompt: Any
    text_seq: np.ndarray
    text_bert: Optional[np.ndarray]
    future: Future
    t_submit: float = field(default_factory=time.perf_counter)


@dataclass
class SchedulerStats:
    batches: int = 0
    requests: int = 0
    max_batch: int = 0
    busy_s: float = 0.0
    latency_ms: List[float] = field(default_factory=list)

    def summary(self) -> dict:
        lat = np.asarray(self.latency_ms, dtype=np.float64)
        return {"batches": self.batches, "requests": self.requests, "max_batch": self.max_batch,
                "mean_batch": self.requests / self.batches if self.batches else 0.0, "busy_s": self.busy_s,
                "latency_ms_p50": float(np.percentile(lat, 50)) if lat.size else None,
                "latency_ms_p99": float(np.percentile(lat, 99)) if lat.size else None}


class BatchScheduler:
    """One worker thread per model replica; ``submit`` is thread-safe and returns a Future of the
    float32 waveform (same semantics as ``GENIE.tts`` for that sentence)."""

    _STOP = object()

    def __init__(self, model, synthesizer=None, sampling=None, max_batch: int = 128, max_wait_ms: float = 2.0,
                 name: str = "replica0"):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        if synthesizer is None:
            from .Core.Inference import tts_client
            synthesizer = tts_client
        self.model, self.synth, self.sampling = model, synthesizer, sampling
        self.max_batch, self.max_wait = int(max_batch), max_wait_ms / 1000.0
        self.stats = SchedulerStats()
        self._q: "queue.Queue" = queue.Queue()
        self._load = 0                       # queued + running phonemes (dispatch cost estimate)
        self._lock = threading.Lock()
        self._closed = False
        self._thread = threading.Thread(target=self._loop, name=f"genie-batcher-{name}", daemon=True)
        self._thread.start()

    # ---- client side -------------------------------------------------------------------
    @property
    def load(self) -> int:
        return self._load

    def submit(self, prompt, text_seq, text_bert=None) -> Future:
        if self._closed:
            raise RuntimeError("scheduler is closed")
        seq = np.ascontiguousarray(np.asarray(text_seq).reshape(-1), dtype=np.int64)
        if seq.size == 0:
            raise ValueError("empty phoneme sequence")
        fut: Future = Future()
        with self._lock:
            self._load += int(seq.size)
        self._q.put(_Request(prompt, seq, None if text_bert is None else np.asarray(text_bert, dtype=np.float32), fut))
        return fut

    def close(self, wait: bool = True) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(self._STOP)
        if wait:
            self._thread.join()

    # ---- worker ------------------------------------------------------------------------
    def _gather(self) -> Optional[List[_Request]]:
        first = self._q.get()
        if first is self._STOP:
            return None
        batch = [first]
        deadline = time.perf_counter() + self.max_wait
        while len(batch) < self.max_batch:
            remaining = deadline - time.perf_counter()
            try:
                nxt = self._q.get(timeout=remaining) if remaining > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if nxt is self._STOP:
                self._q.put(self._STOP)          # finish this batch, stop on the next turn
                break
            batch.append(nxt)
        return batch

    def _run(self, batch: List[_Request]) -> None:
        berts: Optional[List[np.ndarray]] = None
        if any(r.text_bert is not None for r in batch):       # zero rows == "no BERT features" (GetPhonesAndBert.py:60)
            berts = [r.text_bert if r.text_bert is not None else np.zeros((r.text_seq.size, 1024), np.float32)
                     for r in batch]
        t0 = time.perf_counter()
        try:
            kwargs = {"sampling": self.sampling} if self.sampling is not None else {}
            auds = self.synth.tts_batch(self.model, [r.prompt for r in batch], [r.text_seq for r in batch], berts, **kwargs)
            if len(auds) != len(batch):
                raise RuntimeError(f"tts_batch returned {len(auds)} waveforms for {len(batch)} requests")
            for r, a in zip(batch, auds):
                r.future.set_result(a)
        except BaseException as e:                             # the stream must survive a bad batch (TTSPlayer.py:109-114)
            for r in batch:
                if not r.future.done():
                    r.future.set_exception(e)
        t1 = time.perf_counter()
        st = self.stats
        st.batches += 1
        st.requests += len(batch)
        st.max_batch = max(st.max_batch, len(batch))
        st.busy_s += t1 - t0
        st.latency_ms.extend(1000.0 * (t1 - r.t_submit) for r in batch)
        with self._lock:
            self._load -= sum(int(r.text_seq.size) for r in batch)

    def _loop(self) -> None:
        while True:
            batch = self._gather()
            if batch is None:
                break
            self._run(batch)
        # fail whatever was queued behind the stop marker
        while True:
            try:
                r = self._q.get_nowait()
            except queue.Empty:
                break
            if r is not self._STOP and not r.future.done():
                r.future.set_exception(RuntimeError("scheduler closed"))


class ReplicaPool:
    """One ``BatchScheduler`` per GPU replica; requests go to the replica with the least queued work."""

    def __init__(self, schedulers: Sequence[BatchScheduler]):
        if not schedulers:
            raise ValueError("need at least one replica")
        self.schedulers = list(schedulers)
        self._lock = threading.Lock()

    def submit(self, prompts: Sequence[Any], text_seq, text_bert=None) -> Future:
        """``prompts[i]`` is the request's prompt handle on replica i (handles are per device, §8b)."""
        if len(prompts) != len(self.schedulers):
            raise ValueError("one prompt handle per replica")
        with self._lock:
            i = min(range(len(self.schedulers)), key=lambda j: self.schedulers[j].load)
            return self.schedulers[i].submit(prompts[i], text_seq, text_bert)

    def close(self) -> None:
        for s in self.schedulers:
            s.close()
