"""Request scheduling in front of one model replica (BASELINE config 5).

The reference serves one sentence at a time from a single worker thread
(``Core/TTSPlayer.py:74-120``: ``tts_client.tts`` per sentence) and a second concurrent ``/tts``
clears the first one's queues (``TTSPlayer.py:185-186``).  On a B200 a decode step costs about the
same for 1 or 100 utterances (it is latency-bound), so the server-side win is to keep
as many utterances as possible inside every step:

* ``ContinuousBatcher`` — the serving scheduler.  One thread per (GPU replica, character) drives the
  C-ABI slot pool (``genie_t2s_pool_*``): waiting requests are ADMITTED into free decode slots (their
  prefill runs as one ragged batch) while the other slots are mid-decode, every tick advances all
  active slots a few steps, finished slots are read, released and vocoded in small batches.  A
  request never waits for a batch to drain; per-request results do not depend on what shares the pool.
* ``BatchScheduler`` — window batching for offline lists: whatever arrives within ``max_wait_ms``
  becomes ONE ``GENIE.tts_batch`` call.
* ``ReplicaPool`` — the per-box front: one scheduler per GPU replica, least-loaded dispatch by
  queued phoneme count — utterances are independent, no collective (§8e).
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
    sampling: Any = None
    generation: int = 0


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


@dataclass
class _SlotJob:
    req: "_Request"
    slot: int
    t_admit: float


class ContinuousBatcher:
    """Continuous batching on the C-ABI slot pool of ONE model handle (one GPU replica of one character).

    ``submit`` is thread-safe and returns a Future of the float32 waveform of that sentence — the same
    result ``GENIE.tts`` gives for it (reference slicing quirks and EOS strip included).  The scheduler thread
    loops: admit -> a few decode steps -> collect finished slots -> vocode what is ready.
    """

    _STOP = object()

    def __init__(self, model, n_slots: int = 128, kv_capacity: int = 1024, max_prompt_tokens: int = 512,
                 max_steps: int = 500, steps_per_tick: int = 8, max_admit: int = 32, vits_max_batch: int = 32,
                 vits_window_ms: float = 4.0, sampling=None, name: str = "replica0"):
        from .engine import SamplingParams
        self.model = model
        self.n_slots, self.kv_capacity, self.max_steps = int(n_slots), int(kv_capacity), int(max_steps)
        self.max_prompt_tokens = int(max_prompt_tokens)
        self.steps_per_tick, self.max_admit = int(steps_per_tick), int(max_admit)
        self.vits_max_batch, self.vits_window = int(vits_max_batch), vits_window_ms / 1000.0
        self.sampling = sampling or SamplingParams(max_steps=max_steps)
        self.stats = SchedulerStats()
        self.first_token_ms: List[float] = []
        self._q: "queue.Queue" = queue.Queue()
        self._load = 0
        self._lock = threading.Lock()
        self._closed = False
        self._generation = 0                  # bumped by cancel_all(): older requests are dropped
        self._ready = threading.Event()
        self._error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._loop, name=f"genie-cbatch-{name}", daemon=True)
        self._thread.start()
        self._ready.wait()
        if self._error is not None:
            raise self._error

    # ---- client side -------------------------------------------------------------------
    @property
    def load(self) -> int:
        return self._load

    def submit(self, prompt, text_seq, text_bert=None, sampling=None) -> Future:
        if self._closed:
            raise RuntimeError("scheduler is closed")
        seq = np.ascontiguousarray(np.asarray(text_seq).reshape(-1), dtype=np.int64)
        if seq.size == 0:
            raise ValueError("empty phoneme sequence")
        fut: Future = Future()
        r = _Request(prompt, seq, None if text_bert is None else np.asarray(text_bert, dtype=np.float32), fut)
        r.sampling = sampling
        r.generation = self._generation
        with self._lock:
            self._load += int(seq.size)
        self._q.put(r)
        return fut

    def cancel_all(self) -> None:
        """Drop everything queued or decoding (the reference's ``stop``: TTSPlayer.py:208-224)."""
        self._generation += 1

    def close(self, wait: bool = True) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(self._STOP)
        if wait:
            self._thread.join()

    # ---- scheduler thread ----------------------------------------------------------------
    def _fail(self, r: "_Request", e: BaseException) -> None:
        if not r.future.done():
            r.future.set_exception(e)
        with self._lock:
            self._load -= int(r.text_seq.size)

    def _fits(self, r: "_Request", steps: int) -> Optional[str]:
        need = r.prompt.ref_len + int(r.text_seq.size) + r.prompt.n_prompt_tokens + steps + 1
        if need > self.kv_capacity:
            return f"request needs {need} KV rows, slots hold {self.kv_capacity}"
        if r.prompt.n_prompt_tokens > self.max_prompt_tokens:
            return "reference audio longer than the pool was sized for"
        if r.text_seq.min() < 0 or r.text_seq.max() >= 732:
            return "phoneme id out of range"
        return None

    def _loop(self) -> None:
        from .Core.Inference import finish_t2s, strip_eos
        from .engine import SamplingParams
        m = self.model
        try:
            m.pool_create(self.n_slots, self.kv_capacity, self.max_prompt_tokens, self.max_steps)
        except BaseException as e:
            self._error = e
            self._ready.set()
            return
        self._ready.set()
        free = list(range(self.n_slots - 1, -1, -1))        # pop() hands out the lowest slot first
        active: dict = {}                                   # slot -> _SlotJob
        vits_wait: List[tuple] = []                         # (request, semantic tokens, t_ready)
        stopping = False
        while True:
            # ---- 1. admission: everything waiting, up to the free slots / max_admit per tick
            new: List[_Request] = []
            block = not active and not vits_wait and not stopping
            while free and len(new) < self.max_admit and len(new) < len(free):
                try:
                    r = self._q.get(timeout=0.05) if (block and not new) else self._q.get_nowait()
                except queue.Empty:
                    break
                if r is self._STOP:
                    stopping = True
                    break
                if r.generation != self._generation:
                    self._fail(r, RuntimeError("cancelled"))
                    continue
                sp = r.sampling or self.sampling
                steps = sp.fixed_steps if sp.fixed_steps > 0 else (sp.max_steps if sp.max_steps > 0 else 500)
                why = self._fits(r, min(steps, self.max_steps))
                if why or steps > self.max_steps:
                    self._fail(r, ValueError(why or f"max_steps {steps} exceeds the pool's {self.max_steps}"))
                    continue
                new.append(r)
            if new:
                slots = [free.pop() for _ in new]
                try:
                    berts = None
                    if any(r.text_bert is not None for r in new):
                        berts = [r.text_bert for r in new]
                    m.pool_admit(slots, [r.prompt for r in new], [r.text_seq for r in new], berts,
                                 [r.sampling or self.sampling for r in new])
                    now = time.perf_counter()
                    for r, sl in zip(new, slots):
                        active[sl] = _SlotJob(r, sl, now)
                        self.first_token_ms.append(1000.0 * (now - r.t_submit))
                except BaseException as e:                  # a bad admission must not take the pool down
                    for r, sl in zip(new, slots):
                        self._fail(r, e)
                        try:
                            m.pool_release(sl)
                        except Exception:
                            pass
                        free.append(sl)
                    free.sort(reverse=True)
            # ---- 2. cancellation of what is decoding
            for sl in [s_ for s_, j in active.items() if j.req.generation != self._generation]:
                j = active.pop(sl)
                m.pool_release(sl)
                free.append(sl)
                self._fail(j.req, RuntimeError("cancelled"))
            # ---- 3. a few decode steps for every active slot, then collect the finished ones
            if active:
                t0 = time.perf_counter()
                try:
                    m.pool_step(self.steps_per_tick)
                    state, _ = m.pool_poll()
                    for sl in [s_ for s_ in list(active) if state[s_] == 2]:
                        j = active.pop(sl)
                        y, idx = m.pool_read(sl)
                        m.pool_release(sl)
                        free.append(sl)
                        sem = strip_eos(finish_t2s(y, idx)).reshape(-1)
                        vits_wait.append((j.req, sem, time.perf_counter()))
                except BaseException as e:                  # device-side failure: fail what was decoding, rebuild
                    for sl, j in list(active.items()):
                        self._fail(j.req, e)
                    active.clear()
                    try:
                        m.pool_create(self.n_slots, self.kv_capacity, self.max_prompt_tokens, self.max_steps)
                    except BaseException:
                        stopping = True
                    free = list(range(self.n_slots - 1, -1, -1))
                free.sort(reverse=True)
                self.stats.busy_s += time.perf_counter() - t0
            # ---- 4. vocode: a full batch, or whatever has waited vits_window, or everything when idle
            if vits_wait:
                oldest = time.perf_counter() - vits_wait[0][2]
                if len(vits_wait) >= self.vits_max_batch or oldest >= self.vits_window or not active:
                    batch, vits_wait = vits_wait[:self.vits_max_batch], vits_wait[self.vits_max_batch:]
                    self._vocode(batch)
            if stopping and not active and not vits_wait:
                break
        while True:                                         # fail whatever was queued behind the stop marker
            try:
                r = self._q.get_nowait()
            except queue.Empty:
                break
            if r is not self._STOP:
                self._fail(r, RuntimeError("scheduler closed"))

    def _vocode(self, batch: List[tuple]) -> None:
        t0 = time.perf_counter()
        keep = [(r, sem) for r, sem, _ in batch if len(sem) > 0 and r.generation == self._generation]
        for r, sem, _ in batch:
            if len(sem) == 0 and not r.future.done():
                r.future.set_result(np.zeros(0, np.float32))
            elif r.generation != self._generation and not r.future.done():
                r.future.set_exception(RuntimeError("cancelled"))
        try:
            if keep:
                auds = self.model.vits_decode([r.prompt for r, _ in keep], [r.text_seq for r, _ in keep],
                                              [np.where(sem >= 1024, 0, sem) for _, sem in keep])
                for (r, _), a in zip(keep, auds):
                    r.future.set_result(a)
        except BaseException as e:
            for r, _ in keep:
                if not r.future.done():
                    r.future.set_exception(e)
        t1 = time.perf_counter()
        st = self.stats
        st.batches += 1
        st.requests += len(batch)
        st.max_batch = max(st.max_batch, len(batch))
        st.busy_s += t1 - t0
        st.latency_ms.extend(1000.0 * (t1 - r.t_submit) for r, _, _ in batch)
        with self._lock:
            self._load -= sum(int(r.text_seq.size) for r, _, _ in batch)


class ReplicaPool:
    """One scheduler (``ContinuousBatcher`` or ``BatchScheduler``) per GPU replica; a request goes to the replica
    with the least queued work."""

    def __init__(self, schedulers: Sequence[Any]):
        if not schedulers:
            raise ValueError("need at least one replica")
        self.schedulers = list(schedulers)
        self._lock = threading.Lock()

    def pick(self) -> int:
        with self._lock:
            return min(range(len(self.schedulers)), key=lambda j: self.schedulers[j].load)

    def submit(self, prompts: Sequence[Any], text_seq, text_bert=None, **kw) -> Future:
        """``prompts[i]`` is the request's prompt handle on replica i (handles are per device, §8b)."""
        if len(prompts) != len(self.schedulers):
            raise ValueError("one prompt handle per replica")
        with self._lock:
            i = min(range(len(self.schedulers)), key=lambda j: self.schedulers[j].load)
            return self.schedulers[i].submit(prompts[i], text_seq, text_bert, **kw)

    def close(self) -> None:
        for s in self.schedulers:
            s.close()
