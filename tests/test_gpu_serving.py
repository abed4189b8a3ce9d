"""Continuous batching (C-ABI slot pool), execution contexts, handle lifetimes, the sampler in isolation and the
REST surface — all through the C-ABI on a B200."""
import json
import os
import threading
import time

import numpy as np
import pytest
import torch

from synth import make_prompt_inputs, make_text_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def v2(v2_dir):
    from genie_tts.engine import B200Model
    from oracle import gsv_port as P
    m = B200Model(v2_dir)
    yield m, P.PortModel(v2_dir)
    m.close()


def _prompt(m, pr):
    return m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"], pr.get("sv_emb"))


# ---------------------------------------------------------------------------------------------- sampler (K7 / K8)
def _penalised(raw, hist, penalty=1.35):
    lg = raw.astype(np.float32).copy()
    h = np.unique(hist)
    s = lg[h]
    lg[h] = np.where(s < 0, s * np.float32(penalty), s / np.float32(penalty))
    return lg


def test_sampler_matches_port_with_injected_noise_and_topk_membership(v2):
    """argmax(p / noise) semantics (negative draws, +-0 scores of masked tokens), top-k with ties, the stop flag and
    the top-p extension: the sampler kernel vs the port's sample_token on the SAME noise, row by row."""
    from genie_tts.engine import SamplingParams
    from oracle import gsv_port as P
    m, _ = v2
    rng = np.random.default_rng(5)
    rows, draws, ld = 48, 6, 40
    logits = rng.standard_normal((rows, 1025)).astype(np.float32) * 1.5
    logits[::7, 1024] += 6.0                                   # some rows: EOS is the raw arg-max (stop flag)
    logits[3, 10] = logits[3, 11] = logits[3].max() + 1.0       # an exact tie inside the top-k
    hist = rng.integers(0, 1024, (rows, ld)).astype(np.int64)
    hist_len = rng.integers(1, ld + 1, rows).astype(np.int32)
    noise = rng.standard_normal((draws, rows, 1025)).astype(np.float32)
    for kw in (dict(), dict(top_k=5, temperature=0.7, repetition_penalty=1.2), dict(top_p=0.8), dict(top_k=40, top_p=0.5)):
        tok, stop = m.debug_sample(logits, hist, hist_len, SamplingParams(seed=1, **kw), noise=noise, n_draws=draws)
        same = total = 0
        for r in range(rows):
            h = torch.as_tensor(hist[r, :hist_len[r]])
            for d in range(draws):
                t, s = P.sample_token(torch.as_tensor(logits[r]), h, torch.as_tensor(noise[d, r]),
                                      top_k=kw.get("top_k", 15), temperature=kw.get("temperature", 1.0),
                                      penalty=kw.get("repetition_penalty", 1.35), top_p=kw.get("top_p", 1.0))
                total += 1
                same += int(t == tok[r, d] and bool(s) == bool(stop[r, d]))
        assert same >= total - 2, (kw, same, total)            # float ties between expf implementations at most
    # greedy == arg-max of the penalised logits; stop flag as the graph computes it
    tok, stop = m.debug_sample(logits, hist, hist_len, SamplingParams(greedy=True))
    for r in range(rows):
        lg = _penalised(logits[r], hist[r, :hist_len[r]])
        assert tok[r, 0] == int(np.argmax(lg))
        assert bool(stop[r, 0]) == (int(np.argmax(logits[r])) == 1024 or tok[r, 0] == 1024)


def test_sampler_philox_distribution_chi2(v2):
    """Philox sampling draws from the same distribution as argmax(p / N(0,1)) of the reference graphs (which is NOT
    multinomial): two-sample chi-square of the kernel's draws against a numpy Monte-Carlo of the port's rule."""
    from genie_tts.engine import SamplingParams
    m, _ = v2
    rng = np.random.default_rng(9)
    raw = (rng.standard_normal(1025) * 1.2).astype(np.float32)
    hist = rng.integers(0, 1024, (1, 30)).astype(np.int64)
    rows, draws = 1024, 32
    tok, _ = m.debug_sample(np.repeat(raw[None], rows, 0), np.repeat(hist, rows, 0), np.full(rows, 30, np.int32),
                            SamplingParams(seed=77), n_draws=draws)
    tok = tok.reshape(-1)
    lg = _penalised(raw, hist[0])
    kth = np.sort(lg)[-15]
    p = np.where(lg < kth, 0.0, np.exp(lg - lg.max())).astype(np.float32)
    p /= p.sum()
    top = np.nonzero(p > 0)[0]
    assert np.isin(tok, top).mean() > 0.999                    # inside the top-k set (all-negative draws: 2^-15)
    n = len(tok)
    q = np.random.default_rng(10).standard_normal((n, len(top))).astype(np.float32)
    mc = top[np.argmax(p[top][None, :] / q, axis=1)]
    a = np.asarray([(tok == t).sum() for t in top], np.float64)
    b = np.asarray([(mc == t).sum() for t in top], np.float64)
    keep = (a + b) >= 10
    chi2 = float((((a - b) ** 2) / np.maximum(a + b, 1))[keep].sum())
    assert chi2 < 50.0, (chi2, a.tolist(), b.tolist())          # df <= 14: P(chi2 > 50) ~ 1e-5
    assert a.max() / n < 0.9                                   # not degenerate


def test_sampled_decode_tokens_inside_topk_and_fresh_seeds(v2):
    """Sampling inside the real decode loop: every sampled token belongs to the top-k set of that step's penalised
    logits; an explicit seed reproduces, seed=None (the default, like the reference's unseeded graphs) does not."""
    from genie_tts.engine import SamplingParams
    m, _ = v2
    pr, tx = make_prompt_inputs(seed=51, Lr=10, Ts=20, n_audio=32000), make_text_inputs(seed=52, Lt=8)
    prompt = _prompt(m, pr)
    try:
        m.record_logits(True)
        ys, _ = m.t2s_generate([prompt], [tx["text_seq"]], None, SamplingParams(seed=7, max_steps=24, fixed_steps=24))
        lg = m.read_logits().reshape(-1, 1025)
        m.record_logits(False)
        y, Ly = ys[0], prompt.n_prompt_tokens
        assert len(y) == Ly + 25 and len(lg) == 25
        for t in range(25):
            pen = _penalised(lg[t], y[:Ly + t])
            assert pen[y[Ly + t]] >= np.sort(pen)[-15], t
        a, _ = m.t2s_generate([prompt], [tx["text_seq"]], None, SamplingParams(seed=7, max_steps=24, fixed_steps=24))
        assert np.array_equal(a[0], y)                          # graph replay, same seed
        b, _ = m.t2s_generate([prompt], [tx["text_seq"]], None, SamplingParams(max_steps=24, fixed_steps=24))
        c, _ = m.t2s_generate([prompt], [tx["text_seq"]], None, SamplingParams(max_steps=24, fixed_steps=24))
        assert not np.array_equal(b[0], c[0])                   # fresh noise per call
    finally:
        m.record_logits(False)
        prompt.close()


# ---------------------------------------------------------------------------------------------- slot pool / contexts
def _drive_pool(ctx, plan, prompts, txs, sp):
    """Admit per ``plan`` [(slots, request indices)], decoding a few steps between admissions; returns {request: (y, idx)}."""
    owner, done = {}, {}

    def collect():
        state, _ = ctx.pool_poll()
        for sl in [s for s in list(owner) if state[s] == 2]:
            done[owner.pop(sl)] = ctx.pool_read(sl)
            ctx.pool_release(sl)

    for slots, reqs in plan:
        ctx.pool_admit(slots, [prompts[i % 2] for i in reqs], [txs[i]["text_seq"] for i in reqs],
                       [txs[i]["text_bert"] for i in reqs], [sp] * len(reqs))
        owner.update(dict(zip(slots, reqs)))
        ctx.pool_step(3)                                    # others keep decoding between admissions
        collect()
    for _ in range(60):
        if not owner:
            break
        ctx.pool_step(8)
        collect()
    return done


def test_slot_pool_admission_while_decoding_matches_batch_api(v2):
    """Requests admitted into arbitrary free slots at different times, sharing decode steps with requests that are
    mid-way, must produce exactly the tokens the batch API gives each of them; released slots are reused.
    Pass 1: fp32 cache rows on both sides (pool vs each request alone through the batch API, a different kernel
    path) — exact.  Pass 2: the default fp16 rows (pool vs all 14 in one batch call) — exact."""
    from genie_tts.engine import SamplingParams
    m, _ = v2
    prs = [make_prompt_inputs(seed=1200 + i, Lr=12 + 6 * i, Ts=40 + 16 * i, n_audio=32000, bert=(i == 1)) for i in range(2)]
    prompts = [_prompt(m, p) for p in prs]
    txs = [make_text_inputs(seed=1300 + i, Lt=8 + 3 * (i % 6), bert=(i % 4 == 0)) for i in range(14)]
    sp = SamplingParams(greedy=True, max_steps=40)
    plan = [([17, 0, 5], [0, 1, 2]), ([1, 2, 23, 9], [3, 4, 5, 6]), ([3], [7]), ([4, 6, 7, 8, 10, 11], [8, 9, 10, 11, 12, 13])]
    ctxs = []
    try:
        for kv16 in (0, 1):
            m.set_option("kv_fp16", kv16)
            ctx = m.create_context()                            # inherits the option
            ctxs.append(ctx)
            if kv16 == 0:
                ref = [m.t2s_generate([prompts[i % 2]], [txs[i]["text_seq"]], [txs[i]["text_bert"]], sp) for i in range(14)]
                ref = [(r[0][0], r[1][0]) for r in ref]
            else:
                ys, idx = m.t2s_generate([prompts[i % 2] for i in range(14)], [t["text_seq"] for t in txs],
                                         [t["text_bert"] for t in txs], sp)
                ref = list(zip(ys, idx))
            ctx.pool_create(n_slots=24, kv_capacity=256, max_prompt_tokens=64, max_steps=40)
            done = _drive_pool(ctx, plan, prompts, txs, sp)
            assert len(done) == 14
            for i in range(14):
                assert np.array_equal(done[i][0], ref[i][0]) and done[i][1] == ref[i][1], (kv16, i)
        m.set_option("kv_fp16", 1)
        ctx = ctxs[-1]
        # reuse of released slots, seeded sampling per slot, budget instead of stop flags: 5 requests admitted
        # together decode exactly like the batch API's 5 (same seed -> same Philox keys per utterance index)
        sps = SamplingParams(seed=99, max_steps=12, fixed_steps=12)
        r5, i5 = m.t2s_generate([prompts[0]] * 5, [txs[i]["text_seq"] for i in range(5)], None, sps)
        ctx.pool_admit([17, 3, 9, 0, 21], [prompts[0]] * 5, [txs[i]["text_seq"] for i in range(5)], None, [sps] * 5)
        while ctx.pool_step(8):
            pass
        for k, sl in enumerate([17, 3, 9, 0, 21]):
            y, idx = ctx.pool_read(sl)
            assert np.array_equal(y, r5[k]) and idx == i5[k]
            ctx.pool_release(sl)
        with pytest.raises(Exception):
            ctx.pool_read(17)                                   # slot not in use
        with pytest.raises(Exception):                          # does not fit a slot
            ctx.pool_admit([0], [prompts[0]], [np.arange(300) % 700], None, [sp])
        state, _ = ctx.pool_poll()
        assert (state == 0).all()
    finally:
        m.set_option("kv_fp16", 1)
        for p in prompts:
            p.close()
        for c in ctxs:
            c.close()


def test_two_contexts_decode_concurrently_from_two_threads(v2):
    """>1 session per model: two execution contexts on the same weights, driven from two host threads at the same
    time (one on a caller-owned torch stream), each reproduce the single-handle tokens."""
    from genie_tts.engine import SamplingParams
    m, _ = v2
    pr = make_prompt_inputs(seed=1400, Lr=16, Ts=48, n_audio=32000)
    prompt = _prompt(m, pr)
    txs = [make_text_inputs(seed=1410 + i, Lt=9 + i) for i in range(12)]
    sp = SamplingParams(greedy=True, max_steps=30)
    user_stream = torch.cuda.Stream()
    ctxs = [m.create_context(), m.create_context(cuda_stream=user_stream.cuda_stream)]
    try:
        ref = m.t2s_generate([prompt] * 12, [t["text_seq"] for t in txs], None, sp)
        out, err = [None, None], []

        def work(k):
            try:
                for _ in range(3):
                    out[k] = ctxs[k].t2s_generate([prompt] * 12, [t["text_seq"] for t in txs], None, sp)
            except Exception as e:                              # noqa: BLE001
                err.append(e)

        th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not err, err
        for k in range(2):
            assert all(np.array_equal(a, b) for a, b in zip(out[k][0], ref[0])) and out[k][1] == ref[1]
    finally:
        for c in ctxs:
            c.close()
        prompt.close()


def test_tts_batch_stream_pipelined_equals_sequential(v2):
    """GENIE.tts_batch_stream keeps two batches in flight on two execution contexts (T2S of one overlaps SoVITS of
    the other, bulk stages on a lower-priority stream): every batch must come back in order with exactly the
    waveforms the sequential call produces (seeded sampling, seeded z_p noise)."""
    from genie_tts.Core.Inference import GENIE
    from genie_tts.engine import SamplingParams
    m, _ = v2
    g = GENIE()
    pr = make_prompt_inputs(seed=1800, Lr=20, Ts=64, n_audio=32000)
    prompt = _prompt(m, pr)
    sp = SamplingParams(seed=21, max_steps=20, fixed_steps=20)
    batches = []
    for k in range(5):
        n = 9 + 3 * k
        txs = [make_text_inputs(seed=1810 + 40 * k + i, Lt=8 + (i % 7)) for i in range(n)]
        batches.append(([prompt] * n, [t["text_seq"] for t in txs], None))
    try:
        ref = [g.tts_batch(m, *b, sampling=sp) for b in batches]
        for depth in (2, 3):
            got = list(g.tts_batch_stream(m, batches, sampling=sp, depth=depth))
            assert len(got) == len(ref)
            for a, b in zip(got, ref):
                assert len(a) == len(b)
                for x, y in zip(a, b):
                    assert x.shape == y.shape and np.array_equal(x, y)
    finally:
        prompt.close()


def test_model_and_prompt_lifetimes(v2_dir):
    """closing a model while prompts built on it are alive must not leave dangling handles: the model
    closes its prompts (their HBM goes with it), later prompt.close() / __del__ are no-ops, and a prompt destroyed
    after its model at the C level is harmless."""
    import ctypes as C
    from genie_tts import _native as N
    from genie_tts.engine import B200Model, SamplingParams
    m = B200Model(v2_dir)
    pr = make_prompt_inputs(seed=1500, Lr=10, Ts=20, n_audio=32000)
    p1, p2 = _prompt(m, pr), _prompt(m, pr)
    ys, _ = m.t2s_generate([p1], [make_text_inputs(seed=1501, Lt=8)["text_seq"]], None, SamplingParams(greedy=True, max_steps=4))
    assert len(ys[0]) > 0
    raw = C.c_void_p(p2._h.value)
    p2._h = C.c_void_p(0)              # take p2 out of the Python registry: destroyed by hand AFTER the model
    m._prompts.discard(p2)
    m.close()
    assert p1.closed and m.closed
    p1.close()                         # idempotent
    N.lib().genie_prompt_destroy(raw)  # C level: prompt outlives its model
    m2 = B200Model(v2_dir)             # the device is still healthy
    p3 = _prompt(m2, pr)
    with pytest.raises(N.GenieNativeError):
        m2.t2s_generate([p3], [np.asarray([3, 5000])], None, SamplingParams(greedy=True, max_steps=2))   # id out of range
    ys2, _ = m2.t2s_generate([p3], [make_text_inputs(seed=1501, Lt=8)["text_seq"]], None, SamplingParams(greedy=True, max_steps=4))
    assert np.array_equal(ys[0], ys2[0])
    m2.close()


def test_second_device_in_one_process(v2_dir):
    """kernels needing > 48 KB of dynamic shared memory must launch on every device of the process."""
    from genie_tts import _native as N
    from genie_tts.engine import B200Model, SamplingParams
    if N.lib().genie_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pr, tx = make_prompt_inputs(seed=1600, Lr=10, Ts=20, n_audio=32000), make_text_inputs(seed=1601, Lt=8)
    outs = []
    for dev in (0, 1):
        m = B200Model(v2_dir, device=dev)
        p = _prompt(m, pr)
        ys, _ = m.t2s_generate([p] * 12, [tx["text_seq"]] * 12, None, SamplingParams(greedy=True, max_steps=6))
        aud = m.vits_decode([p], [tx["text_seq"]], [ys[0][-5:] % 1024], seed=3)[0]
        outs.append((ys[0], aud))
        m.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.abs(outs[0][1] - outs[1][1]).max() < 1e-5


# ---------------------------------------------------------------------------------------------- REST surface
def test_rest_concurrent_tts_requests_are_isolated():
    """Config-5 behaviour behind the unchanged REST surface: concurrent /tts requests (each several sentences) are
    served by continuous batching and every request receives exactly its own audio, in sentence order — the
    reference's second concurrent request would clear the first one's queues (TTSPlayer.py:185-186)."""
    import requests
    import uvicorn
    from genie_tts import Internal, Server
    from genie_tts.Audio.ReferenceAudio import ReferenceAudio
    from genie_tts.GetPhonesAndBert import set_text_frontend
    from genie_tts.Service import SynthesisService

    def frontend(text, language):                               # synthetic G2P: length and ids from the text
        body = text.lstrip("。")
        n = 6 + (sum(map(ord, body)) % 9)
        rng = np.random.default_rng(sum(map(ord, body)))
        seq = rng.integers(0, 732, (1, n)).astype(np.int64)
        seq[0, 0] = 3
        return seq, None

    set_text_frontend(frontend)
    # two scheduler contexts on the one GPU (own slot pools / streams / graphs, shared weights): requests are spread
    # over both by queued phonemes, results must not depend on which one served a sentence
    svc = SynthesisService(devices=[0], n_slots=32, kv_capacity=320, max_prompt_tokens=64, max_steps=40,
                           contexts_per_gpu=2)
    Server.set_service(svc)
    pr = make_prompt_inputs(seed=1700, Lr=14, Ts=48, n_audio=32000)
    ref = ReferenceAudio.from_features("synthetic-ref", pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
    from conftest import fixture_dir
    svc.load_character("Mika", fixture_dir("v2sharp", 0), "Japanese")   # natural stops: every sentence its own length
    Internal.set_reference_features("Mika", ref)
    svc.set_reference("Mika", ref)
    cfg = uvicorn.Config(Server.app, host="127.0.0.1", port=18765, log_level="warning")
    server = uvicorn.Server(cfg)
    th = threading.Thread(target=server.run, daemon=True)
    th.start()
    for _ in range(100):
        if server.started:
            break
        time.sleep(0.05)
    try:
        # greedy + a small step budget keep the test short; per-request seeds are irrelevant under greedy, so the
        # expected audio LENGTH of a request is what its sentences give when synthesised alone
        from genie_tts.engine import SamplingParams
        svc.pool_cfg["sampling"] = SamplingParams(greedy=True, max_steps=40)
        texts = [f"文{i}あ。文{i}い、長い文です。文{i}う!" for i in range(12)]

        def expected_len(text):
            st = svc.submit("Mika", text, True)
            return [len(b) for b in st.chunks(timeout=120)]

        want = [expected_len(t) for t in texts]
        assert all(len(w) >= 2 for w in want)                   # really several sentences per request
        got, err = [None] * len(texts), []

        def client(i):
            try:
                r = requests.post("http://127.0.0.1:18765/tts", json={"character_name": "Mika", "text": texts[i],
                                                                      "split_sentence": True}, stream=True, timeout=120)
                assert r.status_code == 200
                got[i] = b"".join(r.iter_content(chunk_size=None))
            except Exception as e:                              # noqa: BLE001
                err.append(e)

        cl = [threading.Thread(target=client, args=(i,)) for i in range(len(texts))]
        [c.start() for c in cl]
        [c.join() for c in cl]
        assert not err, err
        for i in range(len(texts)):
            assert len(got[i]) == sum(want[i]) and len(got[i]) % (1280 * 2) == 0, i
        assert len({len(g) for g in got}) > 1                   # different requests, different audio
        r = requests.post("http://127.0.0.1:18765/tts", json={"character_name": "Nobody", "text": "x"}, timeout=30)
        assert r.status_code == 404
        assert requests.post("http://127.0.0.1:18765/stop", timeout=30).status_code == 200
        st = requests.get("http://127.0.0.1:18765/stats", timeout=30).json()
        assert len(st["mika"]) == 2 and all(r["requests"] > 0 for r in st["mika"])
        assert sum(r["requests"] for r in st["mika"]) >= 12 * 2
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, "rest_isolation_stats.json"), "w") as f:
                json.dump(st, f, indent=1)
    finally:
        server.should_exit = True
        th.join(timeout=10)
        svc.unload_character("Mika")
        svc.close()
        Server.set_service(None)
        set_text_frontend(None)
