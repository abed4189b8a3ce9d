"""Config-5 shape without the HTTP layer: N concurrent synthetic requests (JA20-shaped sentences, 90-token budget)
against ONE replica through genie_tts.Scheduler.BatchScheduler; prints request latency p50 / p99 (submit ->
waveform ready), batch statistics and the sustained audio-seconds per second."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.Core.Inference import GENIE
from genie_tts.Scheduler import BatchScheduler
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
MAXB = int(sys.argv[2]) if len(sys.argv) > 2 else 256
CLIENTS = 32
m = B200Model(fixture_dir("v2", 0))
pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
rng = np.random.default_rng(0)
txs = [make_text_inputs(seed=300 + i, Lt=int(rng.integers(40, 61)))["text_seq"] for i in range(N)]
sp = SamplingParams(seed=7, max_steps=90, fixed_steps=90)
g = GENIE()
for rounds in range(2):                                   # round 0 warms graphs / workspaces
    sch = BatchScheduler(m, synthesizer=g, sampling=sp, max_batch=MAXB, max_wait_ms=5.0)
    futs = [None] * N
    t0 = time.perf_counter()

    def client(k):
        for i in range(k, N, CLIENTS):
            futs[i] = sch.submit(prompt, txs[i])

    th = [threading.Thread(target=client, args=(k,)) for k in range(CLIENTS)]
    [t.start() for t in th]
    [t.join() for t in th]
    auds = [f.result(timeout=600) for f in futs]
    dt = time.perf_counter() - t0
    st = sch.stats.summary()
    sch.close()
audio_s = sum(len(a) for a in auds) / 32000.0
print(f"{N} concurrent requests, max_batch {MAXB}: {st['batches']} batches (mean {st['mean_batch']:.0f}), "
      f"latency p50 {st['latency_ms_p50']:.0f} ms p99 {st['latency_ms_p99']:.0f} ms, "
      f"{audio_s / dt:.0f} audio-s/s sustained ({audio_s:.0f} s of audio in {dt:.2f} s)")
