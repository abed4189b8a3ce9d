"""Concurrency diagnostic: two execution contexts on one GPU running (a) decode + decode, (b) SoVITS + SoVITS,
(c) decode + SoVITS at the same time, against the same stages alone.  Prints wall ms per combination."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs

B, STEPS = 100, 90
m = B200Model(fixture_dir("v2", 0))
ctxs = m.pipeline_contexts(2)
if os.environ.get("PART"):
    for c in ctxs:
        c.set_option("sm_partition", int(os.environ["PART"]))
pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
rng = np.random.default_rng(0)
seqs = [make_text_inputs(seed=200 + i, Lt=int(rng.integers(40, 61)))["text_seq"] for i in range(B)]
sp = SamplingParams(seed=3, max_steps=STEPS, fixed_steps=STEPS)
sems = [rng.integers(0, 1024, STEPS).astype(np.int64) for _ in range(B)]


def t2s(k):
    ctxs[k].t2s_generate([prompt] * B, seqs, None, sp)


def vits(k):
    ctxs[k].vits_decode([prompt] * B, seqs, sems, seed=5)


def timed(jobs, reps=3):
    best = 1e9
    for _ in range(reps):
        th = [threading.Thread(target=f, args=(k,)) for f, k in jobs]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        best = min(best, 1000 * (time.perf_counter() - t0))
    return best


for k in (0, 1):
    t2s(k); vits(k)
print(f"PART={os.environ.get('PART')} CUDA_DEVICE_MAX_CONNECTIONS={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')} GENIE_STREAM_PRIO={os.environ.get('GENIE_STREAM_PRIO')}")
print(f"t2s alone            {timed([(t2s, 0)]):8.1f} ms")
print(f"vits alone           {timed([(vits, 0)]):8.1f} ms")
print(f"t2s  + t2s           {timed([(t2s, 0), (t2s, 1)]):8.1f} ms")
print(f"vits + vits          {timed([(vits, 0), (vits, 1)]):8.1f} ms")
print(f"t2s  + vits          {timed([(t2s, 0), (vits, 1)]):8.1f} ms")
