import os, sys, time
import numpy as np
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs
m = B200Model(fixture_dir("v2", 0))
pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
rng = np.random.default_rng(0)
for B in (100, 50, 16):
    seqs = [make_text_inputs(seed=200 + i, Lt=int(rng.integers(40, 61)))["text_seq"] for i in range(B)]
    sp = SamplingParams(seed=3, max_steps=90, fixed_steps=90)
    for rep in range(3):
        m.t2s_prefill([prompt] * B, seqs, None, sp)
        m.t2s_read()
        t0 = time.perf_counter()
        m.t2s_decode_steps(90)
        t1 = time.perf_counter()
        m.t2s_read()
        t2 = time.perf_counter()
    print(f"B={B}: host enqueue of 90 graph launches {1e3*(t1-t0):.1f} ms, until done {1e3*(t2-t0):.1f} ms, device decode {m.last_timing()['decode_ms']:.1f} ms")
