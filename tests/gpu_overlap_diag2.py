"""Does the decode of two contexts overlap at all?  Decode-only timing (prefill excluded) for several batch sizes
and launch modes: one context alone vs two contexts at the same time."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs

m = B200Model(fixture_dir("v2", 0))
ctxs = m.pipeline_contexts(2)
pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
rng = np.random.default_rng(0)
sp = SamplingParams(seed=3, max_steps=90, fixed_steps=90)


def decode_only(k, B, seqs, out):
    ctxs[k].t2s_prefill([prompt] * B, seqs, None, sp)
    ctxs[k].t2s_read()
    bar.wait()
    t0 = time.perf_counter()
    ctxs[k].t2s_decode_steps(90)
    ctxs[k].t2s_read()
    out[k] = 1000 * (time.perf_counter() - t0)


for graph in (1, 0):
    for c in ctxs:
        c.set_option("use_graph", graph)
    for B in (100, 16, 6):
        seqs = [make_text_inputs(seed=200 + i, Lt=int(rng.integers(40, 61)))["text_seq"] for i in range(B)]
        res = {}
        for n in (1, 2):
            best = None
            for rep in range(3):
                bar = threading.Barrier(n)
                out = {}
                th = [threading.Thread(target=decode_only, args=(k, B, seqs, out)) for k in range(n)]
                [t.start() for t in th]
                [t.join() for t in th]
                v = max(out.values())
                best = v if best is None else min(best, v)
            res[n] = best
        print(f"use_graph={graph} B={B}: decode alone {res[1]:.1f} ms, two contexts at once {res[2]:.1f} ms (x{res[2] / res[1]:.2f})")
