import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs
m = B200Model(fixture_dir("v2", 0))
what = sys.argv[1]
if what == "t2s":
    prs = [make_prompt_inputs(seed=100 + i, Lr=12 + 5 * (i % 3), Ts=40 + 8 * (i % 4), n_audio=32000) for i in range(2)]
    prompts = [m.make_prompt(p["ref_seq"], p["ref_bert"], p["ssl_content"], p["ref_audio"]) for p in prs]
    txs = [make_text_inputs(seed=200 + i, Lt=9 + 3 * (i % 5)) for i in range(10)]
    sp = SamplingParams(greedy=True, max_steps=10)
    for ug in (1, 0):
        m.set_option("use_graph", ug)
        for B in (10, 1, 3, 1):
            ys, idx = m.t2s_generate([prompts[i % 2] for i in range(B)], [t["text_seq"] for t in txs[:B]], None, sp)
            print("graph", ug, "B", B, "len0", len(ys[0]), "Ly", prompts[0].n_prompt_tokens, "idx", idx[:3], m.last_timing()["steps"])
else:
    B = int(sys.argv[2]); T = int(sys.argv[3])
    pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
    prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
    rng = np.random.default_rng(0)
    txs = [make_text_inputs(seed=200 + i, Lt=int(rng.integers(40, 61))) for i in range(B)]
    sems = [rng.integers(0, 1024, T - (i % 3)) for i in range(B)]
    a = m.vits_decode([prompt] * B, [t["text_seq"] for t in txs], sems)
    print("vits ok", B, T, len(a), a[0][:4], m.last_timing())
