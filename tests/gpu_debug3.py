import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs
m = B200Model(fixture_dir("v2", 0))
steps = 10
sp = SamplingParams(greedy=True, max_steps=steps)
prs = [make_prompt_inputs(seed=100 + i, Lr=12 + 5 * (i % 3), Ts=40 + 8 * (i % 4), n_audio=32000 + 6400 * (i % 2), bert=(i % 2 == 0)) for i in range(4)]
prompts = [m.make_prompt(p["ref_seq"], p["ref_bert"], p["ssl_content"], p["ref_audio"]) for p in prs]
B = 10
txs = [make_text_inputs(seed=200 + i, Lt=9 + 3 * (i % 5), bert=(i % 3 == 0)) for i in range(B)]
pid = [i % 4 for i in range(B)]
ys, idx = m.t2s_generate([prompts[p] for p in pid], [t["text_seq"] for t in txs], [t["text_bert"] for t in txs], sp)
print("batch lens", [len(y) for y in ys], idx)
for ug in (1, 0):
    m.set_option("use_graph", ug)
    for b in (0, 3, 7, 9, 3):
        y1, i1 = m.t2s_generate([prompts[pid[b]]], [txs[b]["text_seq"]], [txs[b]["text_bert"]], sp)
        print("graph", ug, "b", b, "len", len(y1[0]), "Ly", prompts[pid[b]].n_prompt_tokens, "idx", i1, "eq", np.array_equal(y1[0], ys[b]), m.last_timing()["steps"])
