"""Small profiling target: one T2S call (+ optional vocoder) at batch B for a few decode steps."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs
B = int(sys.argv[1]); steps = int(sys.argv[2]); vits = len(sys.argv) > 3 and sys.argv[3] == "vits"
m = B200Model(fixture_dir("v2", 0))
m.set_option("use_graph", int(os.environ.get("USE_GRAPH", "1")))
m.set_option("time_attention", int(os.environ.get("TIME_ATT", "0")))
if os.environ.get("BRANCHES"):
    m.set_option("decode_branches", int(os.environ["BRANCHES"]))
if os.environ.get("SPLIT"):
    m.set_option("decode_split_min", int(os.environ["SPLIT"]))
pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600)
prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"])
rng = np.random.default_rng(0)
txs = [make_text_inputs(seed=200 + i, Lt=int(rng.integers(40, 61))) for i in range(B)]
sp = SamplingParams(seed=3, max_steps=steps, fixed_steps=steps)
from genie_tts import _native as N
for it in range(2):
    if it == 1:
        N.lib().genie_profiler_range(1)
    ys, idx = m.t2s_generate([prompt] * B, [t["text_seq"] for t in txs], None, sp)
    print("t2s", m.last_timing())
    if vits:
        nt = int(os.environ.get("VITS_TOKENS", steps))
        sems = [rng.integers(0, 1024, nt).astype(np.int64) if nt != steps else y[-steps:] % 1024 for y in ys]
        a = m.vits_decode([prompt] * B, [t["text_seq"] for t in txs], sems)
        print("vits", m.last_timing())
N.lib().genie_profiler_range(0)
