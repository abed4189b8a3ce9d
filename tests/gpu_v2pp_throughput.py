"""Informational: stage timings of the bench workload shape on the V2ProPlus fixture (1.5x wider generator)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir
from genie_tts.engine import B200Model, SamplingParams
from synth import make_prompt_inputs, make_text_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ver = sys.argv[2] if len(sys.argv) > 2 else "v2ProPlus"
m = B200Model(fixture_dir(ver, 1 if ver == "v2ProPlus" else 0))
pr = make_prompt_inputs(seed=1, Lr=60, Ts=264, n_audio=169600, v2pp=(ver == "v2ProPlus"))
prompt = m.make_prompt(pr["ref_seq"], pr["ref_bert"], pr["ssl_content"], pr["ref_audio"], pr.get("sv_emb"))
rng = np.random.default_rng(0)
txs = [make_text_inputs(seed=200 + i, Lt=int(rng.integers(40, 61))) for i in range(B)]
sp = SamplingParams(seed=3, max_steps=90, fixed_steps=90)
for it in range(3):
    ys, idx = m.t2s_generate([prompt] * B, [t["text_seq"] for t in txs], None, sp)
    t = m.last_timing()
    sems = [y[-90:] % 1024 for y in ys]
    a = m.vits_decode([prompt] * B, [t_["text_seq"] for t_ in txs], sems)
    t2 = m.last_timing()
audio_s = sum(len(x) for x in a) / 32000.0
tot = t["t2s_ms"] + t2["vits_ms"]
print(f"{ver} B={B}: prefill {t['prefill_ms']:.1f} decode {t['decode_ms']:.1f} vits {t2['vits_ms']:.1f} "
      f"(generator {t2['generator_ms']:.1f}) ms -> {audio_s / (tot * 1e-3):.0f} audio-s/s (stage sum)")
