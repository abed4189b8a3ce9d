"""Generate tests/golden/*.npz by executing the REFERENCE's graph files
(/root/reference/src/genie_tts/Data/**/Models/*.onnx) with oracle/onnx_interp.py
over the reference host loop restatement (oracle/ref_pipeline.py) on seeded
fixtures.  Build container only; the vectors pin oracle/gsv_port.py and the
CUDA path where /root/reference is absent (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "genie-tts_b200")]
from conftest import fixture_dir  # noqa: E402
from oracle import ref_pipeline as R  # noqa: E402
from synth import make_prompt_inputs, make_text_inputs, make_zp_noise  # noqa: E402

CASES = {
    # name: (version, fixture seed, prompt kw, text kw, max_steps)
    "v2_small": ("v2", 0, dict(seed=11, Lr=20, Ts=60, n_audio=64000, bert=True), dict(seed=12, Lt=15, bert=True), 12),
    "v2_ja20": ("v2", 0, dict(seed=21, Lr=60, Ts=264, n_audio=169600), dict(seed=22, Lt=50), 24),
    "v2pp_small": ("v2ProPlus", 1, dict(seed=31, Lr=20, Ts=60, n_audio=64000, v2pp=True), dict(seed=32, Lt=15), 12),
}


def main():
    for name, (ver, fseed, pkw, tkw, steps) in CASES.items():
        d = fixture_dir(ver, fseed)
        s = R.load_sessions(d)
        pr = make_prompt_inputs(**pkw)
        tx = make_text_inputs(**tkw)
        zp = make_zp_noise(pkw["seed"] + 100, steps + 2)
        R.set_sampler_mode(s, greedy=True, zp_noise=zp)
        for prog in (s.first_stage, s.stage):
            prog.trace = {}
            prog.keep = {"/ar_predict_layer/MatMul_output_0"}
        col = {}
        toks = R.t2s_cpu(s, pr["ref_seq"], pr["ref_bert"], tx["text_seq"], tx["text_bert"], pr["ssl_content"],
                         max_steps=steps, collect=col)
        out = {
            "prompts": col["prompts"][0], "y_full": col["y_full"][0], "idx": np.int64(col["idx"]),
            "tokens": toks, "x": col["x"][0].astype(np.float32),
            "logits_first": s.first_stage.trace["/ar_predict_layer/MatMul_output_0"].numpy().reshape(-1),
            "logits_last": s.stage.trace["/ar_predict_layer/MatMul_output_0"].numpy().reshape(-1),
            "k0_first_stage": col["kv0"][0][:, 0, :].astype(np.float32),
        }
        sem = R.strip_eos(toks)
        if s.prompt_encoder is None:
            audio = R.vocode(s, tx["text_seq"], sem, ref_audio_32k=pr["ref_audio"])
        else:
            ge, gea = R.prompt_global_emb(s, pr["ref_audio"], pr["sv_emb"])
            out["ge"], out["ge_advanced"] = ge.reshape(-1), gea.reshape(-1)
            audio = R.vocode(s, tx["text_seq"], sem, ge=ge, ge_advanced=gea)
        out["audio"] = audio.astype(np.float32)
        out["semantic"] = sem.reshape(-1)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
