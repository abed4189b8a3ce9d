"""Extract the architecture constants baked into the reference's T2S graphs
(not weights: they live in Constant nodes) into the package data file
genie-tts_b200/genie_tts/Data/t2s_constants.json.  Build container only.

  * pe_div_term: t2s_stage_decoder_fp32.onnx `/ar_audio_position/Constant_1_output_0`
    (256 fp32; the exporter's exp() is not correctly rounded, so the table is
    kept bit-exact instead of being recomputed)
  * top_k 15 (`onnx::Reshape_2129`), repetition penalty 1.35 (`/Constant_13_output_0`),
    temperature 1.0 (`/Constant_15_output_0`), EOS 1024 (`/Constant_20_output_0`)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "genie-tts_b200"))
from genie_tts.onnx_reader import load_model  # noqa: E402

REF = "/root/reference/src/genie_tts/Data/v2/Models/"


def const(model, out_name):
    for n in model.graph.nodes:
        if n.op_type == "Constant" and n.outputs == [out_name]:
            return n.attrs["value"].t.numpy()
    raise KeyError(out_name)


def main():
    st = load_model(REF + "t2s_stage_decoder_fp32.onnx")
    fs = load_model(REF + "t2s_first_stage_decoder_fp32.onnx")
    en = load_model(REF + "t2s_encoder_fp32.onnx")
    div = const(st, "/ar_audio_position/Constant_1_output_0")
    assert np.array_equal(div, const(fs, "/ar_audio_position/Constant_1_output_0"))
    assert np.array_equal(div, const(en, "/encoder/ar_text_position/Constant_1_output_0"))
    out = {
        "pe_div_term_f32_hex": div.astype("<f4").tobytes().hex(),
        "top_k": int(const(st, "onnx::Reshape_2129")),
        "repetition_penalty": float(const(st, "/Constant_13_output_0")),
        "temperature": float(const(st, "/Constant_15_output_0")),
        "eos": int(const(st, "/Constant_20_output_0")),
        "max_decode_steps": 500,   # src/genie_tts/Core/Inference.py:95
        "vits_noise_scale": 0.5,   # vits_fp32.onnx#[6494]
    }
    dst = os.path.join(ROOT, "genie-tts_b200", "genie_tts", "Data")
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(dst, "t2s_constants.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: (v if not isinstance(v, str) else v[:16] + "...") for k, v in out.items()})


if __name__ == "__main__":
    main()
