"""Dump the initialiser tables (name, dims, dtype, offset, length, bin) of the
reference's six graph templates into tests/golden/model_schema.json.

Run in the build container only (needs /root/reference).  The schema is what
the fixture writer (tests/fixture_models.py) uses to emit a model directory in
the converter's on-disk layout (reference: src/genie_tts/Converter/v2/
T2SConverter.py:45-79, VITSConverter.py:44-81, EncoderConverter.py:38-106,
v2ProPlus/PromptEncoderConverter.py:43-87) when the graph templates themselves
are not available (GPU box).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "genie-tts_b200"))
from genie_tts.onnx_reader import load_model  # noqa: E402

REF = "/root/reference/src/genie_tts/Data"
GRAPHS = {
    "v2": ["t2s_encoder_fp32", "t2s_first_stage_decoder_fp32", "t2s_stage_decoder_fp32", "vits_fp32"],
    "v2ProPlus": ["vits_fp32", "prompt_encoder_fp32"],
}


def main():
    out = {}
    for ver, names in GRAPHS.items():
        for nm in names:
            m = load_model(os.path.join(REF, ver, "Models", nm + ".onnx"), with_nodes=False)
            rows = []
            for t in m.graph.initializers:
                assert t.is_external, (nm, t.name)
                rows.append([t.name, list(t.dims), t.data_type,
                             int(t.external["offset"]), int(t.external["length"])])
            out[f"{ver}/{nm}"] = {
                "inputs": [[i.name, i.elem_type, [d if isinstance(d, int) else str(d) for d in i.shape]]
                           for i in m.graph.inputs],
                "outputs": [o.name for o in m.graph.outputs],
                "initializers": rows,
            }
    with open(os.path.join(ROOT, "tests", "golden", "model_schema.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print({k: len(v["initializers"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
