"""Random-init model directories in the converter's on-disk layout.

Layout facts (reference): ``t2s_shared_fp16.bin`` / ``vits_fp16.bin`` /
``prompt_encoder_fp16.bin`` hold the tensors back to back in key-file order as
fp16, and each initialiser's external_data ``offset``/``length`` are in fp32
byte units (src/genie_tts/Converter/v2/T2SConverter.py:45-79,
v2/VITSConverter.py:44-81, v2ProPlus/PromptEncoderConverter.py:43-87);
``t2s_encoder_fp32.bin`` is true fp32 (v2/EncoderConverter.py:38-106).

When the reference's graph templates are available (``oracle/_ref/graphs``,
copied there by ``__graft_entry__.build()`` in the build container, or
/root/reference itself) the fixture uses them verbatim, so the directory is
byte-compatible with ``ModelManager.load_character`` and the graph interpreter
oracle can execute it.  Otherwise a weights-only ``.onnx`` (initialiser table
only) is written from tests/golden/model_schema.json — enough for the product
loader, which never looks at nodes.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
from typing import Dict, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "genie-tts_b200"))
from genie_tts.onnx_reader import write_weights_only_model  # noqa: E402

SCHEMA_PATH = os.path.join(ROOT, "tests", "golden", "model_schema.json")
REF_DATA = "/root/reference/src/genie_tts/Data"
REF_GRAPHS = os.path.join(ROOT, "oracle", "_ref", "graphs")

_FILES = {
    # graph file -> (schema key per version, bin file, bin dtype)
    "t2s_encoder_fp32": ("t2s_encoder_fp32.bin", np.float32),
    "t2s_first_stage_decoder_fp32": ("t2s_shared_fp16.bin", np.float16),
    "t2s_stage_decoder_fp32": ("t2s_shared_fp16.bin", np.float16),
    "vits_fp32": ("vits_fp16.bin", np.float16),
    "prompt_encoder_fp32": ("prompt_encoder_fp16.bin", np.float16),
}


def template_path(version: str, graph: str) -> Optional[str]:
    """Where a graph template lives, or None.  T2S graphs are shared by both versions."""
    for ver in (version, "v2"):
        for base in (REF_GRAPHS, REF_DATA):
            p = os.path.join(base, ver, "Models" if base == REF_DATA else "", graph + ".onnx")
            if os.path.exists(p):
                return p
    return None


def have_templates(version: str = "v2") -> bool:
    graphs = ["t2s_encoder_fp32", "t2s_first_stage_decoder_fp32", "t2s_stage_decoder_fp32", "vits_fp32"]
    if version == "v2ProPlus":
        graphs.append("prompt_encoder_fp32")
    return all(template_path(version, g) for g in graphs)


def _schema() -> Dict:
    with open(SCHEMA_PATH) as f:
        return json.load(f)


def _init_tensor(rng: np.random.Generator, name: str, dims) -> np.ndarray:
    """Non-degenerate random init: every tensor is non-trivial (biases, LN
    affine and weight_g are perturbed) so a kernel that drops one is caught."""
    n = int(np.prod(dims)) if len(dims) else 1
    shape = tuple(dims)
    last = name.rsplit(".", 1)[-1]
    if name.endswith("alpha"):
        return np.full(shape, 1.0 + 0.1 * rng.standard_normal(), np.float32)
    if last in ("gamma",) or "norm" in name and last == "weight":
        return (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
    if last in ("beta", "bias") or name.endswith("in_proj_bias"):
        return (0.05 * rng.standard_normal(shape)).astype(np.float32)
    if last == "weight_g":
        # weight-normed convs: unit gain, except the second conv of every HiFi-GAN
        # residual pair (0.3) so 15 stacked residual adds keep the pre-tanh signal O(1)
        gain = 0.3 if ".convs2." in name else 1.0
        return (gain * (1.0 + 0.1 * rng.standard_normal(shape))).astype(np.float32)
    if "embedding" in name or "codebook" in name or "emb_rel" in name:
        return (0.5 * rng.standard_normal(shape)).astype(np.float32)
    if last == "weight" and len(shape) == 1:      # PReLU slope etc.
        return (0.25 + 0.05 * rng.standard_normal(shape)).astype(np.float32)
    if len(shape) >= 2:
        if ".ups." in name:                       # ConvTranspose1d weight [Cin, Cout, k]
            fan_in = shape[0] * shape[2] / max(1, _stride_of(name))
        else:
            fan_in = int(np.prod(shape[1:]))
        return (rng.standard_normal(shape) / np.sqrt(max(fan_in, 1))).astype(np.float32)
    return (0.1 * rng.standard_normal(n).reshape(shape)).astype(np.float32)


def _stride_of(name: str) -> int:
    for i, s in enumerate((10, 8, 2, 2, 2)):
        if f".ups.{i}." in name:
            return s
    return 1


EOS_BOOST = 1.6
RESIDUAL_BRANCH_GAIN = 0.15


def write_fixture(out_dir: str, version: str = "v2", seed: int = 0) -> str:
    """Write a complete character model directory; returns ``out_dir``.

    A ``version`` with the suffix "sharp" ("v2sharp", "v2ProPlussharp") is that fixture with two changes to the T2S
    weights that make greedy decoding a SENSITIVE parity probe; the suffix "sens" applies only the first (no EOS
    boost: long decodes that never stop).  (1) ``linear2.weight`` / ``out_proj.weight`` of every layer are scaled by
    RESIDUAL_BRANCH_GAIN: with plain N(0, 1/fan_in) init 24 post-LN ReLU layers sit deep in the ordered phase —
    the final hidden state (hence the logits) is the same to 2 decimals for every input token, step and sentence,
    so token equality says little about attention / KV indexing.  With small residual branches the residual
    stream keeps the token and position identity, logits change from step to step and a wrong cache row flips
    tokens.  (2) Row 1024 (EOS) of ``ar_predict_layer.weight`` is scaled by EOS_BOOST so the natural-stop path
    (stop flag = argmax(raw)==1024 or token==1024, stage#[1807-1821]; host slicing Inference.py:105-109) fires
    after a few to a few dozen steps, different for every sentence (incl. idx 0 and EOS as the first-stage token)."""
    sens = version.endswith(("sharp", "sens"))
    sharp = version.endswith("sharp")
    if sens:
        version = version[:-5] if sharp else version[:-4]
    os.makedirs(out_dir, exist_ok=True)
    schema = _schema()
    graphs = ["t2s_encoder_fp32", "t2s_first_stage_decoder_fp32", "t2s_stage_decoder_fp32", "vits_fp32"]
    if version == "v2ProPlus":
        graphs.append("prompt_encoder_fp32")
    written_bins: Dict[str, bool] = {}
    for gname in graphs:
        key = f"{version}/{gname}" if f"{version}/{gname}" in schema else f"v2/{gname}"
        rows = schema[key]["initializers"]
        bin_name, bin_dt = _FILES[gname]
        # graph file
        tpl = template_path(version, gname)
        dst = os.path.join(out_dir, gname + ".onnx")
        if tpl is not None:
            shutil.copyfile(tpl, dst)
        else:
            write_weights_only_model(dst, [(r[0], r[1], r[2], bin_name, r[3], r[4]) for r in rows],
                                     graph_name=gname)
        if bin_name in written_bins:
            continue
        written_bins[bin_name] = True
        # tensors back to back in table order (== key-file order; offsets verified cumulative)
        rng = np.random.default_rng([seed, sum(map(ord, bin_name))])
        total = max(r[3] + r[4] for r in rows)
        blob = np.zeros(total // 4, dtype=np.float32)
        off_expect = 0
        for (name, dims, dt, off, ln) in sorted(rows, key=lambda r: r[3]):
            assert dt == 1 and off == off_expect, (name, off, off_expect)
            arr = _init_tensor(rng, name, dims)
            assert arr.size * 4 == ln, (name, arr.shape, ln)
            if sharp and name == "ar_predict_layer.weight":
                arr[1024] *= EOS_BOOST
            if sens and name.startswith("transformer_encoder.") and name.endswith(("linear2.weight", "out_proj.weight")):
                arr *= RESIDUAL_BRANCH_GAIN
            blob[off // 4: (off + ln) // 4] = arr.reshape(-1)
            off_expect = off + ln
        with open(os.path.join(out_dir, bin_name), "wb") as f:
            f.write(blob.astype(bin_dt).tobytes())
    return out_dir


if __name__ == "__main__":
    print(write_fixture(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "v2",
                        int(sys.argv[3]) if len(sys.argv) > 3 else 0))
