"""Model-directory weight tables.

A converted character directory is ``*.onnx`` graphs whose initialisers point
into side ``.bin`` files (reference: src/genie_tts/ModelManager.py:59-114,
231-310).  ``external_data`` offset/length are in fp32 byte units even for the
fp16 bins (the reference up-casts the whole blob and slices it,
ModelManager.py:75-103), so the fp16 element index is ``offset/4``.  This module
only builds name -> view tables over the memory-mapped bins; nothing is
up-cast on the host — the device loader consumes the fp16 payload as stored.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from .onnx_reader import load_model

T2S_ENCODER = "t2s_encoder_fp32.onnx"
T2S_FIRST_STAGE = "t2s_first_stage_decoder_fp32.onnx"
T2S_STAGE = "t2s_stage_decoder_fp32.onnx"
T2S_BIN = "t2s_shared_fp16.bin"
VITS = "vits_fp32.onnx"
VITS_BIN = "vits_fp16.bin"
PROMPT_ENCODER = "prompt_encoder_fp32.onnx"
PROMPT_ENCODER_BIN = "prompt_encoder_fp16.bin"


@dataclass
class WeightTable:
    """name -> ndarray view (fp16 for the fp16 bins, fp32 for the encoder bin)."""
    tensors: Dict[str, np.ndarray]

    def __getitem__(self, k: str) -> np.ndarray:
        return self.tensors[k]

    def __contains__(self, k: str) -> bool:
        return k in self.tensors

    def f32(self, k: str) -> np.ndarray:
        return np.asarray(self.tensors[k], dtype=np.float32)


def read_table(onnx_path: str, fp16_bin: Optional[str]) -> WeightTable:
    if not os.path.exists(onnx_path):
        raise FileNotFoundError(f"ONNX Model not found: {onnx_path}")
    if fp16_bin is not None and not os.path.exists(fp16_bin):
        raise FileNotFoundError(f"FP16 Weight file not found: {fp16_bin}")
    model = load_model(onnx_path, with_nodes=False)
    base = os.path.dirname(onnx_path)
    blob16 = np.memmap(fp16_bin, dtype=np.float16, mode="r") if fp16_bin else None
    blobs32: Dict[str, np.ndarray] = {}
    out: Dict[str, np.ndarray] = {}
    for t in model.graph.initializers:
        if not t.is_external:
            if t.raw is not None:
                out[t.name] = t.numpy()
            continue
        off = int(t.external.get("offset", 0))
        ln = int(t.external.get("length", 0))
        if blob16 is not None:
            if (off + ln) // 4 > blob16.shape[0]:
                raise ValueError(f"tensor {t.name} exceeds {fp16_bin}")
            out[t.name] = blob16[off // 4:(off + ln) // 4].reshape(t.dims)
        else:
            loc = t.external["location"]
            if loc not in blobs32:
                p = os.path.join(base, loc)
                if not os.path.exists(p):
                    raise FileNotFoundError(f"weight file not found: {p}")
                blobs32[loc] = np.memmap(p, dtype=np.uint8, mode="r")
            out[t.name] = blobs32[loc][off:off + ln].view(np.float32).reshape(t.dims)
    return WeightTable(out)


@dataclass
class ModelDirTables:
    encoder: WeightTable       # fp32, 7 tensors
    t2s: WeightTable           # fp16, 291 tensors (shared by both decoder graphs)
    vits: WeightTable          # fp16
    prompt_encoder: Optional[WeightTable]
    is_v2pp: bool


def read_model_dir(model_dir: str) -> ModelDirTables:
    """V2 vs V2ProPlus is decided by the presence of prompt_encoder_fp32.onnx
    (reference: ModelManager.py:287-293)."""
    j = lambda n: os.path.join(model_dir, n)  # noqa: E731
    for f in (T2S_ENCODER, T2S_FIRST_STAGE, T2S_STAGE, VITS):
        if not os.path.exists(j(f)):
            raise FileNotFoundError(f"文件 {os.path.normpath(j(f))} 不存在！")
    pe = None
    if os.path.exists(j(PROMPT_ENCODER)):
        pe = read_table(j(PROMPT_ENCODER), j(PROMPT_ENCODER_BIN))
    return ModelDirTables(
        encoder=read_table(j(T2S_ENCODER), None),
        t2s=read_table(j(T2S_STAGE), j(T2S_BIN)),
        vits=read_table(j(VITS), j(VITS_BIN)),
        prompt_encoder=pe,
        is_v2pp=pe is not None,
    )
