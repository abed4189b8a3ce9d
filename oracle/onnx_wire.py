"""TEST INFRASTRUCTURE — the oracle's OWN reader of .onnx files (protobuf wire format), independent of the
product's loader (genie_tts/onnx_reader.py): a wire-format or offset bug in one of the two cannot cancel out in
a parity test.  Schema-table driven generic decoder; field numbers are those of onnx.proto (ir_version 9, the
version of all six reference graphs) as listed in SURVEY.md Appendix A.

Weight materialisation follows the reference loader exactly:
  * fp16 side file: src/genie_tts/ModelManager.py:74-103 — the whole .bin is up-cast to fp32 and every initialiser
    takes blob32[offset/4 : (offset+length)/4] (offset / length are in fp32 byte units, `location` is ignored);
  * t2s_encoder: src/genie_tts/ModelManager.py:282-286 — onnxruntime resolves the fp32 .bin named by `location`.
"""
from __future__ import annotations

import os
import struct
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np

# message -> {field number: (attribute name, kind)}; kind: int | float | str | bytes | msg:<Message>, "rep " prefix
SCHEMA = {
    "Model": {7: ("graph", "msg:Graph")},
    "Graph": {1: ("nodes", "rep msg:Node"), 2: ("name", "str"), 5: ("initializers", "rep msg:Tensor"),
              11: ("inputs", "rep msg:ValueInfo"), 12: ("outputs", "rep msg:ValueInfo")},
    "Node": {1: ("inputs", "rep str"), 2: ("outputs", "rep str"), 3: ("name", "str"), 4: ("op_type", "str"),
             5: ("attributes", "rep msg:Attr")},
    "Attr": {1: ("name", "str"), 20: ("type", "int"), 2: ("f", "float"), 3: ("i", "int"), 4: ("s", "bytes"),
             5: ("t", "msg:Tensor"), 6: ("g", "msg:Graph"), 7: ("floats", "rep float"), 8: ("ints", "rep int"),
             9: ("strings", "rep bytes")},
    "Tensor": {1: ("dims", "rep int"), 2: ("data_type", "int"), 4: ("float_data", "rep float"),
               5: ("int32_data", "rep int"), 7: ("int64_data", "rep int"), 8: ("name", "str"),
               9: ("raw_data", "bytes"), 13: ("external_data", "rep msg:KV"), 14: ("data_location", "int")},
    "KV": {1: ("key", "str"), 2: ("value", "str")},
    "ValueInfo": {1: ("name", "str")},
}
NP_DTYPE = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16,
            11: np.float64}
ONNX_DTYPES = NP_DTYPE


def _varint(b: bytes, p: int):
    v = s = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7F) << s
        if c < 0x80:
            return v, p
        s += 7


def _s64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _decode(msg: str, b: bytes) -> SimpleNamespace:
    spec = SCHEMA[msg]
    out = SimpleNamespace(**{name: ([] if kind.startswith("rep ") else None) for name, kind in spec.values()})
    p, n = 0, len(b)
    while p < n:
        key, p = _varint(b, p)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            raw, p = _varint(b, p)
        elif wt == 1:
            raw, p = b[p:p + 8], p + 8
        elif wt == 5:
            raw, p = b[p:p + 4], p + 4
        elif wt == 2:
            ln, p = _varint(b, p)
            raw, p = b[p:p + ln], p + ln
        else:
            raise ValueError(f"wire type {wt} in {msg}")
        if fno not in spec:
            continue
        name, kind = spec[fno]
        rep = kind.startswith("rep ")
        base = kind[4:] if rep else kind
        if base == "int":
            if wt == 2:                                  # packed
                vals, q = [], 0
                while q < len(raw):
                    v, q = _varint(raw, q)
                    vals.append(_s64(v))
            else:
                vals = [_s64(raw)]
        elif base == "float":
            vals = list(struct.unpack(f"<{len(raw) // 4}f", raw))
        elif base == "str":
            vals = [bytes(raw).decode("utf-8")]
        elif base == "bytes":
            vals = [bytes(raw)]
        else:
            vals = [_decode(base[4:], bytes(raw))]
        if rep:
            getattr(out, name).extend(vals)
        else:
            setattr(out, name, vals[-1])
    return out


def _finish_tensor(t: SimpleNamespace) -> SimpleNamespace:
    t.dims = [int(d) for d in t.dims]
    t.external = {kv.key: kv.value for kv in t.external_data}
    t.is_external = t.data_location == 1 or bool(t.external)

    def numpy() -> np.ndarray:
        dt = NP_DTYPE[t.data_type]
        if t.raw_data is not None:
            a = np.frombuffer(t.raw_data, dtype=dt)
        elif t.float_data:
            a = np.asarray(t.float_data, dtype=dt)
        elif t.int64_data:
            a = np.asarray(t.int64_data, dtype=dt)
        elif t.int32_data:
            a = np.asarray(t.int32_data, dtype=np.int32).astype(dt)   # fp16 bit patterns are not used by these graphs
        else:
            a = np.zeros(0, dtype=dt)
        return np.array(a).reshape(t.dims)
    t.numpy = numpy
    return t


def _finish_graph(g: SimpleNamespace) -> SimpleNamespace:
    for t in g.initializers:
        _finish_tensor(t)
    for nd in g.nodes:
        table = {}
        for a in nd.attributes:
            ty = a.type
            if ty == 1:
                v = a.f
            elif ty == 2:
                v = a.i
            elif ty == 3:
                v = a.s
            elif ty == 4:
                v = _finish_tensor(a.t)
            elif ty == 5:
                v = _finish_graph(a.g)
            elif ty == 6:
                v = list(a.floats)
            elif ty == 7:
                v = list(a.ints)
            elif ty == 8:
                v = list(a.strings)
            else:
                raise ValueError(f"attribute type {ty}")
            table[a.name] = v
            a.value = (lambda val: (lambda: val))(v)
        nd.attrs = {a.name: a for a in nd.attributes}
        nd.attr = (lambda tb: (lambda name, default=None: tb.get(name, default)))(table)
    return g


# names the interpreter annotates with
Model = Graph = Node = SimpleNamespace


def load_model(path: str) -> SimpleNamespace:
    with open(path, "rb") as f:
        m = _decode("Model", f.read())
    _finish_graph(m.graph)
    return m


def read_tensors(onnx_path: str, fp16_bin: Optional[str] = None) -> Dict[str, np.ndarray]:
    """name -> fp32 array for every initialiser of a graph file, materialised as the reference loader does."""
    g = load_model(onnx_path).graph
    blob32 = None
    if fp16_bin is not None:
        blob32 = np.fromfile(fp16_bin, dtype=np.float16).astype(np.float32)       # ModelManager.py:75-77
    side: Dict[str, np.ndarray] = {}
    out: Dict[str, np.ndarray] = {}
    for t in g.initializers:
        if not t.is_external:
            out[t.name] = t.numpy()
            continue
        off, ln = int(t.external.get("offset", 0)), int(t.external.get("length", 0))
        if blob32 is not None:
            arr = blob32[off // 4:(off + ln) // 4]                                # ModelManager.py:80-103
        else:
            loc = t.external["location"]                                          # ModelManager.py:282-286
            if loc not in side:
                side[loc] = np.fromfile(os.path.join(os.path.dirname(onnx_path), loc), dtype=np.uint8)
            arr = side[loc][off:off + ln].view(NP_DTYPE[t.data_type])
        out[t.name] = np.array(arr, dtype=np.float32).reshape(t.dims)
    return out
