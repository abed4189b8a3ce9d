"""Minimal ONNX ``ModelProto`` reader/writer on the raw protobuf wire format.

The converted model directory (``*.onnx`` + ``.bin``) stays the weight format
(reference: src/genie_tts/ModelManager.py:59-114 reads it with the ``onnx``
package, which is not a dependency here).  The product loader only needs the
initialiser table (name, dims, dtype, external-data offset/length); the test
oracle additionally walks nodes and attributes.

Field numbers follow onnx.proto3 (ModelProto.graph=7, GraphProto.node=1,
initializer=5, input=11, output=12; TensorProto dims=1, data_type=2, name=8,
raw_data=9, external_data=13, data_location=14 ...).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

# TensorProto.DataType -> numpy
ONNX_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64,
               9: np.bool_, 10: np.float16, 11: np.float64}
NP_TO_ONNX = {np.dtype(v): k for k, v in ONNX_DTYPES.items()}


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val, wt) -> List[int]:
    if wt == 0:
        return [_signed(val)]
    out = []
    pos = 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(_signed(v))
    return out


@dataclass
class Tensor:
    name: str = ""
    dims: Tuple[int, ...] = ()
    data_type: int = 0
    raw: Optional[bytes] = None          # inline payload, if any
    external: Dict[str, str] = field(default_factory=dict)
    data_location: int = 0

    @property
    def is_external(self) -> bool:
        return self.data_location == 1

    @property
    def numel(self) -> int:
        n = 1
        for d in self.dims:
            n *= d
        return n

    def numpy(self) -> np.ndarray:
        if self.raw is None:
            raise ValueError(f"tensor {self.name} has no inline data")
        return np.frombuffer(self.raw, dtype=ONNX_DTYPES[self.data_type]).reshape(self.dims).copy()


def _parse_tensor(buf: memoryview) -> Tensor:
    t = Tensor()
    dims: List[int] = []
    floats: List[bytes] = []
    i32: List[int] = []
    i64: List[int] = []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            dims.extend(_packed_varints(val, wt))
        elif fno == 2:
            t.data_type = val
        elif fno == 4:
            floats.append(bytes(val))
        elif fno == 5:
            i32.extend(_packed_varints(val, wt))
        elif fno == 7:
            i64.extend(_packed_varints(val, wt))
        elif fno == 8:
            t.name = bytes(val).decode()
        elif fno == 9:
            t.raw = bytes(val)
        elif fno == 13:
            k = v = ""
            for f2, _, v2 in _fields(val):
                if f2 == 1:
                    k = bytes(v2).decode()
                elif f2 == 2:
                    v = bytes(v2).decode()
            t.external[k] = v
        elif fno == 14:
            t.data_location = val
    t.dims = tuple(dims)
    if t.raw is None and not t.is_external:
        dt = ONNX_DTYPES.get(t.data_type)
        if floats:
            t.raw = b"".join(floats)
        elif i64:
            t.raw = np.asarray(i64, dtype=np.int64).astype(dt).tobytes()
        elif i32:
            if t.data_type == 10:   # fp16 bit patterns travel in int32_data
                t.raw = np.asarray(i32, dtype=np.uint16).tobytes()
            else:
                t.raw = np.asarray(i32, dtype=np.int32).astype(dt).tobytes()
        elif t.numel == 0:
            t.raw = b""
    return t


@dataclass
class Attribute:
    name: str = ""
    type: int = 0
    f: float = 0.0
    i: int = 0
    s: bytes = b""
    t: Optional[Tensor] = None
    g: Optional["Graph"] = None
    floats: List[float] = field(default_factory=list)
    ints: List[int] = field(default_factory=list)

    def value(self):
        return {1: self.f, 2: self.i, 3: self.s, 4: self.t, 5: self.g,
                6: self.floats, 7: self.ints}.get(self.type)


def _parse_attr(buf: memoryview) -> Attribute:
    a = Attribute()
    for fno, wt, val in _fields(buf):
        if fno == 1:
            a.name = bytes(val).decode()
        elif fno == 20:
            a.type = val
        elif fno == 2:
            a.f = struct.unpack("<f", val)[0]
        elif fno == 3:
            a.i = _signed(val)
        elif fno == 4:
            a.s = bytes(val)
        elif fno == 5:
            a.t = _parse_tensor(val)
        elif fno == 6:
            a.g = _parse_graph(val)
        elif fno == 7:
            if wt == 5:
                a.floats.append(struct.unpack("<f", val)[0])
            else:
                a.floats.extend(np.frombuffer(bytes(val), dtype="<f4").tolist())
        elif fno == 8:
            a.ints.extend(_packed_varints(val, wt))
    if a.type == 0:   # exporter omitted the type: infer
        if a.t is not None:
            a.type = 4
        elif a.g is not None:
            a.type = 5
        elif a.ints:
            a.type = 7
        elif a.floats:
            a.type = 6
        elif a.s:
            a.type = 3
    return a


@dataclass
class Node:
    op_type: str = ""
    name: str = ""
    inputs: List[str] = field(default_factory=list)
    outputs: List[str] = field(default_factory=list)
    attrs: Dict[str, Attribute] = field(default_factory=dict)

    def attr(self, name, default=None):
        a = self.attrs.get(name)
        return default if a is None else a.value()


def _parse_node(buf: memoryview) -> Node:
    n = Node()
    for fno, _, val in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(val).decode())
        elif fno == 2:
            n.outputs.append(bytes(val).decode())
        elif fno == 3:
            n.name = bytes(val).decode()
        elif fno == 4:
            n.op_type = bytes(val).decode()
        elif fno == 5:
            a = _parse_attr(val)
            n.attrs[a.name] = a
    return n


@dataclass
class ValueInfo:
    name: str = ""
    elem_type: int = 0
    shape: Tuple[object, ...] = ()


def _parse_value_info(buf: memoryview) -> ValueInfo:
    vi = ValueInfo()
    for fno, _, val in _fields(buf):
        if fno == 1:
            vi.name = bytes(val).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(val):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:
                        dims = []
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            d: object = None
                            for f5, _, v5 in _fields(v4):
                                if f5 == 1:
                                    d = _signed(v5)
                                elif f5 == 2:
                                    d = bytes(v5).decode()
                            dims.append(d)
                        vi.shape = tuple(dims)
    return vi


@dataclass
class Graph:
    name: str = ""
    nodes: List[Node] = field(default_factory=list)
    initializers: List[Tensor] = field(default_factory=list)
    inputs: List[ValueInfo] = field(default_factory=list)
    outputs: List[ValueInfo] = field(default_factory=list)


def _parse_graph(buf: memoryview, with_nodes: bool = True) -> Graph:
    g = Graph()
    for fno, _, val in _fields(buf):
        if fno == 1:
            if with_nodes:
                g.nodes.append(_parse_node(val))
        elif fno == 2:
            g.name = bytes(val).decode()
        elif fno == 5:
            g.initializers.append(_parse_tensor(val))
        elif fno == 11:
            g.inputs.append(_parse_value_info(val))
        elif fno == 12:
            g.outputs.append(_parse_value_info(val))
    return g


@dataclass
class Model:
    ir_version: int = 0
    producer_name: str = ""
    producer_version: str = ""
    opset: int = 0
    graph: Graph = field(default_factory=Graph)


def load_model(path: str, with_nodes: bool = True) -> Model:
    """Parse an ``.onnx`` file.  ``with_nodes=False`` skips node decoding (the
    product loader only needs the initialiser table)."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    m = Model()
    for fno, _, val in _fields(data):
        if fno == 1:
            m.ir_version = val
        elif fno == 2:
            m.producer_name = bytes(val).decode()
        elif fno == 3:
            m.producer_version = bytes(val).decode()
        elif fno == 7:
            m.graph = _parse_graph(val, with_nodes)
        elif fno == 8:
            for f2, _, v2 in _fields(val):
                if f2 == 2:
                    m.opset = max(m.opset, v2)
    return m


# --------------------------------------------------------------------------
# writer: just enough to emit a weights-only model (initialiser table with
# external_data), used by the fixture generator when no graph template exists.
# --------------------------------------------------------------------------

def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_field(fno: int, wt: int, payload) -> bytes:
    key = _enc_varint((fno << 3) | wt)
    if wt == 0:
        return key + _enc_varint(payload)
    return key + _enc_varint(len(payload)) + payload


def encode_external_tensor(name: str, dims, data_type: int, location: str,
                           offset: int, length: int) -> bytes:
    body = b"".join(_enc_field(1, 0, int(d)) for d in dims)
    body += _enc_field(2, 0, data_type)
    body += _enc_field(8, 2, name.encode())
    for k, v in (("location", location), ("offset", str(offset)), ("length", str(length))):
        entry = _enc_field(1, 2, k.encode()) + _enc_field(2, 2, v.encode())
        body += _enc_field(13, 2, entry)
    body += _enc_field(14, 0, 1)
    return body


def write_weights_only_model(path: str, tensors, graph_name: str = "weights_only") -> None:
    """``tensors``: iterable of (name, dims, onnx_dtype, location, offset, length)."""
    g = _enc_field(2, 2, graph_name.encode())
    for (name, dims, dt, loc, off, ln) in tensors:
        g += _enc_field(5, 2, encode_external_tensor(name, dims, dt, loc, off, ln))
    m = _enc_field(1, 0, 9) + _enc_field(2, 2, b"genie_b200-fixture")
    m += _enc_field(7, 2, g)
    m += _enc_field(8, 2, _enc_field(2, 0, 20))
    with open(path, "wb") as f:
        f.write(m)
