"""TEST INFRASTRUCTURE — CPU oracle, never on the product path.

Node-for-node fp32 executor (torch CPU) for the ONNX graphs that ARE the
reference's arithmetic: src/genie_tts/Data/{v2,v2ProPlus}/Models/*.onnx, which
the reference feeds to onnxruntime==1.22.1 (pyproject.toml:28; CPU EP selected
at src/genie_tts/ModelManager.py:125).  onnxruntime is a third-party dependency
that is absent from /root/reference and not installable offline, so this file
restates the published ONNX operator semantics (opset 20) for the 54 op types
the six graphs use and executes the reference's own
graph files.  Weights are materialised exactly as
src/genie_tts/ModelManager.py:74-103 does: fp16 ``.bin`` -> fp32, sliced by
each initialiser's external_data offset/length (fp32 byte units).

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md §4), and onnxruntime cannot run here, so *parity unpinned* against
the real runtime; what is pinned is (a) this interpreter executing the
reference's graph files, and (b) the independent restatement in
``oracle/gsv_port.py`` agreeing with it (tests/test_oracle_port.py) and with
the committed vectors in tests/golden/.

Two switches the reference lacks: ``RandomNormalLike`` is routed
through ``rng_hook`` so tests can make the samplers greedy (noise == 1) and
inject a known ``z_p`` noise into the vocoder.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .onnx_wire import ONNX_DTYPES, Graph, Model, Node, load_model   # the oracle's own wire reader

_TORCH_DT = {1: torch.float32, 6: torch.int32, 7: torch.int64, 9: torch.bool,
             10: torch.float16, 11: torch.float64, 2: torch.uint8, 3: torch.int8}


def _t(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def _ints(x) -> List[int]:
    return [int(v) for v in _t(x).reshape(-1).tolist()]


class OnnxProgram:
    """A parsed graph + materialised initialisers; ``run`` mirrors
    ``InferenceSession.run(None, feed)`` (reference call sites:
    src/genie_tts/Core/Inference.py:47,55,76,88,102)."""

    def __init__(self, onnx_path: str, fp16_bin: Optional[str] = None):
        self.path = onnx_path
        self.model: Model = load_model(onnx_path)
        g = self.model.graph
        self.input_names = [i.name for i in g.inputs]
        self.output_names = [o.name for o in g.outputs]
        self.weights: Dict[str, torch.Tensor] = {}
        blob16 = None
        if fp16_bin is not None:
            # ModelManager.py:75-77: whole blob fp16 -> fp32
            blob16 = np.fromfile(fp16_bin, dtype=np.float16)
        blob32_cache: Dict[str, np.ndarray] = {}
        base = os.path.dirname(onnx_path)
        for t in g.initializers:
            if t.is_external:
                off = int(t.external.get("offset", 0))
                ln = int(t.external.get("length", 0))
                if blob16 is not None:
                    # ModelManager.py:80-103: index the up-cast blob in fp32 byte units
                    arr = blob16[off // 4: (off + ln) // 4].astype(np.float32)
                else:
                    # ModelManager.py:282-286: ORT resolves the fp32 .bin next to the graph
                    loc = t.external["location"]
                    if loc not in blob32_cache:
                        blob32_cache[loc] = np.fromfile(os.path.join(base, loc), dtype=np.uint8)
                    arr = blob32_cache[loc][off: off + ln].view(ONNX_DTYPES[t.data_type])
                self.weights[t.name] = torch.from_numpy(np.array(arr).reshape(t.dims))
            else:
                self.weights[t.name] = torch.from_numpy(t.numpy())
        self.rng_hook: Optional[Callable[[str, torch.Tensor], torch.Tensor]] = None
        self.trace: Optional[Dict[str, torch.Tensor]] = None   # set to {} to keep all values
        self.keep: Optional[set] = None                        # names to keep when trace is set

    # -- public ------------------------------------------------------------
    def get_inputs(self):
        class _I:
            def __init__(self, n):
                self.name = n
        return [_I(n) for n in self.input_names]

    def run(self, output_names, feed: Dict[str, np.ndarray]) -> List[np.ndarray]:
        env: Dict[str, torch.Tensor] = dict(self.weights)
        for k, v in feed.items():
            env[k] = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else _t(v)
        with torch.no_grad():
            self._exec(self.model.graph, env)
        names = output_names or self.output_names
        return [env[n].numpy() if env[n].dtype != torch.bool or env[n].dim() else np.bool_(env[n].item())
                for n in names]

    # -- execution ---------------------------------------------------------
    def _exec(self, g: Graph, env: Dict[str, torch.Tensor]) -> None:
        for node in g.nodes:
            fn = getattr(self, "op_" + node.op_type, None)
            if fn is None:
                raise NotImplementedError(node.op_type)
            ins = [env[i] if i != "" else None for i in node.inputs]
            outs = fn(node, *ins) if node.op_type != "If" else self.op_If(node, ins[0], env)
            if not isinstance(outs, (tuple, list)):
                outs = (outs,)
            for name, val in zip(node.outputs, outs):
                if name:
                    env[name] = val
                    if self.trace is not None and (self.keep is None or name in self.keep):
                        self.trace[name] = val

    # -- elementwise -------------------------------------------------------
    def op_Add(self, n, a, b): return a + b
    def op_Sub(self, n, a, b): return a - b
    def op_Mul(self, n, a, b): return a * b

    def op_Div(self, n, a, b):
        if not a.is_floating_point() and not b.is_floating_point():
            return torch.div(a, b, rounding_mode="trunc")
        return a / b

    def op_Neg(self, n, a): return -a
    def op_Exp(self, n, a): return torch.exp(a)
    def op_Sqrt(self, n, a): return torch.sqrt(a)
    def op_Sin(self, n, a): return torch.sin(a)
    def op_Cos(self, n, a): return torch.cos(a)
    def op_Tanh(self, n, a): return torch.tanh(a)
    def op_Sigmoid(self, n, a): return torch.sigmoid(a)
    def op_Relu(self, n, a): return torch.relu(a)
    def op_Softplus(self, n, a): return torch.nn.functional.softplus(a)
    def op_Pow(self, n, a, b): return torch.pow(a, b.to(a.dtype) if b.is_floating_point() else b)
    def op_Max(self, n, *xs):
        r = xs[0]
        for x in xs[1:]:
            r = torch.maximum(r, x)
        return r

    def op_LeakyRelu(self, n, a):
        return torch.nn.functional.leaky_relu(a, n.attr("alpha", 0.01))

    def op_PRelu(self, n, a, slope):
        return torch.where(a < 0, a * slope, a)

    def op_Equal(self, n, a, b): return a == b
    def op_Less(self, n, a, b): return a < b
    def op_Greater(self, n, a, b): return a > b
    def op_Not(self, n, a): return ~a
    def op_Or(self, n, a, b): return a | b
    def op_Where(self, n, c, a, b): return torch.where(c, a, b)

    def op_Cast(self, n, a):
        return a.to(_TORCH_DT[n.attr("to")])

    # -- constants / shapes ------------------------------------------------
    def op_Constant(self, n):
        a = n.attrs.get("value")
        if a is not None:
            return torch.from_numpy(a.t.numpy())
        for key, conv in (("value_float", lambda v: torch.tensor(v, dtype=torch.float32)),
                          ("value_int", lambda v: torch.tensor(v, dtype=torch.int64)),
                          ("value_ints", lambda v: torch.tensor(v, dtype=torch.int64)),
                          ("value_floats", lambda v: torch.tensor(v, dtype=torch.float32))):
            if key in n.attrs:
                return conv(n.attrs[key].value())
        raise ValueError("Constant without value")

    def op_ConstantOfShape(self, n, shape):
        a = n.attrs.get("value")
        val = torch.from_numpy(a.t.numpy()) if a is not None else torch.zeros(1)
        return torch.full(_ints(shape), val.reshape(-1)[0].item(), dtype=val.dtype)

    def op_Shape(self, n, a):
        s = list(a.shape)
        st = n.attr("start", 0)
        en = n.attr("end", None)
        s = s[st:en] if en is not None else s[st:]
        return torch.tensor(s, dtype=torch.int64)

    def op_Reshape(self, n, a, shape):
        shp = _ints(shape)
        if not n.attr("allowzero", 0):
            shp = [a.shape[i] if d == 0 else d for i, d in enumerate(shp)]
        return a.reshape(shp)

    def op_Transpose(self, n, a):
        perm = n.attr("perm")
        if perm is None:
            perm = list(range(a.dim()))[::-1]
        return a.permute(perm)

    def op_Squeeze(self, n, a, axes=None):
        if axes is None:
            return a.squeeze()
        for ax in sorted([x % a.dim() for x in _ints(axes)], reverse=True):
            a = a.squeeze(ax)
        return a

    def op_Unsqueeze(self, n, a, axes):
        ax = _ints(axes)
        rank = a.dim() + len(ax)
        for x in sorted(v % rank for v in ax):
            a = a.unsqueeze(x)
        return a

    def op_Concat(self, n, *xs):
        return torch.cat([x for x in xs], dim=n.attr("axis"))

    def op_Split(self, n, a, split=None):
        axis = n.attr("axis", 0)
        if split is not None:
            return torch.split(a, _ints(split), dim=axis)
        k = n.attr("num_outputs", len(n.outputs))
        return torch.split(a, -(-a.shape[axis] // k), dim=axis)

    def op_Slice(self, n, a, starts, ends, axes=None, steps=None):
        starts, ends = _ints(starts), _ints(ends)
        axes = _ints(axes) if axes is not None else list(range(len(starts)))
        steps = _ints(steps) if steps is not None else [1] * len(starts)
        for s, e, ax, st in zip(starts, ends, axes, steps):
            ax %= a.dim()
            d = a.shape[ax]
            if st > 0:
                s = max(0, min(d, s + d if s < 0 else s))
                e = max(0, min(d, e + d if e < 0 else e))
                idx = torch.arange(s, e, st)
            else:
                s = max(-1, min(d - 1, s + d if s < 0 else s))
                e = max(-1, min(d - 1, e + d if e < 0 else e)) if e > -(1 << 62) else -1
                if e < -1:
                    e = -1
                idx = torch.arange(s, e, st)
            a = a.index_select(ax, idx)
        return a

    def op_Gather(self, n, a, idx):
        axis = n.attr("axis", 0) % a.dim()
        idx = idx.to(torch.int64)
        idx = torch.where(idx < 0, idx + a.shape[axis], idx)
        out = a.index_select(axis, idx.reshape(-1))
        return out.reshape(list(a.shape[:axis]) + list(idx.shape) + list(a.shape[axis + 1:]))

    def op_GatherElements(self, n, a, idx):
        axis = n.attr("axis", 0)
        idx = torch.where(idx < 0, idx + a.shape[axis], idx)
        return torch.gather(a, axis, idx)

    def op_ScatterElements(self, n, a, idx, upd):
        axis = n.attr("axis", 0)
        idx = torch.where(idx < 0, idx + a.shape[axis], idx)
        return a.clone().scatter_(axis, idx, upd)

    def op_Expand(self, n, a, shape):
        shp = _ints(shape)
        shp = list(torch.broadcast_shapes(tuple(a.shape), tuple(shp)))
        return a.expand(shp)

    def op_Tile(self, n, a, reps):
        return a.repeat(_ints(reps))

    def op_Pad(self, n, a, pads, value=None, axes=None):
        mode = n.attr("mode", b"constant")
        mode = mode.decode() if isinstance(mode, bytes) else mode
        p = _ints(pads)
        rank = a.dim()
        ax = [x % rank for x in _ints(axes)] if axes is not None else list(range(rank))
        k = len(ax)
        begin = [0] * rank
        end = [0] * rank
        for i, x in enumerate(ax):
            begin[x] = p[i]
            end[x] = p[i + k]
        # negative pads crop
        for d in range(rank):
            if begin[d] < 0:
                a = a.narrow(d, -begin[d], a.shape[d] + begin[d])
                begin[d] = 0
            if end[d] < 0:
                a = a.narrow(d, 0, a.shape[d] + end[d])
                end[d] = 0
        tp: List[int] = []
        for d in range(rank - 1, -1, -1):
            tp += [begin[d], end[d]]
        if mode == "constant":
            v = float(value.reshape(-1)[0]) if value is not None and value.numel() else 0.0
            return torch.nn.functional.pad(a, tp, mode="constant", value=v)
        # reflect: torch wants <= (rank-1) trailing padded dims on a batched tensor
        nz = [d for d in range(rank) if begin[d] or end[d]]
        if not nz:
            return a
        lo = min(nz)
        tp = tp[: 2 * (rank - lo)]
        lead = a.shape[:lo]
        x = a.reshape((1, -1) + tuple(a.shape[lo:])) if lo > 0 else a.reshape((1, 1) + tuple(a.shape))
        x = torch.nn.functional.pad(x, tp, mode=mode)
        return x.reshape(tuple(lead) + tuple(x.shape[2:]))

    def op_CumSum(self, n, a, axis):
        return torch.cumsum(a, dim=int(axis))

    # -- reductions / normalisation -----------------------------------------
    def op_ReduceSum(self, n, a, axes=None):
        keep = bool(n.attr("keepdims", 1))
        if axes is None or axes.numel() == 0:
            if n.attr("noop_with_empty_axes", 0):
                return a
            return a.sum() if not keep else a.sum().reshape([1] * a.dim())
        return a.sum(dim=_ints(axes), keepdim=keep)

    def op_ReduceL2(self, n, a, axes=None):
        keep = bool(n.attr("keepdims", 1))
        ax = _ints(axes) if axes is not None else n.attr("axes")
        return torch.sqrt((a * a).sum(dim=ax, keepdim=keep))

    def op_ArgMax(self, n, a):
        axis = n.attr("axis", 0)
        keep = bool(n.attr("keepdims", 1))
        # first occurrence of the maximum (select_last_index=0)
        m = a.max(dim=axis, keepdim=True).values
        is_max = (a == m) | (torch.isnan(a) if a.is_floating_point() else torch.zeros_like(a, dtype=torch.bool))
        idx = torch.arange(a.shape[axis]).reshape([-1 if d == axis % a.dim() else 1 for d in range(a.dim())])
        big = torch.where(is_max, idx, torch.full_like(idx, a.shape[axis]))
        r = big.min(dim=axis, keepdim=keep).values
        return r.to(torch.int64)

    def op_Softmax(self, n, a):
        return torch.softmax(a, dim=n.attr("axis", -1))

    def op_LayerNormalization(self, n, a, w, b=None):
        axis = n.attr("axis", -1) % a.dim()
        return torch.nn.functional.layer_norm(a, a.shape[axis:], w, b, n.attr("epsilon", 1e-5))

    def op_TopK(self, n, a, k):
        v, i = torch.topk(a, int(k.reshape(-1)[0]), dim=n.attr("axis", -1),
                          largest=bool(n.attr("largest", 1)), sorted=True)
        return v, i

    # -- linear algebra ------------------------------------------------------
    def op_MatMul(self, n, a, b): return torch.matmul(a, b)

    def op_Gemm(self, n, a, b, c=None):
        if n.attr("transA", 0):
            a = a.t()
        if n.attr("transB", 0):
            b = b.t()
        y = n.attr("alpha", 1.0) * (a @ b)
        if c is not None:
            y = y + n.attr("beta", 1.0) * c
        return y

    def op_Conv(self, n, x, w, b=None):
        pads = n.attr("pads", [0] * (2 * (x.dim() - 2)))
        nd = x.dim() - 2
        strides = n.attr("strides", [1] * nd)
        dil = n.attr("dilations", [1] * nd)
        grp = n.attr("group", 1)
        if pads[:nd] != pads[nd:]:
            tp: List[int] = []
            for d in range(nd - 1, -1, -1):
                tp += [pads[d], pads[d + nd]]
            x = torch.nn.functional.pad(x, tp)
            pads = [0] * (2 * nd)
        f = torch.nn.functional.conv1d if nd == 1 else torch.nn.functional.conv2d
        return f(x, w, b, stride=strides, padding=pads[:nd], dilation=dil, groups=grp)

    def op_ConvTranspose(self, n, x, w, b=None):
        nd = x.dim() - 2
        pads = n.attr("pads", [0] * (2 * nd))
        assert pads[:nd] == pads[nd:]
        f = torch.nn.functional.conv_transpose1d if nd == 1 else torch.nn.functional.conv_transpose2d
        return f(x, w, b, stride=n.attr("strides", [1] * nd), padding=pads[:nd],
                 output_padding=n.attr("output_padding", [0] * nd),
                 groups=n.attr("group", 1), dilation=n.attr("dilations", [1] * nd))

    def op_STFT(self, n, signal, frame_step, window=None, frame_length=None):
        # ONNX STFT: signal [B, L, 1] real; output [B, frames, bins, 2]
        onesided = bool(n.attr("onesided", 1))
        x = signal[..., 0] if signal.dim() == 3 else signal
        hop = int(frame_step)
        nfft = int(frame_length) if frame_length is not None else window.shape[0]
        win = window if window is not None else torch.ones(nfft)
        frames = x.unfold(-1, nfft, hop) * win
        spec = torch.fft.rfft(frames, n=nfft) if onesided else torch.fft.fft(frames, n=nfft)
        return torch.view_as_real(spec).to(torch.float32)

    def op_RandomNormalLike(self, n, a):
        if self.rng_hook is not None:
            return self.rng_hook(n.name, a)
        return torch.randn_like(a) * n.attr("scale", 1.0) + n.attr("mean", 0.0)

    def op_If(self, n, cond, env):
        g = n.attrs["then_branch" if bool(cond.reshape(-1)[0]) else "else_branch"].g
        sub = dict(env)
        self._exec(g, sub)
        outs = [sub[o.name] for o in g.outputs]
        return outs
