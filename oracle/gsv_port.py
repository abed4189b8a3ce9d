"""TEST INFRASTRUCTURE — CPU oracle ("port"), never on the product path.

Independent fp32 restatement (torch CPU) of the arithmetic in the reference's
six ONNX graphs, written as ordinary model code instead of a graph walk, plus
the reference host loop.  Every function cites the graph nodes it restates as
``<graph>#[node-index]`` (node index = position in graph.node)
or the reference file:line.  It is pinned against the reference's own graph
files executed by ``oracle/onnx_interp.py`` (tests/test_cpu.py, run in
the build container) and against the committed vectors in tests/golden/.
The real runtime (onnxruntime 1.22.1) is not installable here: parity with it
is unpinned.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .onnx_wire import read_tensors   # the oracle's own wire reader (independent of the product loader)

EOS = 1024
N_LAYER = 24
D_MODEL = 512
N_HEAD = 16


def _tt(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.array(a, dtype=np.float32))


class PortModel:
    """fp32 copies of all tensors of one model directory (ModelManager.py:74-103:
    fp16 blob -> fp32)."""

    def __init__(self, model_dir: str):
        j = lambda n: os.path.join(model_dir, n)  # noqa: E731
        # V2 vs V2ProPlus by the presence of the prompt encoder (src/genie_tts/ModelManager.py:287-293)
        self.is_v2pp = os.path.exists(j("prompt_encoder_fp32.onnx"))
        self.enc = {k: _tt(v) for k, v in read_tensors(j("t2s_encoder_fp32.onnx")).items()}
        self.t2s = {k: _tt(v) for k, v in
                    read_tensors(j("t2s_stage_decoder_fp32.onnx"), j("t2s_shared_fp16.bin")).items()}
        self.vits = {k[len("vq_model."):] if k.startswith("vq_model.") else k: _tt(v)
                     for k, v in read_tensors(j("vits_fp32.onnx"), j("vits_fp16.bin")).items()}
        self.pe = ({k: _tt(v) for k, v in
                    read_tensors(j("prompt_encoder_fp32.onnx"), j("prompt_encoder_fp16.bin")).items()}
                   if self.is_v2pp else None)
        self._wn_cache: Dict[str, torch.Tensor] = {}

    def wn(self, prefix: str) -> torch.Tensor:
        """weight-norm fold w = g * v / ||v||_2 over axes (1,2)
        (vits#[6508-6510] and 130 more; recomputed per call by the reference)."""
        if prefix not in self._wn_cache:
            v = self.vits[prefix + ".weight_v"]
            g = self.vits[prefix + ".weight_g"]
            self._wn_cache[prefix] = g * v / torch.sqrt((v * v).sum(dim=(1, 2), keepdim=True))
        return self._wn_cache[prefix]


# ---------------------------------------------------------------------------
# T2S
# ---------------------------------------------------------------------------

def sine_pe(n: int, start: int = 1) -> torch.Tensor:
    """Interleaved sin/cos table for positions start..start+n-1 (1-based in the
    graphs: CumSum of ones, t2s_encoder#[66-79], stage#[17-30]).  div_term is the
    256-vector constant at t2s_encoder#[70] == exp(arange(0,512,2) * -ln(1e4)/512)."""
    pos = torch.arange(start, start + n, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, D_MODEL, 2, dtype=torch.float32) * (-math.log(10000.0) / D_MODEL))
    ang = pos * div
    pe = torch.stack([torch.sin(ang), torch.cos(ang)], dim=-1).reshape(n, D_MODEL)
    return pe


def vq_prompts(m: PortModel, ssl_content: torch.Tensor) -> torch.Tensor:
    """t2s_encoder#[2-48]: Conv1d(768->768,k=2,s=2) then nearest codebook entry
    by argmax of -(|x|^2 - 2 x.E^T + |E|^2); returns int64 [T/2]."""
    h = F.conv1d(ssl_content, m.enc["vits.ssl_proj.weight"], m.enc["vits.ssl_proj.bias"], stride=2)
    x = h[0].t()                                             # [T/2, 768]
    e = m.enc["vits.quantizer.vq.layers.0._codebook.embed"]  # [1024, 768]
    dist = -((x * x).sum(1, keepdim=True) - (x * 2.0) @ e.t() + (e.t() * e.t()).sum(0, keepdim=True))
    return torch.argmax(dist, dim=-1)


def t2s_encode_text(m: PortModel, ref_seq, text_seq, ref_bert, text_bert) -> torch.Tensor:
    """t2s_encoder#[49-83]: x = Emb[ref||text] + bert @ W^T + b ; x*1.0 + alpha*PE(1..L)."""
    seq = torch.cat([ref_seq.reshape(-1), text_seq.reshape(-1)])
    bert = torch.cat([ref_bert, text_bert], dim=0)           # [L, 1024]
    x = m.enc["encoder.ar_text_embedding.word_embeddings.weight"][seq]
    x = x + (bert @ m.enc["encoder.bert_proj.weight"].t() + m.enc["encoder.bert_proj.bias"])
    return x + m.enc["encoder.ar_text_position.alpha"] * sine_pe(x.shape[0])


def _layer_weights(m: PortModel, i: int):
    p = f"transformer_encoder.layers.{i}."
    w = m.t2s
    return (w[p + "self_attn.in_proj_weight"], w[p + "self_attn.in_proj_bias"],
            w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"],
            w[p + "linear1.weight"], w[p + "linear1.bias"], w[p + "linear2.weight"], w[p + "linear2.bias"],
            w[p + "norm1.weight"], w[p + "norm1.bias"], w[p + "norm2.weight"], w[p + "norm2.bias"])


def _attn(q, k, v, mask=None):
    """16 heads x 32; sqrt(1/sqrt(32)) applied to both q and k^T
    (first_stage#[98-110], stage#[85-96])."""
    S, T = q.shape[0], k.shape[0]
    s = math.sqrt(1.0 / math.sqrt(32.0))
    qh = q.reshape(S, N_HEAD, 32).transpose(0, 1) * s
    kh = k.reshape(T, N_HEAD, 32).transpose(0, 1) * s
    vh = v.reshape(T, N_HEAD, 32).transpose(0, 1)
    sc = qh @ kh.transpose(1, 2)
    if mask is not None:
        sc = sc + mask
    p = torch.softmax(sc, dim=-1)
    return (p @ vh).transpose(0, 1).reshape(S, D_MODEL)


def _block(m: PortModel, i: int, h, k_all, v_all, mask):
    (wi, bi, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2) = _layer_weights(m, i)
    a = _attn(h @ wi[:512].t() + bi[:512], k_all, v_all, mask) @ wo.t() + bo
    h = F.layer_norm(h + a, (D_MODEL,), g1, be1, 1e-5)          # post-LN, first_stage#[118-119]
    f = torch.relu(h @ w1.t() + b1) @ w2.t() + b2               # #[120-126]
    return F.layer_norm(h + f, (D_MODEL,), g2, be2, 1e-5)       # #[127-128]


def t2s_prefill(m: PortModel, x: torch.Tensor, prompts: torch.Tensor):
    """first_stage#[5-1788]: returns (logits[1025], y_emb[Ly,512], k[24,S,512], v[24,S,512])."""
    Lx, Ly = x.shape[0], prompts.shape[0]
    y_emb = m.t2s["ar_audio_embedding.word_embeddings.weight"][prompts]
    y_pos = y_emb + m.t2s["ar_audio_position.alpha"] * sine_pe(Ly)
    h = torch.cat([x, y_pos], dim=0)
    S = Lx + Ly
    # additive mask (#[26-56]): text rows see text only; audio rows see text + causal audio
    mask = torch.zeros(S, S)
    mask[:Lx, Lx:] = float("-inf")
    causal = torch.triu(torch.ones(Ly, Ly, dtype=torch.bool), diagonal=1)
    mask[Lx:, Lx:] = torch.where(causal, float("-inf"), 0.0)
    ks, vs = [], []
    for i in range(N_LAYER):
        wi, bi = _layer_weights(m, i)[:2]
        k = h @ wi[512:1024].t() + bi[512:1024]
        v = h @ wi[1024:].t() + bi[1024:]
        ks.append(k)
        vs.append(v)
        h = _block(m, i, h, k, v, mask)
    logits = h[-1] @ m.t2s["ar_predict_layer.weight"].t()      # #[1785-1788], no bias
    return logits, y_emb, torch.stack(ks), torch.stack(vs)


def t2s_decode_step(m: PortModel, token: int, pos: int, k_cache, v_cache, T: int):
    """stage#[12-1774] for one new token at 1-based audio position ``pos``;
    appends K/V at row T of the caches (in place of the reference's Concat,
    stage#[63-64]) and returns logits[1025]."""
    e = m.t2s["ar_audio_embedding.word_embeddings.weight"][token]
    h = (e + m.t2s["ar_audio_position.alpha"] * sine_pe(1, start=pos)[0]).unsqueeze(0)
    for i in range(N_LAYER):
        wi, bi = _layer_weights(m, i)[:2]
        k_cache[i, T] = h[0] @ wi[512:1024].t() + bi[512:1024]
        v_cache[i, T] = h[0] @ wi[1024:].t() + bi[1024:]
        h = _block(m, i, h, k_cache[i, :T + 1], v_cache[i, :T + 1], None)   # no mask, #[94-96]
    return h[0] @ m.t2s["ar_predict_layer.weight"].t()


def sample_token(logits: torch.Tensor, history: torch.Tensor, noise: Optional[torch.Tensor] = None,
                 top_k: int = 15, temperature: float = 1.0, penalty: float = 1.35,
                 top_p: float = 1.0) -> Tuple[int, bool]:
    """stage#[1775-1821]: repetition penalty over every token in ``history``
    (gather -> where(s<0, s*p, s/p) -> scatter, so each distinct token is
    penalised once from the raw logit), /temperature, top-k keeping ties
    (logits < kth -> -inf), softmax, argmax(probs / noise); noise None == 1
    (greedy).  stop = argmax(raw logits)==EOS or token==EOS."""
    raw = logits
    s = raw[history]
    lg = raw.clone()
    lg[history] = torch.where(s < 0, s * penalty, s / penalty)
    if 0.0 < top_p < 1.0:
        # EXTENSION (not in the reference graphs; upstream GPT-SoVITS logits_to_probs order and rule):
        # descending sort, drop every token whose inclusive cumulative probability exceeds top_p, keep the first
        srt, order = torch.sort(lg, descending=True, stable=True)
        remove = torch.cumsum(torch.softmax(srt, dim=-1), dim=-1) > top_p
        remove[0] = False
        lg[order[remove]] = float("-inf")
    lg = lg / temperature
    kth = torch.topk(lg, top_k).values[-1]
    lg = torch.where(lg < kth, torch.tensor(float("-inf")), lg)
    probs = torch.softmax(lg, dim=-1)
    q = probs if noise is None else probs / noise
    tok = int(torch.argmax(q))
    stop = int(torch.argmax(raw)) == EOS or tok == EOS
    return tok, stop


@dataclass
class T2SResult:
    tokens: np.ndarray                 # what t2s_cpu returns: int64 [1,1,N]
    y_full: np.ndarray                 # prompts + all generated tokens before the [:, -idx:] slice
    idx: int
    logits: List[np.ndarray] = field(default_factory=list)


def _round_cache(kc, vc, rows: slice, mode: Optional[str]) -> None:
    """What-if switch for the KV-precision study: store the cache rows ``rows`` rounded through fp16
    ("kv"), only V ("v") or only K ("k").  None = the reference's fp32 cache."""
    if mode in ("kv", "k"):
        kc[:, rows] = kc[:, rows].half().float()
    if mode in ("kv", "v"):
        vc[:, rows] = vc[:, rows].half().float()


def t2s_generate(m: PortModel, ref_seq, ref_bert, text_seq, text_bert, ssl_content,
                 max_steps: int = 500, noise_fn=None, keep_logits: bool = False,
                 force_tokens: Optional[int] = None, kv_fp16: Optional[str] = None) -> T2SResult:
    """src/genie_tts/Core/Inference.py:63-109 over the port's stages, including
    the loop quirks: y[0,-1]=0 (:108) and y[:, -idx:] (:109; idx==0 returns all
    of y).  ``force_tokens``: ignore the stop flag and run exactly that many
    loop iterations (bench token budget)."""
    ref_seq = torch.as_tensor(ref_seq).reshape(-1)
    text_seq = torch.as_tensor(text_seq).reshape(-1)
    x = t2s_encode_text(m, ref_seq, text_seq, torch.as_tensor(ref_bert), torch.as_tensor(text_bert))
    prompts = vq_prompts(m, torch.as_tensor(ssl_content))
    logits, _, k0, v0 = t2s_prefill(m, x, prompts)
    S = k0.shape[1]
    Ly = prompts.shape[0]
    cap = S + max_steps + 1
    kc = torch.zeros(N_LAYER, cap, D_MODEL)
    vc = torch.zeros(N_LAYER, cap, D_MODEL)
    kc[:, :S] = k0
    vc[:, :S] = v0
    _round_cache(kc, vc, slice(0, S), kv_fp16)
    y = prompts.tolist()
    all_logits = [logits.numpy().copy()] if keep_logits else []
    tok, _ = sample_token(logits, torch.tensor(y), None if noise_fn is None else noise_fn(0))
    y.append(tok)                                            # first_stage#[1820]
    idx = 0
    T = S
    n_iter = max_steps if force_tokens is None else force_tokens
    for idx in range(0, n_iter):
        logits = t2s_decode_step(m, y[-1], Ly + idx + 1, kc, vc, T)
        _round_cache(kc, vc, slice(T, T + 1), kv_fp16)   # the step itself saw the new token's k / v in fp32
        T += 1
        if keep_logits:
            all_logits.append(logits.numpy().copy())
        tok, stop = sample_token(logits, torch.tensor(y), None if noise_fn is None else noise_fn(idx + 1))
        y.append(tok)
        if stop and force_tokens is None:
            break
    y_full = np.asarray(y, dtype=np.int64)[None]
    yy = y_full.copy()
    yy[0, -1] = 0
    return T2SResult(tokens=np.expand_dims(yy[:, -idx:], axis=0), y_full=y_full, idx=idx, logits=all_logits)


# ---------------------------------------------------------------------------
# SoVITS
# ---------------------------------------------------------------------------

def _mish(x):
    return x * torch.tanh(F.softplus(x))


def spectrogram(ref_audio: torch.Tensor) -> torch.Tensor:
    """vits#[3-79]: reflect-pad (2048-640)/2=704, STFT n_fft 2048 hop 640 with the
    periodic Hann constant, magnitude sqrt(re^2+im^2+1e-6), first 704 bins -> [1,704,F]."""
    x = F.pad(ref_audio.reshape(1, 1, -1), (704, 704), mode="reflect")[0, 0]
    win = torch.hann_window(2048, periodic=True)
    fr = x.unfold(0, 2048, 640) * win
    sp = torch.fft.rfft(fr, n=2048)
    mag = torch.sqrt(sp.real ** 2 + sp.imag ** 2 + 1e-6)      # [F, 1025]
    return mag[:, :704].t().unsqueeze(0)


def mel_style_encoder(w: Dict[str, torch.Tensor], prefix: str, spec: torch.Tensor) -> torch.Tensor:
    """vits#[109-271] / prompt_encoder#[..-268]: MelStyleEncoder -> [1, C_out]."""
    g = lambda n: w[prefix + n]  # noqa: E731
    x = spec.transpose(1, 2)                                             # [1,F,704]
    x = _mish(x @ g("spectral.0.fc.weight").t() + g("spectral.0.fc.bias"))
    x = _mish(x @ g("spectral.3.fc.weight").t() + g("spectral.3.fc.bias"))
    x = x.transpose(1, 2)                                                # [1,128,F]
    for i in (0, 1):                                                     # Conv1dGLU k=5 + residual
        y = F.conv1d(x, g(f"temporal.{i}.conv1.conv.weight"), g(f"temporal.{i}.conv1.conv.bias"), padding=2)
        a, b = y.split(128, dim=1)
        x = x + a * torch.sigmoid(b)
    x = x.transpose(1, 2)                                                # [1,F,128]
    res = x
    Fn = x.shape[1]
    q = (x @ g("slf_attn.w_qs.weight").t() + g("slf_attn.w_qs.bias")).reshape(1, Fn, 2, 64).permute(2, 0, 1, 3).reshape(2, Fn, 64)
    k = (x @ g("slf_attn.w_ks.weight").t() + g("slf_attn.w_ks.bias")).reshape(1, Fn, 2, 64).permute(2, 0, 1, 3).reshape(2, Fn, 64)
    v = (x @ g("slf_attn.w_vs.weight").t() + g("slf_attn.w_vs.bias")).reshape(1, Fn, 2, 64).permute(2, 0, 1, 3).reshape(2, Fn, 64)
    att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(128.0), dim=2)  # temperature sqrt(d_model), #[227]
    o = (att @ v).reshape(2, 1, Fn, 64).permute(1, 2, 0, 3).reshape(1, Fn, 128)
    x = o @ g("slf_attn.fc.weight").t() + g("slf_attn.fc.bias") + res
    x = x @ g("fc.fc.weight").t() + g("fc.fc.bias")                      # [1,F,C]
    return x.sum(dim=1) / Fn                                             # masked mean, no padding


def ref_enc_v2(m: PortModel, ref_audio_32k) -> torch.Tensor:
    """V2 global embedding ge [1,512,1] (vits#[3-271]); a pure function of the
    reference audio that the reference recomputes in every vocoder call."""
    return mel_style_encoder(m.vits, "ref_enc.", spectrogram(torch.as_tensor(ref_audio_32k))).unsqueeze(-1)


def prompt_encoder_v2pp(m: PortModel, ref_audio_32k, sv_emb) -> Tuple[torch.Tensor, torch.Tensor]:
    """prompt_encoder#[0-280] -> ge [1,1024,1], ge_advanced [1,512,1]
    (reference call site src/genie_tts/Audio/ReferenceAudio.py:68-76)."""
    w = m.pe
    ge = mel_style_encoder(w, "ref_enc.", spectrogram(torch.as_tensor(ref_audio_32k))).unsqueeze(-1)
    sv = torch.as_tensor(sv_emb) @ w["sv_emb.weight"].t() + w["sv_emb.bias"]
    ge = ge + sv.unsqueeze(-1)
    ge = torch.where(ge < 0, ge * w["prelu.weight"].reshape(1, -1, 1), ge)          # PRelu #[275]
    gea = (ge.transpose(1, 2) @ w["ge_to512.weight"].t() + w["ge_to512.bias"]).transpose(1, 2)
    return ge, gea


def _rel_attention(w, p: str, x: torch.Tensor, n_heads: int = 2, window: int = 4) -> torch.Tensor:
    """VITS MultiHeadAttention with windowed relative-position keys/values
    (vits#[314-765]); x [C,T] -> [C,T].  The graph's pad/reshape skew equals a
    band |i-j|<=window addressed as emb_rel[j-i+window]."""
    C, T = x.shape
    d = C // n_heads
    q = (w[p + "conv_q.weight"][:, :, 0] @ x + w[p + "conv_q.bias"][:, None]).reshape(n_heads, d, T).transpose(1, 2)
    k = (w[p + "conv_k.weight"][:, :, 0] @ x + w[p + "conv_k.bias"][:, None]).reshape(n_heads, d, T).transpose(1, 2)
    v = (w[p + "conv_v.weight"][:, :, 0] @ x + w[p + "conv_v.bias"][:, None]).reshape(n_heads, d, T).transpose(1, 2)
    q = q / math.sqrt(d)
    sc = q @ k.transpose(1, 2)                                            # [h,T,T]
    ek = w[p + "emb_rel_k"][0]                                            # [9,d]
    ev = w[p + "emb_rel_v"][0]
    rel = q @ ek.t()                                                      # [h,T,9]
    ii = torch.arange(T).unsqueeze(1)
    jj = torch.arange(T).unsqueeze(0)
    r = jj - ii + window
    band = (r >= 0) & (r <= 2 * window)
    rc = r.clamp(0, 2 * window)
    sc = sc + torch.where(band, torch.gather(rel, 2, rc.unsqueeze(0).expand(n_heads, T, T)), torch.tensor(0.0))
    pr = torch.softmax(sc, dim=-1)
    out = pr @ v                                                          # [h,T,d]
    # value side: sum_r p[i,i+r] * emb_rel_v[r+window]
    pw = torch.zeros(n_heads, T, 2 * window + 1)
    pw.scatter_add_(2, rc.unsqueeze(0).expand(n_heads, T, T), torch.where(band, pr, torch.tensor(0.0)))
    out = out + pw @ ev
    o = out.transpose(1, 2).reshape(C, T)
    return w[p + "conv_o.weight"][:, :, 0] @ o + w[p + "conv_o.bias"][:, None]


def _ln_c(x, g, b):
    """LayerNorm over channels of a [C,T] tensor (vits#[767-769])."""
    return F.layer_norm(x.t(), (x.shape[0],), g, b, 1e-5).t()


def _vits_encoder(w, p: str, x: torch.Tensor, n_layers: int) -> torch.Tensor:
    """attentions.Encoder, post-LN, FFN conv k=3 same-pad (vits#[313-1837])."""
    for i in range(n_layers):
        y = _rel_attention(w, f"{p}attn_layers.{i}.", x)
        x = _ln_c(x + y, w[f"{p}norm_layers_1.{i}.gamma"], w[f"{p}norm_layers_1.{i}.beta"])
        h = torch.relu(F.conv1d(x.unsqueeze(0), w[f"{p}ffn_layers.{i}.conv_1.weight"],
                                w[f"{p}ffn_layers.{i}.conv_1.bias"], padding=1))
        y = F.conv1d(h, w[f"{p}ffn_layers.{i}.conv_2.weight"], w[f"{p}ffn_layers.{i}.conv_2.bias"], padding=1)[0]
        x = _ln_c(x + y, w[f"{p}norm_layers_2.{i}.gamma"], w[f"{p}norm_layers_2.{i}.beta"])
    return x


def _mrte(w, ssl: torch.Tensor, text: torch.Tensor, ge512: torch.Tensor) -> torch.Tensor:
    """vits#[4891-4964]: cross attention 4 heads x 128, Q from ssl, K/V from text."""
    p = "enc_p.mrte."
    c1 = lambda n, x: w[p + n + ".weight"][:, :, 0] @ x + w[p + n + ".bias"][:, None]  # noqa: E731
    s = c1("c_pre", ssl)                                                 # [512,T]
    t = c1("text_pre", text)                                             # [512,L]
    T, L = s.shape[1], t.shape[1]
    q = c1("cross_attention.conv_q", s).reshape(4, 128, T).transpose(1, 2) / math.sqrt(128.0)
    k = c1("cross_attention.conv_k", t).reshape(4, 128, L).transpose(1, 2)
    v = c1("cross_attention.conv_v", t).reshape(4, 128, L).transpose(1, 2)
    o = (torch.softmax(q @ k.transpose(1, 2), dim=-1) @ v).transpose(1, 2).reshape(512, T)
    x = c1("cross_attention.conv_o", o) + s + ge512.reshape(512, 1)
    return c1("c_post", x)


def enc_p(m: PortModel, codes: torch.Tensor, text_seq: torch.Tensor, ge512: torch.Tensor):
    """vits#[273-6489]: codebook dequant, x2 nearest upsample, TextEncoder -> (m_p, logs_p) [192,2T]."""
    w = m.vits
    q = w["quantizer.vq.layers.0._codebook.embed"][codes].t()            # [768,T]  #[273-279]
    q = q.repeat_interleave(2, dim=1)                                    # #[280-292]
    y = w["enc_p.ssl_proj.weight"][:, :, 0] @ q + w["enc_p.ssl_proj.bias"][:, None]
    y = _vits_encoder(w, "enc_p.encoder_ssl.", y, 3)
    t = w["enc_p.text_embedding.weight"][text_seq].t()                   # [192,L]  #[1843-1845]
    t = _vits_encoder(w, "enc_p.encoder_text.", t, 6)
    y = _mrte(w, y, t, ge512)
    y = _vits_encoder(w, "enc_p.encoder2.", y, 3)
    st = w["enc_p.proj.weight"][:, :, 0] @ y + w["enc_p.proj.bias"][:, None]
    return st[:192], st[192:]


def flow_reverse(m: PortModel, z: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """vits#[6500-7820]: 4 x (flip, mean-only residual coupling reverse); z [192,T], g [gin,1]."""
    w = m.vits
    for fi in (6, 4, 2, 0):
        z = torch.flip(z, dims=[0])                                      # Slice step -1, #[6500]
        x0, x1 = z[:96], z[96:]
        p = f"flow.flows.{fi}."
        h = w[p + "pre.weight"][:, :, 0] @ x0 + w[p + "pre.bias"][:, None]
        cond = m.wn(p + "enc.cond_layer")[:, :, 0] @ g + w[p + "enc.cond_layer.bias"][:, None]   # [1536,1]
        out = torch.zeros_like(h)
        for li in range(4):
            xin = F.conv1d(h.unsqueeze(0), m.wn(f"{p}enc.in_layers.{li}"), w[f"{p}enc.in_layers.{li}.bias"],
                           padding=2)[0] + cond[li * 384:(li + 1) * 384]
            acts = torch.tanh(xin[:192]) * torch.sigmoid(xin[192:])
            rs = m.wn(f"{p}enc.res_skip_layers.{li}")[:, :, 0] @ acts + w[f"{p}enc.res_skip_layers.{li}.bias"][:, None]
            if li < 3:
                h = h + rs[:192]
                out = out + rs[192:]
            else:
                out = out + rs
        mean = w[p + "post.weight"][:, :, 0] @ out + w[p + "post.bias"][:, None]
        z = torch.cat([x0, x1 - mean], dim=0)                            # logs == 0, #[7815-7819]
    return z


def generator(m: PortModel, z: torch.Tensor, g: torch.Tensor, collect: Optional[dict] = None) -> torch.Tensor:
    """vits#[7822-8452] HiFi-GAN; z [192,T] -> audio [T*640]."""
    w = m.vits
    x = F.conv1d(z.unsqueeze(0), w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
    x = x + (w["dec.cond.weight"][:, :, 0] @ g + w["dec.cond.bias"][:, None]).unsqueeze(0)
    if collect is not None:
        collect["g_pre"] = x[0].t().contiguous()
    n_up = sum(1 for k in w if k.startswith("dec.ups.") and k.endswith("weight_g"))
    for i in range(n_up):
        x = F.leaky_relu(x, 0.1)
        wu = m.wn(f"dec.ups.{i}")                                        # [Cin,Cout,k]
        k = wu.shape[2]
        stride = (10, 8, 2, 2, 2)[i]
        x = F.conv_transpose1d(x, wu, w[f"dec.ups.{i}.bias"], stride=stride, padding=(k - stride) // 2)
        xs = None
        for j in range(3):
            r = x
            kk = (3, 7, 11)[j]
            for c, dil in enumerate((1, 3, 5)):
                t = F.leaky_relu(r, 0.1)
                t = F.conv1d(t, m.wn(f"dec.resblocks.{i * 3 + j}.convs1.{c}"),
                             w[f"dec.resblocks.{i * 3 + j}.convs1.{c}.bias"], padding=dil * (kk - 1) // 2, dilation=dil)
                t = F.leaky_relu(t, 0.1)
                t = F.conv1d(t, m.wn(f"dec.resblocks.{i * 3 + j}.convs2.{c}"),
                             w[f"dec.resblocks.{i * 3 + j}.convs2.{c}.bias"], padding=(kk - 1) // 2)
                r = t + r
            xs = r if xs is None else xs + r
        x = xs / 3.0
        if collect is not None:
            collect[f"g_s{i}"] = xs[0].t().contiguous()
    x = F.leaky_relu(x, 0.01)                                            # #[8450]
    x = F.conv1d(x, w["dec.conv_post.weight"], None, padding=3)          # no bias, #[8451]
    return torch.tanh(x)[0, 0]


def vits_decode(m: PortModel, text_seq, pred_semantic, ge: torch.Tensor, ge_advanced: Optional[torch.Tensor] = None,
                zp_noise: Optional[torch.Tensor] = None, noise_scale: float = 0.5,
                collect: Optional[dict] = None) -> np.ndarray:
    """vits#[273-8452] given the global embedding(s).  V2: ge [1,512,1] everywhere.
    V2ProPlus: MRTE gets ge_advanced [1,512,1]; flow and generator get ge [1,1024,1]."""
    codes = torch.as_tensor(pred_semantic).reshape(-1)
    text = torch.as_tensor(text_seq).reshape(-1)
    ge = torch.as_tensor(ge)
    g_m = (torch.as_tensor(ge_advanced) if ge_advanced is not None else ge).reshape(-1, 1)
    g = ge.reshape(-1, 1)
    m_p, logs_p = enc_p(m, codes, text, g_m)
    noise = torch.zeros_like(m_p) if zp_noise is None else torch.as_tensor(zp_noise).reshape(192, -1)[:, :m_p.shape[1]]
    z_p = m_p + noise * torch.exp(logs_p) * noise_scale                   # #[6490-6495]
    z = flow_reverse(m, z_p, g)
    if collect is not None:
        collect["z"] = z.t().contiguous()
    return generator(m, z, g, collect).numpy()
