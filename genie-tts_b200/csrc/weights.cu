// Model store: raw tensors (as stored in the model directory) -> kernel-ready
// device layouts.  Replaces load_session_with_fp16_conversion
// (reference src/genie_tts/ModelManager.py:59-114): the fp16 payload is copied to
// the device as is; T2S matrices stay fp16 (exact), vectors become fp32, VITS conv
// weights are weight-norm-folded once (the reference redoes vits#[6508-6510] x131
// in every call) and repacked [Cout][k][Cin] for channels-last implicit GEMM.
#include "model.h"
#include <cstdlib>

namespace genie {

std::atomic<unsigned long long> g_launches{0};
thread_local unsigned long long* t_capture_counter = nullptr;
int g_sync_debug = []() { const char* e = getenv("GENIE_SYNC_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
thread_local int g_pdl_now = 1;
int g_pdl = []() { const char* e = getenv("GENIE_PDL"); return (e && e[0] == '0') ? 0 : 1; }();

Model::~Model() {
  cudaSetDevice(device);
  t2s_session.reset();                 // events of the slot pool and the captured steps before the streams go
  t2s_graphs.reset();
  for (void* p : ctx_owned) cudaFree(p);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (stream_bulk) cudaStreamDestroy(stream_bulk);
  if (ev_bulk) cudaEventDestroy(ev_bulk);
  if (stream2) cudaStreamDestroy(stream2);
  if (stream3) cudaStreamDestroy(stream3);
  if (stream4) cudaStreamDestroy(stream4);
  if (ev_join3) cudaEventDestroy(ev_join3);
  if (ev_join4) cudaEventDestroy(ev_join4);
  if (stream && stream_owned) cudaStreamDestroy(stream);
  // the weights (owner) are released with the last handle that shares them
}

namespace {

__global__ void to_f32_kernel(const void* src, int f16, float* dst, long long n, float scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = f16 ? __half2float(reinterpret_cast<const __half*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
  dst[i] = v * scale;
}

// one block per d0: optional weight-norm over (d1,d2), scale, and repack.
// layout 0: out[d0][d2][d1]   (Conv1d   [Cout,Cin,k] -> [Cout][k][Cin])
// layout 1: out[d2][d1][d0]   (ConvT1d  [Cin,Cout,k] -> [k][Cout][Cin])
__global__ void prep_weight_kernel(const void* v, int v_f16, const void* g, int g_f16, int D0, int D1, int D2,
                                   float scale, float* out, int layout) {
  const int d0 = blockIdx.x;
  const long long n = (long long)D1 * D2;
  auto ld = [&](long long i) {
    return v_f16 ? __half2float(reinterpret_cast<const __half*>(v)[d0 * n + i])
                 : reinterpret_cast<const float*>(v)[d0 * n + i];
  };
  __shared__ float red[32];
  float gain = 1.f, inv = 1.f;
  if (g) {
    float ss = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { float t = ld(i); ss = fmaf(t, t, ss); }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    inv = 1.f / sqrtf(tot);
    gain = g_f16 ? __half2float(reinterpret_cast<const __half*>(g)[d0]) : reinterpret_cast<const float*>(g)[d0];
  }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    int d1 = (int)(i / D2), d2 = (int)(i % D2);
    float val = g ? (ld(i) * inv) * gain : ld(i);      // (v / ||v||) * g, vits#[6509-6510]
    val *= scale;
    long long o = layout == 0 ? ((long long)d0 * D2 + d2) * D1 + d1 : ((long long)d2 * D1 + d1) * D0 + d0;
    out[o] = val;
  }
}

// fp32 weights addressed like ConvGemm (co*co_stride + tap*tap_stride + ci) -> fp16 hi (+lo) [Cout][kpad]
__global__ void pack_tc_kernel(const float* __restrict__ w, long long co_stride, long long tap_stride, int ntaps,
                               int Cin, int Cout, int kpad, __half* __restrict__ hi, __half* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * kpad) return;
  const int co = (int)(i / kpad), kk = (int)(i % kpad);
  float v = 0.f;
  if (kk < ntaps * Cin) {
    const int tap = kk / Cin, ci = kk - tap * Cin;
    v = w[co * co_stride + tap * tap_stride + ci];
  }
  const __half h = __float2half_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2half_rn(v - __half2float(h));
}

struct Finalizer {
  Model& m;
  cudaStream_t s;
  explicit Finalizer(Model& mm) : m(mm), s(mm.stream) {}

  const RawTensor& raw(int g, const std::string& name) {
    auto it = m.raw[g].find(name);
    GENIE_CHECK(it != m.raw[g].end(), "missing tensor '" + name + "' in graph " + std::to_string(g));
    return it->second;
  }
  bool has(int g, const std::string& name) { return m.raw[g].count(name) != 0; }

  float* f32(int g, const std::string& name, float scale = 1.f) {
    const RawTensor& t = raw(g, name);
    if (!t.f16 && scale == 1.f) return reinterpret_cast<float*>(t.d);
    float* d = dev_alloc<float>(m.owner->owned, t.numel);
    m.weight_bytes += t.numel * 4;
    to_f32_kernel<<<(unsigned)((t.numel + 255) / 256), 256, 0, s>>>(t.d, t.f16, d, t.numel, scale);
    GENIE_LAUNCHED("to_f32");
    return d;
  }
  Linear linear(int g, const std::string& w, const std::string& b, bool keep_f16) {
    const RawTensor& t = raw(g, w);
    GENIE_CHECK(t.dims.size() == 2, "linear weight must be 2-D: " + w);
    Linear L;
    L.N = (int)t.dims[0]; L.K = (int)t.dims[1];
    if (t.f16 && keep_f16) { L.w = t.d; L.w_f16 = 1; } else { L.w = f32(g, w); L.w_f16 = 0; }
    L.b = b.empty() ? nullptr : f32(g, b);
    return L;
  }
  // Conv1d weight [Cout,Cin,k] (optionally weight-normed: <prefix>.weight_v/.weight_g)
  Conv conv(int g, const std::string& prefix, bool bias = true, float scale = 1.f) {
    const bool wn = has(g, prefix + ".weight_v");
    const RawTensor& v = raw(g, prefix + (wn ? ".weight_v" : ".weight"));
    GENIE_CHECK(v.dims.size() == 3, "conv weight must be 3-D: " + prefix);
    Conv c;
    c.Cout = (int)v.dims[0]; c.Cin = (int)v.dims[1]; c.k = (int)v.dims[2];
    float* out = dev_alloc<float>(m.owner->owned, v.numel);
    m.weight_bytes += v.numel * 4;
    const RawTensor* gt = wn ? &raw(g, prefix + ".weight_g") : nullptr;
    prep_weight_kernel<<<c.Cout, 256, 0, s>>>(v.d, v.f16, gt ? gt->d : nullptr, gt ? gt->f16 : 0, c.Cout, c.Cin, c.k,
                                              scale, out, 0);
    GENIE_LAUNCHED("prep_weight");
    c.w = out;
    c.b = bias ? f32(g, prefix + ".bias") : nullptr;
    return c;
  }
  ConvT convt(int g, const std::string& prefix, int stride, float scale) {
    const RawTensor& v = raw(g, prefix + ".weight_v");
    const RawTensor& gt = raw(g, prefix + ".weight_g");
    ConvT c;
    c.Cin = (int)v.dims[0]; c.Cout = (int)v.dims[1]; c.k = (int)v.dims[2];
    c.stride = stride; c.pad = (c.k - stride) / 2;
    float* out = dev_alloc<float>(m.owner->owned, v.numel);
    m.weight_bytes += v.numel * 4;
    prep_weight_kernel<<<c.Cin, 256, 0, s>>>(v.d, v.f16, gt.d, gt.f16, c.Cin, c.Cout, c.k, scale, out, 1);
    GENIE_LAUNCHED("prep_weight");
    c.w = out;
    c.b = f32(g, prefix + ".bias");
    return c;
  }
  void enc_layers(VitsEncLayer* L, int n, const std::string& p) {
    const int G = GENIE_GRAPH_VITS;
    for (int i = 0; i < n; ++i) {
      std::string a = p + "attn_layers." + std::to_string(i) + ".";
      L[i].q = conv(G, a + "conv_q"); L[i].k = conv(G, a + "conv_k");
      L[i].v = conv(G, a + "conv_v"); L[i].o = conv(G, a + "conv_o");
      L[i].rel_k = f32(G, a + "emb_rel_k"); L[i].rel_v = f32(G, a + "emb_rel_v");
      std::string f = p + "ffn_layers." + std::to_string(i) + ".";
      L[i].ff1 = conv(G, f + "conv_1"); L[i].ff2 = conv(G, f + "conv_2");
      L[i].g1 = f32(G, p + "norm_layers_1." + std::to_string(i) + ".gamma");
      L[i].b1 = f32(G, p + "norm_layers_1." + std::to_string(i) + ".beta");
      L[i].g2 = f32(G, p + "norm_layers_2." + std::to_string(i) + ".gamma");
      L[i].b2 = f32(G, p + "norm_layers_2." + std::to_string(i) + ".beta");
    }
  }
  MelStyle mel_style(int g, const std::string& p) {
    MelStyle e;
    e.fc1 = linear(g, p + "spectral.0.fc.weight", p + "spectral.0.fc.bias", false);
    e.fc2 = linear(g, p + "spectral.3.fc.weight", p + "spectral.3.fc.bias", false);
    e.t0 = conv(g, p + "temporal.0.conv1.conv"); e.t1 = conv(g, p + "temporal.1.conv1.conv");
    e.wq = linear(g, p + "slf_attn.w_qs.weight", p + "slf_attn.w_qs.bias", false);
    e.wk = linear(g, p + "slf_attn.w_ks.weight", p + "slf_attn.w_ks.bias", false);
    e.wv = linear(g, p + "slf_attn.w_vs.weight", p + "slf_attn.w_vs.bias", false);
    e.fo = linear(g, p + "slf_attn.fc.weight", p + "slf_attn.fc.bias", false);
    e.fc = linear(g, p + "fc.fc.weight", p + "fc.fc.bias", false);
    e.out_dim = e.fc.N;
    return e;
  }

  TcW pack(const float* w, long long co_stride, long long tap_stride, int ntaps, int Cin, int Cout) {
    TcW t;
    t.kpad = ((ntaps * Cin + 63) / 64) * 64;
    const long long n = (long long)Cout * t.kpad;
    __half* hi = dev_alloc<__half>(m.owner->owned, n);
    __half* lo = dev_alloc<__half>(m.owner->owned, n);
    m.weight_bytes += n * 4;
    pack_tc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, co_stride, tap_stride, ntaps, Cin, Cout, t.kpad, hi, lo);
    GENIE_LAUNCHED("pack_tc");
    t.hi = hi; t.lo = lo;
    if (tap_stride == Cin && co_stride == (long long)ntaps * Cin && pretile_w128_supported(Cin, Cout, ntaps)) {
      const long long nt = pretile_w128_halves(Cin, Cout, ntaps);     // Cout rounded up to whole 128-row tiles
      __half* tiles = dev_alloc<__half>(m.owner->owned, nt);
      m.weight_bytes += nt * 2;
      launch_pretile_w128(hi, Cout, t.kpad, Cin, ntaps, tiles, s);
      t.tiles = tiles;
    }
    return t;
  }
  void pack_conv(Conv& c) { c.tc = pack(c.w, (long long)c.k * c.Cin, c.Cin, c.k, c.Cin, c.Cout); }
  void pack_convt(ConvT& c) {
    const long long tap_sz = (long long)c.Cout * c.Cin;
    for (int r = 0; r < c.stride && r < 10; ++r) {
      const int ntaps = (c.k - r + c.stride - 1) / c.stride;
      if (ntaps > 0) c.tc[r] = pack(c.w + r * tap_sz, c.Cin, c.stride * tap_sz, ntaps, c.Cin, c.Cout);
    }
    if (c.k == c.stride && c.pad == 0 && c.Cin % 4 == 0) {
      c.tc_fused = pack(c.w, c.Cin, 0, 1, c.Cin, c.Cout * c.k);
      float* bf = dev_alloc<float>(m.owner->owned, (size_t)c.Cout * c.k);
      for (int r = 0; r < c.k; ++r) {
        if (c.b) GENIE_CUDA(cudaMemcpyAsync(bf + (size_t)r * c.Cout, c.b, (size_t)c.Cout * 4, cudaMemcpyDeviceToDevice, s));
        else GENIE_CUDA(cudaMemsetAsync(bf + (size_t)r * c.Cout, 0, (size_t)c.Cout * 4, s));
      }
      c.bias_fused = bf;
    }
  }
  void pack_linear(Linear& L) {
    if (L.w_f16 && L.K % 64 == 0) { L.tc.hi = reinterpret_cast<const __half*>(L.w); L.tc.lo = nullptr; L.tc.kpad = L.K; }
  }
  void pack_enc(VitsEncLayer* L, int n) {
    for (int i = 0; i < n; ++i) {
      pack_conv(L[i].q); pack_conv(L[i].k); pack_conv(L[i].v); pack_conv(L[i].o);
      pack_conv(L[i].ff1); pack_conv(L[i].ff2);
    }
  }
  void pack_all() {
    for (int i = 0; i < 24; ++i) {
      pack_linear(m.layers[i].qkv); pack_linear(m.layers[i].out);
      pack_linear(m.layers[i].ff1); pack_linear(m.layers[i].ff2);
    }
    pack_linear(m.predict);
    pack_conv(m.ssl_proj); pack_conv(m.enc_proj);
    pack_enc(m.enc_ssl, 3); pack_enc(m.enc_text, 6); pack_enc(m.enc2, 3);
    pack_conv(m.mrte_c_pre); pack_conv(m.mrte_text_pre); pack_conv(m.mrte_q); pack_conv(m.mrte_k);
    pack_conv(m.mrte_v); pack_conv(m.mrte_o); pack_conv(m.mrte_c_post);
    for (int f = 0; f < 4; ++f) {
      pack_conv(m.flow[f].pre); pack_conv(m.flow[f].post);
      for (int l = 0; l < 4; ++l) { pack_conv(m.flow[f].wn[l].in); pack_conv(m.flow[f].wn[l].rs); }
    }
    pack_conv(m.dec_pre);
    for (int i = 0; i < m.n_up; ++i) {
      pack_convt(m.ups[i]);
      for (int j = 0; j < 3; ++j)
        for (int c = 0; c < 3; ++c) { pack_conv(m.res[i * 3 + j].c1[c]); pack_conv(m.res[i * 3 + j].c2[c]); }
    }
    m.tc_err = dev_alloc<int>(m.ctx_owned, 1);
    GENIE_CUDA(cudaMemsetAsync(m.tc_err, 0, sizeof(int), s));
  }

  void run() {
    const int E = GENIE_GRAPH_T2S_ENCODER, T = GENIE_GRAPH_T2S, V = GENIE_GRAPH_VITS, P = GENIE_GRAPH_PROMPT_ENCODER;
    m.v2pp = !m.raw[P].empty();
    GENIE_CHECK(m.div_term != nullptr, "genie_model_set_constants must be called before finalize");
    // ---- T2S
    m.text_emb = f32(E, "encoder.ar_text_embedding.word_embeddings.weight");
    m.text_vocab = (int)raw(E, "encoder.ar_text_embedding.word_embeddings.weight").dims[0];
    m.text_alpha = f32(E, "encoder.ar_text_position.alpha");
    m.bert_proj = linear(E, "encoder.bert_proj.weight", "encoder.bert_proj.bias", false);
    m.audio_emb = f32(T, "ar_audio_embedding.word_embeddings.weight");
    m.audio_alpha = f32(T, "ar_audio_position.alpha");
    m.predict = linear(T, "ar_predict_layer.weight", "", true);
    for (int i = 0; i < 24; ++i) {
      std::string p = "transformer_encoder.layers." + std::to_string(i) + ".";
      T2SLayer& L = m.layers[i];
      L.qkv = linear(T, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", true);
      L.out = linear(T, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", true);
      L.ff1 = linear(T, p + "linear1.weight", p + "linear1.bias", true);
      L.ff2 = linear(T, p + "linear2.weight", p + "linear2.bias", true);
      L.ln1_g = f32(T, p + "norm1.weight"); L.ln1_b = f32(T, p + "norm1.bias");
      L.ln2_g = f32(T, p + "norm2.weight"); L.ln2_b = f32(T, p + "norm2.bias");
    }
    // pointer table of the persistent decode step (t2s_persistent.cu): immutable, shared by every context
    if (m.layers[0].qkv.w_f16 && m.predict.w_f16) {
      std::vector<StepLayerPtrs> hl(24);
      for (int l = 0; l < 24; ++l) {
        const T2SLayer& L = m.layers[l];
        hl[l] = StepLayerPtrs{reinterpret_cast<const __half*>(L.qkv.w), reinterpret_cast<const __half*>(L.out.w),
                              reinterpret_cast<const __half*>(L.ff1.w), reinterpret_cast<const __half*>(L.ff2.w),
                              L.qkv.b, L.out.b, L.ff1.b, L.ff2.b, L.ln1_g, L.ln1_b, L.ln2_g, L.ln2_b};
      }
      StepLayerPtrs* d = dev_alloc<StepLayerPtrs>(m.owner->owned, 24);
      GENIE_CUDA(cudaMemcpy(d, hl.data(), sizeof(StepLayerPtrs) * 24, cudaMemcpyHostToDevice));
      m.step_layers_dev = d;
    }
    GENIE_CUDA(cudaDeviceGetAttribute(&m.num_sms, cudaDevAttrMultiProcessorCount, m.device));
    // prompt-time VQ (t2s_encoder#[2-48])
    m.ssl_vq = conv(E, "vits.ssl_proj");
    m.codebook_enc = f32(E, "vits.quantizer.vq.layers.0._codebook.embed");
    {
      float* sq = dev_alloc<float>(m.owner->owned, 1024);
      launch_row_sqnorm(m.codebook_enc, 768, 768, 1024, sq, s);
      m.codebook_enc_sq = sq;
    }
    // ---- VITS
    const std::string q = "vq_model.";
    m.codebook = f32(V, q + "quantizer.vq.layers.0._codebook.embed");
    m.vits_text_emb = f32(V, q + "enc_p.text_embedding.weight");
    m.vits_text_vocab = (int)raw(V, q + "enc_p.text_embedding.weight").dims[0];
    m.ssl_proj = conv(V, q + "enc_p.ssl_proj");
    m.enc_proj = conv(V, q + "enc_p.proj");
    enc_layers(m.enc_ssl, 3, q + "enc_p.encoder_ssl.");
    enc_layers(m.enc_text, 6, q + "enc_p.encoder_text.");
    enc_layers(m.enc2, 3, q + "enc_p.encoder2.");
    m.mrte_c_pre = conv(V, q + "enc_p.mrte.c_pre"); m.mrte_text_pre = conv(V, q + "enc_p.mrte.text_pre");
    m.mrte_q = conv(V, q + "enc_p.mrte.cross_attention.conv_q"); m.mrte_k = conv(V, q + "enc_p.mrte.cross_attention.conv_k");
    m.mrte_v = conv(V, q + "enc_p.mrte.cross_attention.conv_v"); m.mrte_o = conv(V, q + "enc_p.mrte.cross_attention.conv_o");
    m.mrte_c_post = conv(V, q + "enc_p.mrte.c_post");
    const int order[4] = {6, 4, 2, 0};     // flow executed in reverse (vits#[6500-7820])
    for (int f = 0; f < 4; ++f) {
      std::string p = q + "flow.flows." + std::to_string(order[f]) + ".";
      FlowStep& F = m.flow[f];
      F.pre = conv(V, p + "pre"); F.post = conv(V, p + "post"); F.cond = conv(V, p + "enc.cond_layer");
      for (int l = 0; l < 4; ++l) {
        F.wn[l].in = conv(V, p + "enc.in_layers." + std::to_string(l));
        F.wn[l].rs = conv(V, p + "enc.res_skip_layers." + std::to_string(l));
      }
    }
    m.gin = m.flow[0].cond.Cin;
    m.dec_pre = conv(V, q + "dec.conv_pre"); m.dec_cond = conv(V, q + "dec.cond");
    const int strides[5] = {10, 8, 2, 2, 2};
    m.n_up = 5;
    for (int i = 0; i < 5; ++i) {
      // the mean over the 3 resblocks of the previous stage (/3, vits#[7949]) is folded into this
      // stage's transposed-conv weights: leaky-relu is positively homogeneous.
      m.ups[i] = convt(V, q + "dec.ups." + std::to_string(i), strides[i], i == 0 ? 1.f : (1.f / 3.f));
      for (int j = 0; j < 3; ++j) {
        ResBlock& R = m.res[i * 3 + j];
        std::string p = q + "dec.resblocks." + std::to_string(i * 3 + j) + ".";
        for (int c = 0; c < 3; ++c) {
          R.c1[c] = conv(V, p + "convs1." + std::to_string(c));
          R.c2[c] = conv(V, p + "convs2." + std::to_string(c));
        }
        R.k = R.c1[0].k;
      }
    }
    {
      Conv post = conv(V, q + "dec.conv_post", /*bias=*/false, 1.f / 3.f);   // [1][7][C]
      m.conv_post = post.w; m.c_last = post.Cin;
    }
    if (m.v2pp) {
      m.ref_enc = mel_style(P, "ref_enc.");
      m.sv_emb = linear(P, "sv_emb.weight", "sv_emb.bias", true);
      m.ge_to512 = linear(P, "ge_to512.weight", "ge_to512.bias", false);
      m.prelu = f32(P, "prelu.weight");
    } else {
      m.ref_enc = mel_style(V, q + "ref_enc.");
    }
    pack_all();
    m.dft = dev_alloc<float>(m.owner->owned, 1408LL * 2048);
    launch_dft_matrix(m.dft, s);
    GENIE_CUDA(cudaStreamSynchronize(s));
    m.finalized = true;
  }
};

}  // namespace

void model_finalize(Model& m) {
  Finalizer f(m);
  f.run();
}

void run_linear(Model& m, const Linear& L, const float* x, int ldx, float* y, int ldy, int M, int act,
                const float* res, int ldr, int nt, int ksplit, long long split_stride, const Half2Part* a16,
                const Half2Part* y16) {
  ConvGemm p;
  p.x = x; p.ldx = ldx; p.w = L.w; p.w_f16 = L.w_f16; p.w_co_stride = L.K; p.w_tap_stride = 0;
  p.bias = L.b; p.y = y; p.ldy = ldy; p.Cin = L.K; p.Cout = L.N; p.M = M; p.M_out = M; p.act = act;
  p.res = res; p.ldr = ldr; p.ksplit = ksplit; p.split_stride = split_stride;
  if (M <= (m.use_tc ? m.skinny_max_rows : 128) && skinny_gemm_supported(p)) {   // small decode batch: exact weight-streaming path
    launch_skinny_gemm(p, m.stream);
    return;
  }
  if (m.use_tc && L.tc.hi && M >= m.tc_min_rows) {
    // fp16-exact weights: (x_hi + x_lo) . w keeps the fp32 graphs' token parity
    p.tc_w = L.tc.hi; p.tc_wlo = L.tc.lo; p.tc_kpad = L.tc.kpad; p.tc_split_a = m.lin_single_now ? 0 : 1;
    p.tc_nt = nt; p.ksplit = ksplit; p.split_stride = split_stride;
    if (a16) { p.x16 = a16->hi; p.x16_lo = a16->lo; }
    if (y16) { p.y16 = y16->hi; p.y16_lo = y16->lo; }
    launch_tc_conv_gemm(p, m.tc_err, m.stream);
    return;
  }
  GENIE_CHECK(ksplit == 1, "run_linear: split-K needs the tcgen05 path");
  GENIE_CHECK(!a16 && !y16, "run_linear: fp16 hi/lo hand-over needs the tcgen05 path");
  launch_conv_gemm(p, m.stream);
}

BulkStreamScope::BulkStreamScope(Model& mm) : m(mm), keep(mm.stream) {
  if (!m.stream_bulk) return;
  if (m.partition) { m.partition->tok_bulk.lock(); locked = true; }   // one bulk stage at a time on the bulk SMs
  cudaEventRecord(m.ev_bulk, m.stream);
  cudaStreamWaitEvent(m.stream_bulk, m.ev_bulk, 0);
  m.stream = m.stream_bulk;
}
BulkStreamScope::~BulkStreamScope() {
  if (!m.stream_bulk) return;
  m.stream = keep;
  cudaEventRecord(m.ev_bulk, m.stream_bulk);
  cudaStreamWaitEvent(m.stream, m.ev_bulk, 0);
  if (locked) {
    cudaStreamSynchronize(m.stream_bulk);      // the token passes on when the bulk SMs are really free
    m.partition->tok_bulk.unlock();
  }
}

DecodeTokenScope::DecodeTokenScope(Model& mm) : m(mm) {
  if (m.partition) { m.partition->tok_decode.lock(); locked = true; }
}
DecodeTokenScope::~DecodeTokenScope() {
  if (locked) {
    cudaStreamSynchronize(m.stream);
    m.partition->tok_decode.unlock();
  }
}

// Move the handle's streams into the device's SM partitions (see partition.cu).  Captured decode steps are dropped
// (they were recorded on the old streams); the persistent small-batch kernel sizes its grid for the whole device
// and is switched off on a partitioned handle.
void model_enable_partition(Model& m, int decode_sms) {
  GENIE_CHECK(m.stream_owned, "sm_partition: the handle is bound to a caller-owned stream");
  GENIE_CHECK(m.partition == nullptr, "sm_partition: already partitioned");
  GENIE_CUDA(cudaSetDevice(m.device));
  DevicePartition* p = device_partition(m.device, decode_sms);
  GENIE_CUDA(cudaDeviceSynchronize());
  m.t2s_graphs.reset();
  m.t2s_session.reset();
  int least = 0, greatest = 0;
  GENIE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  for (cudaStream_t* s : {&m.stream, &m.stream2, &m.stream3, &m.stream4}) {
    if (*s) cudaStreamDestroy(*s);
    *s = partition_stream(p, true, greatest);
  }
  if (m.stream_bulk) cudaStreamDestroy(m.stream_bulk);
  m.stream_bulk = partition_stream(p, false, least);
  if (!m.ev_bulk) GENIE_CUDA(cudaEventCreateWithFlags(&m.ev_bulk, cudaEventDisableTiming));
  m.partition = p;
  m.decode_sms = p->decode_sms;
  m.persistent_ok = 0;
  ++m.options_gen;
}

void check_tc_error(Model& m) {
  if (!m.tc_err) return;
  int h = 0;
  GENIE_CUDA(cudaMemcpy(&h, m.tc_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (h) {
    cudaMemset(m.tc_err, 0, sizeof(int));
    if (h == 2) throw Error{"input id out of range (phoneme ids must be < the embedding rows, semantic ids < 1025)"};
    throw Error{"tcgen05 pipeline timed out waiting on an mbarrier (tc_gemm.cu)"};
  }
}

void keep_tensor(Model& m, const char* name, const float* dev, long long n) {
  if (!m.keep) return;
  std::vector<float>& v = m.kept[name];
  v.resize(n);
  GENIE_CUDA(cudaMemcpyAsync(v.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaStreamSynchronize(m.stream));
}

}  // namespace genie
