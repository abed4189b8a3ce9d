// SoVITS decoder (vits_fp32.onnx#[273-8452]) for a ragged batch, channels-last
// fp32 activations, plus the per-reference-audio prompt work that the
// reference recomputes in every call (VQ: t2s_encoder#[2-48]; ref_enc:
// vits#[3-271] / prompt_encoder graph).
#include "model.h"
#include "philox.cuh"
#include <algorithm>
#include <cmath>

namespace genie {
namespace {

struct Seg { const int* off; int B; int maxT; int rows; };

// generic conv launcher on segments
struct ConvOpt {
  int dil = 1; float pre_slope = 1.f; int act = ACT_NONE; float act_slope = 0.f;
  const float* res = nullptr; int ldr = 0; const float* bias2 = nullptr; int ldb2 = 0;
  int accumulate = 0; int co0 = 0; int con = -1; int cin0 = 0; int cin = -1;
  const __half* x16 = nullptr; __half* y16 = nullptr;   // fp16 hand-over inside a resblock pair (tc_halo_conv only)
};
void run_conv(Model& m, const Conv& c, const float* x, int ldx, float* y, int ldy, const Seg& sg, const ConvOpt& o = {}) {
  ConvGemm p;
  const int cin = o.cin < 0 ? c.Cin : o.cin;
  const long long K = (long long)c.k * c.Cin;
  p.x = x + o.cin0; p.ldx = ldx;
  p.w = c.w + (long long)o.co0 * K + o.cin0; p.w_f16 = 0; p.w_co_stride = K; p.w_tap_stride = c.Cin;
  p.bias = c.b ? c.b + o.co0 : nullptr;
  p.bias2 = o.bias2 ? o.bias2 + o.co0 : nullptr; p.ldb2 = o.ldb2;
  p.res = o.res; p.ldr = o.ldr; p.y = y; p.ldy = ldy;
  p.Cin = cin; p.Cout = o.con < 0 ? c.Cout : o.con;
  p.ntaps = c.k; p.in_shift_step = o.dil; p.in_shift0 = -o.dil * (c.k - 1) / 2;
  p.pre_slope = o.pre_slope; p.act = o.act; p.act_slope = o.act_slope; p.accumulate = o.accumulate;
  p.in_off = sg.off; p.out_off = sg.off; p.B = sg.B; p.M = sg.maxT; p.M_out = sg.maxT;
  if (!sg.off) { p.B = 1; p.M = sg.rows; p.M_out = sg.rows; }
  if (m.use_tc && c.tc.hi && sg.off != nullptr && o.cin0 == 0 && cin == c.Cin) {
    p.tc_w = c.tc.hi + (long long)o.co0 * c.tc.kpad; p.tc_kpad = c.tc.kpad;
    if (o.co0 == 0 && p.Cout == c.Cout) p.tc_tiles = c.tc.tiles;
    p.x16 = o.x16; p.y16 = o.y16;
    p.tc_wlo = m.tc_vits >= 3 ? c.tc.lo + (long long)o.co0 * c.tc.kpad : nullptr;
    p.tc_split_a = m.tc_vits >= 2;
    launch_tc_conv_gemm(p, m.tc_err, m.stream);
    return;
  }
  launch_conv_gemm(p, m.stream);
}

void run_convt(Model& m, const ConvT& c, const float* x, float* y, const Seg& in, const Seg& out, float pre_slope) {
  const long long tap_sz = (long long)c.Cout * c.Cin;
  if (c.bias_fused && in.off && out.off) {
    // k == stride, pad == 0: one linear layer, every input row read once (was one launch per phase)
    ConvGemm p;
    p.x = x; p.ldx = c.Cin; p.w = c.w; p.w_f16 = 0; p.w_co_stride = c.Cin; p.w_tap_stride = 0;
    p.bias = c.bias_fused; p.y = y; p.ldy = c.Cout * c.k;
    p.Cin = c.Cin; p.Cout = c.Cout * c.k; p.ntaps = 1; p.in_shift0 = 0; p.in_shift_step = 1;
    p.pre_slope = pre_slope;
    p.in_off = in.off; p.out_off = in.off; p.B = in.B; p.M = in.maxT; p.M_out = in.maxT;   // out.off[b] == k * in.off[b]
    if (m.use_tc && c.tc_fused.hi) {
      p.tc_w = c.tc_fused.hi; p.tc_kpad = c.tc_fused.kpad;
      p.tc_wlo = m.tc_vits >= 3 ? c.tc_fused.lo : nullptr;
      p.tc_split_a = m.tc_vits >= 2;
      launch_tc_conv_gemm(p, m.tc_err, m.stream);
    } else {
      launch_conv_gemm(p, m.stream);
    }
    return;
  }
  for (int r = 0; r < c.stride; ++r) {
    int ntaps = (c.k - r + c.stride - 1) / c.stride;
    if (ntaps <= 0) continue;
    ConvGemm p;
    p.x = x; p.ldx = c.Cin; p.w = c.w + r * tap_sz; p.w_f16 = 0; p.w_co_stride = c.Cin;
    p.w_tap_stride = c.stride * tap_sz; p.bias = c.b; p.y = y; p.ldy = c.Cout;
    p.Cin = c.Cin; p.Cout = c.Cout; p.ntaps = ntaps; p.in_shift0 = 0; p.in_shift_step = -1;
    p.out_mul = c.stride; p.out_add = r - c.pad; p.pre_slope = pre_slope;
    // t_out = q*s + r - pad < T_in*s  =>  q <= T_in - 1 + floor((s - 1 + pad - r) / s)
    p.in_off = in.off; p.out_off = out.off; p.B = in.B; p.M = in.maxT; p.M_out = out.maxT;
    p.q_extra = (c.stride - 1 + c.pad - r) / c.stride;
    if (m.use_tc && r < 10 && c.tc[r].hi) {
      p.tc_w = c.tc[r].hi; p.tc_kpad = c.tc[r].kpad;
      p.tc_wlo = m.tc_vits >= 3 ? c.tc[r].lo : nullptr;
      p.tc_split_a = m.tc_vits >= 2;
      launch_tc_conv_gemm(p, m.tc_err, m.stream);
      continue;
    }
    launch_conv_gemm(p, m.stream);
  }
}

// attentions.Encoder (post-LN) with windowed relative positions, vits#[313-1837]
void run_vits_encoder(Model& m, const VitsEncLayer* L, int n, float* x, const Seg& sg, float* q, float* k, float* v,
                      float* att, float* tmp, float* ff) {
  cudaStream_t s = m.stream;
  const int C = 192;
  for (int i = 0; i < n; ++i) {
    run_conv(m, L[i].q, x, C, q, C, sg);
    run_conv(m, L[i].k, x, C, k, C, sg);
    run_conv(m, L[i].v, x, C, v, C, sg);
    Attn a;
    a.q = q; a.k = k; a.v = v; a.o = att; a.ldq = a.ldk = a.ldv = a.ldo = C;
    a.q_off = sg.off; a.kv_off = sg.off; a.B = sg.B; a.H = 2; a.d = 96; a.max_q = sg.maxT;
    a.scale = 1.0f / std::sqrt(96.0f); a.rel_k = L[i].rel_k; a.rel_v = L[i].rel_v; a.window = 4;
    launch_attention(a, s);
    ConvOpt o; o.res = x; o.ldr = C;
    run_conv(m, L[i].o, att, C, tmp, C, sg, o);
    launch_layernorm(tmp, nullptr, L[i].g1, L[i].b1, x, sg.rows, C, s);
    ConvOpt f1; f1.act = ACT_RELU;
    run_conv(m, L[i].ff1, x, C, ff, 768, sg, f1);
    ConvOpt f2; f2.res = x; f2.ldr = C;
    run_conv(m, L[i].ff2, ff, 768, tmp, C, sg, f2);
    launch_layernorm(tmp, nullptr, L[i].g2, L[i].b2, x, sg.rows, C, s);
  }
}

__global__ void zp_noise_kernel(const float* __restrict__ stats, const float* __restrict__ noise, const int* __restrict__ ids,
                                unsigned long long seed, const int* __restrict__ row2utt,
                                const int* __restrict__ off, float* __restrict__ zp, float scale, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * 192) return;
  int c = (int)(i % 192); int r = (int)(i / 192);
  int b = row2utt[r]; int t = r - off[b]; int T2 = off[b + 1] - off[b];
  float mu = stats[(long long)r * 384 + c], logs = stats[(long long)r * 384 + 192 + c];
  // graph layout of the noise is [192, 2T] per utterance (vits#[6490])
  float nz = noise ? noise[192LL * off[b] + (long long)c * T2 + t]
                   : philox_normal(seed, (uint32_t)(ids ? ids[b] : b), (uint32_t)t, (uint32_t)c, 1u);
  zp[i] = mu + nz * expf(logs) * scale;     // vits#[6491-6495]
}

__global__ void i64_to_i32_kernel(const long long* a, int* b, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) b[i] = (int)a[i];
}

// MelStyleEncoder on a [F,704] magnitude spectrogram -> out[C]  (vits#[109-271])
void run_mel_style(Model& m, const MelStyle& e, const float* mag, int F, float* out) {
  cudaStream_t s = m.stream;
  Workspace& ws = m.ws;
  float* a = ws.get<float>("ms.a", (size_t)F * 128);
  float* b = ws.get<float>("ms.b", (size_t)F * 128);
  float* g = ws.get<float>("ms.g", (size_t)F * 256);
  float* q = ws.get<float>("ms.q", (size_t)F * 128);
  float* k = ws.get<float>("ms.k", (size_t)F * 128);
  float* v = ws.get<float>("ms.v", (size_t)F * 128);
  float* o = ws.get<float>("ms.o", (size_t)F * 128);
  float* fc = ws.get<float>("ms.fc", (size_t)F * e.out_dim);
  run_linear(m, e.fc1, mag, 704, a, 128, F, ACT_MISH);
  run_linear(m, e.fc2, a, 128, b, 128, F, ACT_MISH);
  Seg sg{nullptr, 1, F, F};
  run_conv(m, e.t0, b, 128, g, 256, sg);                 // Conv1dGLU k=5: x + a*sigmoid(b)
  launch_glu_residual(g, 256, b, 128, 128, F, s);
  run_conv(m, e.t1, b, 128, g, 256, sg);
  launch_glu_residual(g, 256, b, 128, 128, F, s);
  run_linear(m, e.wq, b, 128, q, 128, F);
  run_linear(m, e.wk, b, 128, k, 128, F);
  run_linear(m, e.wv, b, 128, v, 128, F);
  Attn at;
  at.q = q; at.k = k; at.v = v; at.o = o; at.ldq = at.ldk = at.ldv = at.ldo = 128;
  at.B = 1; at.H = 2; at.d = 64; at.max_q = F; at.scale = 1.0f / std::sqrt(128.0f);   // temperature sqrt(d_model)
  launch_attention(at, s);
  run_linear(m, e.fo, o, 128, a, 128, F, ACT_NONE, b, 128);        // fc + residual
  run_linear(m, e.fc, a, 128, fc, e.out_dim, F);
  launch_mean_rows(fc, e.out_dim, e.out_dim, F, out, s);
}

}  // namespace

// ---------------------------------------------------------------------------
void prompt_build(Model& m, Prompt& p, const int64_t* ref_seq, int Lr, const float* ref_bert, const float* ssl, int Ts,
                  const float* ref_audio, int n_audio, const float* sv_emb, const float* ge_in, int ge_dim,
                  const float* ge_adv_in) {
  GENIE_CHECK(m.finalized, "model not finalized");
  GENIE_CHECK(Lr > 0 && Ts >= 2, "prompt: empty reference");
  GENIE_CUDA(cudaSetDevice(m.device));
  cudaStream_t s = m.stream;
  Workspace& ws = m.ws;
  p.device = m.device; p.model_uid = m.owner->uid; p.Lr = Lr;
  for (int i = 0; i < Lr; ++i)
    GENIE_CHECK(ref_seq[i] >= 0 && ref_seq[i] < m.text_vocab, "prompt: ref_seq id out of range (phoneme table has " +
                                                                  std::to_string(m.text_vocab) + " rows)");
  p.ref_seq = dev_alloc<long long>(p.owned, Lr);
  GENIE_CUDA(cudaMemcpyAsync(p.ref_seq, ref_seq, Lr * sizeof(long long), cudaMemcpyHostToDevice, s));
  if (ref_bert) {
    p.has_bert = true;
    p.ref_bert = dev_alloc<float>(p.owned, (size_t)Lr * 1024);
    GENIE_CUDA(cudaMemcpyAsync(p.ref_bert, ref_bert, (size_t)Lr * 1024 * 4, cudaMemcpyHostToDevice, s));
  }
  // ---- K2: prompt semantic tokens
  {
    const int M = Ts / 2;
    float* cm = ws.get<float>("pr.ssl_cm", (size_t)768 * Ts);
    float* tm = ws.get<float>("pr.ssl_tm", (size_t)768 * Ts);
    float* h = ws.get<float>("pr.h", (size_t)M * 768);
    float* x2 = ws.get<float>("pr.x2", M);
    float* xe = ws.get<float>("pr.xe", (size_t)M * 1024);
    long long* codes = ws.get<long long>("pr.codes", M);
    GENIE_CUDA(cudaMemcpyAsync(cm, ssl, (size_t)768 * Ts * 4, cudaMemcpyHostToDevice, s));
    launch_transpose(cm, 768, Ts, tm, s);                       // [Ts,768]; row pairs = stride-2 k=2 windows
    Linear L; L.w = m.ssl_vq.w; L.w_f16 = 0; L.b = m.ssl_vq.b; L.N = 768; L.K = 1536;
    run_linear(m, L, tm, 1536, h, 768, M);
    launch_row_sqnorm(h, 768, 768, M, x2, s);
    Linear E; E.w = m.codebook_enc; E.w_f16 = 0; E.b = nullptr; E.N = 1024; E.K = 768;
    run_linear(m, E, h, 768, xe, 1024, M);
    launch_vq_argmax(x2, 0, xe, m.codebook_enc_sq, M, codes, s);
    p.Ly = M;
    p.prompts = dev_alloc<int>(p.owned, M);
    i64_to_i32_kernel<<<(M + 255) / 256, 256, 0, s>>>(codes, p.prompts, M);
    GENIE_LAUNCHED("i64_to_i32");
    p.prompts_host.resize(M);
    GENIE_CUDA(cudaMemcpyAsync(p.prompts_host.data(), codes, M * sizeof(long long), cudaMemcpyDeviceToHost, s));
  }
  // ---- K15: global embedding(s)
  p.ge_dim = m.gin;
  p.ge = dev_alloc<float>(p.owned, m.gin);
  if (ge_in) {
    GENIE_CHECK(ge_dim == m.gin, "prompt: ge has the wrong width for this model");
    GENIE_CUDA(cudaMemcpyAsync(p.ge, ge_in, m.gin * 4, cudaMemcpyHostToDevice, s));
  } else {
    GENIE_CHECK(ref_audio && n_audio >= 2048, "prompt: reference audio too short");
    const int F = n_audio / 640;
    float* au = ws.get<float>("pr.audio", n_audio);
    float* fr = ws.get<float>("pr.frames", (size_t)F * 2048);
    float* ri = ws.get<float>("pr.reim", (size_t)F * 1408);
    float* mg = ws.get<float>("pr.mag", (size_t)F * 704);
    GENIE_CUDA(cudaMemcpyAsync(au, ref_audio, (size_t)n_audio * 4, cudaMemcpyHostToDevice, s));
    launch_stft_frames(au, n_audio, fr, F, s);
    Linear Dm; Dm.w = m.dft; Dm.w_f16 = 0; Dm.b = nullptr; Dm.N = 1408; Dm.K = 2048;
    run_linear(m, Dm, fr, 2048, ri, 1408, F);
    launch_magnitude(ri, mg, F, s);
    run_mel_style(m, m.ref_enc, mg, F, p.ge);
    if (m.v2pp) {
      GENIE_CHECK(sv_emb != nullptr, "prompt: V2ProPlus needs sv_emb");
      float* sv = ws.get<float>("pr.sv", 20480);
      float* svo = ws.get<float>("pr.svo", 1024);
      GENIE_CUDA(cudaMemcpyAsync(sv, sv_emb, 20480 * 4, cudaMemcpyHostToDevice, s));
      run_linear(m, m.sv_emb, sv, 20480, svo, 1024, 1);
      launch_prelu_add(p.ge, svo, m.prelu, 1024, s);          // prompt_encoder#[269-275]
    }
  }
  if (m.v2pp) {
    p.ge_mrte = dev_alloc<float>(p.owned, 512);
    if (ge_adv_in) GENIE_CUDA(cudaMemcpyAsync(p.ge_mrte, ge_adv_in, 512 * 4, cudaMemcpyHostToDevice, s));
    else run_linear(m, m.ge_to512, p.ge, 1024, p.ge_mrte, 512, 1);   // prompt_encoder#[276-280]
  } else {
    p.ge_mrte = p.ge;
  }
  // ---- ge-only conditioning, hoisted: flow cond_layer (vits#[6511] x4) and dec.cond (#[7823])
  p.flow_cond = dev_alloc<float>(p.owned, 4 * 1536);
  Seg one{nullptr, 1, 1, 1};
  for (int f = 0; f < 4; ++f) run_conv(m, m.flow[f].cond, p.ge, m.gin, p.flow_cond + f * 1536, 1536, one);
  p.dec_cond = dev_alloc<float>(p.owned, m.dec_cond.Cout);
  run_conv(m, m.dec_cond, p.ge, m.gin, p.dec_cond, m.dec_cond.Cout, one);
  GENIE_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------
void vits_decode(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len,
                 const int64_t* sem, const int* sem_len, const float* zp_noise, unsigned long long seed,
                 float noise_scale, int io_dev, float* audio, int* audio_len, const int* noise_ids) {
  GENIE_CHECK(m.finalized, "model not finalized");
  GENIE_CHECK(B > 0, "empty batch");
  GENIE_CUDA(cudaSetDevice(m.device));
  const BulkStreamScope bulk(m);             // throughput-bound stage: the handle's low-priority stream
  cudaStream_t s = m.stream;
  Workspace& ws = m.ws;
  if (noise_scale < 0.f) noise_scale = m.noise_scale;
  cudaEvent_t ev0, ev1, ev2;
  GENIE_CUDA(cudaEventCreate(&ev0)); GENIE_CUDA(cudaEventCreate(&ev1)); GENIE_CUDA(cudaEventCreate(&ev2));
  GENIE_CUDA(cudaEventRecord(ev0, s));
  const cudaMemcpyKind in_kind = io_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;

  // ---- segment tables: latent rows (2 per token), text rows, and the 5 generator stages
  const int mult[6] = {1, 10, 80, 160, 320, 640};
  std::vector<int> hoff((size_t)8 * (B + 1), 0);
  int* o2 = hoff.data(); int* oL = o2 + (B + 1); int* oTok = oL + (B + 1);
  int maxT2 = 0, maxL = 0, nsem = 0, ntext = 0;
  for (int b = 0; b < B; ++b) {
    GENIE_CHECK(prompts[b] && prompts[b]->model_uid == m.owner->uid, "prompt does not belong to this model");
    GENIE_CHECK(sem_len[b] > 0 && text_len[b] > 0, "vits: empty utterance");
    o2[b + 1] = o2[b] + 2 * sem_len[b]; oL[b + 1] = oL[b] + text_len[b]; oTok[b + 1] = oTok[b] + sem_len[b];
    maxT2 = std::max(maxT2, 2 * sem_len[b]); maxL = std::max(maxL, text_len[b]);
  }
  const int R2 = o2[B]; const int RL = oL[B]; nsem = oTok[B]; ntext = RL;
  if (!io_dev) {   // host-resident ids are validated here; device-resident ones are clamped + flagged by the gather
    for (int i = 0; i < nsem; ++i) GENIE_CHECK(sem[i] >= 0 && sem[i] < 1024, "vits: semantic id out of range [0, 1024)");
    for (int i = 0; i < ntext; ++i)
      GENIE_CHECK(text_seq[i] >= 0 && text_seq[i] < m.vits_text_vocab, "vits: text_seq id out of range");
  }
  for (int st = 1; st <= 5; ++st)
    for (int b = 0; b <= B; ++b) hoff[(size_t)(2 + st) * (B + 1) + b] = o2[b] * mult[st];
  std::vector<int> r2u(R2);
  for (int b = 0; b < B; ++b) for (int r = o2[b]; r < o2[b + 1]; ++r) r2u[r] = b;
  int* d_off = ws.get<int>("v.off", hoff.size());
  int* d_r2u = ws.get<int>("v.r2u", R2);
  GENIE_CUDA(cudaMemcpyAsync(d_off, hoff.data(), hoff.size() * 4, cudaMemcpyHostToDevice, s));
  GENIE_CUDA(cudaMemcpyAsync(d_r2u, r2u.data(), (size_t)R2 * 4, cudaMemcpyHostToDevice, s));
  Seg s2{d_off, B, maxT2, R2}, sL{d_off + (B + 1), B, maxL, RL};
  Seg sg[6];
  sg[0] = s2;
  for (int st = 1; st <= 5; ++st) sg[st] = Seg{d_off + (size_t)(2 + st) * (B + 1), B, maxT2 * mult[st], R2 * mult[st]};

  // per-utterance conditioning tables gathered from the prompts
  float* GE_M = ws.get<float>("v.ge_m", (size_t)B * 512);
  float* FCOND = ws.get<float>("v.fcond", (size_t)B * 4 * 1536);
  const int C0 = m.dec_pre.Cout;
  float* DCOND = ws.get<float>("v.dcond", (size_t)B * C0);
  for (int b = 0; b < B; ++b) {
    GENIE_CUDA(cudaMemcpyAsync(GE_M + (size_t)b * 512, prompts[b]->ge_mrte, 512 * 4, cudaMemcpyDeviceToDevice, s));
    GENIE_CUDA(cudaMemcpyAsync(FCOND + (size_t)b * 6144, prompts[b]->flow_cond, 6144 * 4, cudaMemcpyDeviceToDevice, s));
    GENIE_CUDA(cudaMemcpyAsync(DCOND + (size_t)b * C0, prompts[b]->dec_cond, (size_t)C0 * 4, cudaMemcpyDeviceToDevice, s));
  }

  long long* SEM = ws.get<long long>("v.sem", nsem);
  long long* TXT = ws.get<long long>("v.txt", ntext);
  GENIE_CUDA(cudaMemcpyAsync(SEM, sem, (size_t)nsem * 8, in_kind, s));
  GENIE_CUDA(cudaMemcpyAsync(TXT, text_seq, (size_t)ntext * 8, in_kind, s));

  const size_t RM = (size_t)std::max(R2, RL);
  float* Q768 = ws.get<float>("v.q768", (size_t)R2 * 768);
  float* Y = ws.get<float>("v.y", (size_t)R2 * 192);
  float* TX = ws.get<float>("v.tx", (size_t)RL * 192);
  float* bq = ws.get<float>("v.bq", RM * 512);
  float* bk = ws.get<float>("v.bk", RM * 512);
  float* bv = ws.get<float>("v.bv", RM * 512);
  float* batt = ws.get<float>("v.att", RM * 512);
  float* btmp = ws.get<float>("v.tmp", RM * 512);
  float* bff = ws.get<float>("v.ff", RM * 768);

  // ---- K9: codebook dequant + x2 nearest upsample (vits#[273-292]); ssl_proj
  launch_gather_rows(Q768, 768, m.codebook, 768, SEM, nsem, 2, s, 1024, m.tc_err);
  run_conv(m, m.ssl_proj, Q768, 768, Y, 192, s2);
  // ---- K10: encoder_ssl (3), text embedding + encoder_text (6)
  run_vits_encoder(m, m.enc_ssl, 3, Y, s2, bq, bk, bv, batt, btmp, bff);
  launch_gather_rows(TX, 192, m.vits_text_emb, 192, TXT, ntext, 1, s, m.vits_text_vocab, m.tc_err);
  run_vits_encoder(m, m.enc_text, 6, TX, sL, bq, bk, bv, batt, btmp, bff);
  // ---- K11: MRTE (vits#[4891-4964])
  {
    float* S512 = bff;                                   // c_pre(y) [R2,512]  (bff is >= RM*768)
    float* T512 = ws.get<float>("v.t512", (size_t)RL * 512);
    run_conv(m, m.mrte_c_pre, Y, 192, S512, 512, s2);
    run_conv(m, m.mrte_text_pre, TX, 192, T512, 512, sL);
    run_conv(m, m.mrte_q, S512, 512, bq, 512, s2);
    run_conv(m, m.mrte_k, T512, 512, bk, 512, sL);
    run_conv(m, m.mrte_v, T512, 512, bv, 512, sL);
    Attn a;
    a.q = bq; a.k = bk; a.v = bv; a.o = batt; a.ldq = a.ldk = a.ldv = a.ldo = 512;
    a.q_off = s2.off; a.kv_off = sL.off; a.B = B; a.H = 4; a.d = 128; a.max_q = maxT2;
    a.scale = 1.0f / std::sqrt(128.0f);
    launch_attention(a, s);
    ConvOpt o; o.res = S512; o.ldr = 512; o.bias2 = GE_M; o.ldb2 = 512;     // + c_pre + ge, #[4960-4961]
    run_conv(m, m.mrte_o, batt, 512, btmp, 512, s2, o);
    run_conv(m, m.mrte_c_post, btmp, 512, Y, 192, s2);
  }
  // ---- encoder2 (3), proj, z_p (vits#[4965-6495])
  run_vits_encoder(m, m.enc2, 3, Y, s2, bq, bk, bv, batt, btmp, bff);
  float* STATS = bq;                                     // [R2,384]
  run_conv(m, m.enc_proj, Y, 192, STATS, 384, s2);
  keep_tensor(m, "stats", STATS, (long long)R2 * 384);
  float* Z = ws.get<float>("v.z", (size_t)R2 * 192);
  float* Z2 = ws.get<float>("v.z2", (size_t)R2 * 192);
  {
    const float* nz = nullptr;
    if (zp_noise) {
      float* NZ = ws.get<float>("v.noise", (size_t)R2 * 192);
      GENIE_CUDA(cudaMemcpyAsync(NZ, zp_noise, (size_t)R2 * 192 * 4, in_kind, s));
      nz = NZ;
    }
    const int* d_ids = nullptr;
    if (noise_ids) {                                   // Philox stream per utterance: its index in ITS OWN batch
      int* ids = ws.get<int>("v.noise_ids", B);
      GENIE_CUDA(cudaMemcpyAsync(ids, noise_ids, (size_t)B * 4, cudaMemcpyHostToDevice, s));
      d_ids = ids;
    }
    zp_noise_kernel<<<(unsigned)(((long long)R2 * 192 + 255) / 256), 256, 0, s>>>(STATS, nz, d_ids, seed, d_r2u, s2.off,
                                                                                 Z, noise_scale, R2);
    GENIE_LAUNCHED("zp_noise");
  }
  // ---- K13: flow reverse (vits#[6500-7820])
  {
    float* Hh = bk; float* XIN = bv; float* ACT = batt; float* OUT = btmp; float* MEAN = bff;
    for (int f = 0; f < 4; ++f) {
      const FlowStep& F = m.flow[f];
      launch_flip_channels(Z, Z2, 192, R2, s);
      std::swap(Z, Z2);
      ConvOpt po; po.cin = 96;                           // pre on x0 = z[:, :96]
      run_conv(m, F.pre, Z, 192, Hh, 192, s2, po);
      for (int l = 0; l < 4; ++l) {
        ConvOpt io; io.bias2 = FCOND + f * 1536 + l * 384; io.ldb2 = 6144;
        run_conv(m, F.wn[l].in, Hh, 192, XIN, 384, s2, io);
        launch_gated_act(XIN, 384, ACT, 192, 192, R2, s);
        if (l < 3) {
          ConvOpt r1; r1.con = 192; r1.res = Hh; r1.ldr = 192;               // h += rs[:, :192]
          ConvOpt r2; r2.co0 = 192; r2.con = 192; r2.accumulate = l > 0;      // out += rs[:, 192:]
          run_conv(m, F.wn[l].rs, ACT, 192, OUT, 192, s2, r2);
          run_conv(m, F.wn[l].rs, ACT, 192, Hh, 192, s2, r1);
        } else {
          ConvOpt r2; r2.accumulate = 1;
          run_conv(m, F.wn[l].rs, ACT, 192, OUT, 192, s2, r2);
        }
      }
      run_conv(m, F.post, OUT, 192, MEAN, 96, s2);
      launch_sub_cols(Z, 192, 96, MEAN, 96, 96, R2, s);
    }
  }
  keep_tensor(m, "z", Z, (long long)R2 * 192);
  // ---- K14: HiFi-GAN generator (vits#[7822-8452])
  const size_t gen_floats = (size_t)R2 * 640 * m.c_last;   // every stage output has R2*10240 (V2) floats at most
  const size_t s0_floats = (size_t)R2 * C0;
  float* GX = ws.get<float>("v.gx", std::max(gen_floats, s0_floats));    // running stage input / xs
  GENIE_CUDA(cudaEventRecord(ev1, s));
  const unsigned long long launches_before_gen = g_launches;
  float* UP = ws.get<float>("v.up", gen_floats);
  float* GA = ws.get<float>("v.ga", gen_floats);
  __half* GA16 = ws.get<__half>("v.ga16", gen_floats);
  float* GB = ws.get<float>("v.gb", gen_floats);
  float* GC = ws.get<float>("v.gc", gen_floats);
  {
    ConvOpt o; o.bias2 = DCOND; o.ldb2 = C0;
    run_conv(m, m.dec_pre, Z, 192, GX, C0, s2, o);
    keep_tensor(m, "g_pre", GX, (long long)R2 * C0);
  }
  cudaEvent_t ev_narrow = nullptr;           // start of the first stage with <= 32 channels (HBM-bound part)
  double narrow_bytes = 0.0;                 // its algorithmic bytes: every tensor read or written once per conv
  for (int i = 0; i < m.n_up; ++i) {
    const ConvT& U = m.ups[i];
    if (U.Cout <= 32 && !ev_narrow) {
      GENIE_CUDA(cudaEventCreate(&ev_narrow));
      GENIE_CUDA(cudaEventRecord(ev_narrow, s));
    }
    if (ev_narrow) {
      const double e_in = (double)sg[i].rows * U.Cin, e_out = (double)sg[i + 1].rows * U.Cout;
      const bool h16p = m.use_tc && m.tc_vits == 1 && tc_halo_fp16_pair_ok(U.Cout, 3);
      // upsampling conv: read in, write out; per resblock pair: conv1 reads x (4 B) and writes the hand-over
      // (2 B fp16 / 4 B fp32), conv2 reads it back, reads the residual and writes (4 + 4 B); the last conv of
      // resblocks 1 and 2 also reads the running sum
      // (fused pairs, tc_pair_conv.cu: x in once, y out once = 8 B per element and pair)
      const bool fusedp = m.use_tc && m.tc_vits == 1 && m.fuse_pairs && tc_pair_conv_supported(U.Cout, 3);
      narrow_bytes += 4.0 * (e_in + e_out) +
                      e_out * (9.0 * (fusedp ? 8.0 : 4.0 + 4.0 + 4.0 + (h16p ? 4.0 : 8.0)) + 2.0 * 4.0);
    }
    run_convt(m, U, GX, UP, sg[i], sg[i + 1], 0.1f);
    const Seg& S = sg[i + 1];
    const int C = U.Cout;
    const int dils[3] = {1, 3, 5};
    for (int j = 0; j < 3; ++j) {
      const ResBlock& Rb = m.res[i * 3 + j];
      const float* r = UP;
      float* pp[2] = {GB, GC};
      // narrow stages are HBM-bound: conv1 hands fp16(lrelu(out)) to conv2 (exactly what conv2's loader
      // would have produced from the fp32 tensor), 2 + 2 instead of 4 + 4 bytes per element
      const bool h16 = m.use_tc && m.tc_vits == 1 && S.off != nullptr && tc_halo_fp16_pair_ok(C, Rb.k);
      // 16 / 32 channels: the whole pair in one kernel (x read once, intermediate never leaves the SM)
      const bool fused = m.use_tc && m.tc_vits == 1 && S.off != nullptr && m.fuse_pairs && tc_pair_conv_supported(C, Rb.c2[0].k) &&
                         Rb.c1[0].tc.hi && Rb.c2[0].tc.hi;
      for (int c = 0; c < 3 && fused; ++c) {
        float* dst = c < 2 ? pp[c] : GX;
        ConvGemm q;
        q.x = r; q.ldx = C; q.pre_slope = 0.1f; q.Cin = C; q.Cout = C; q.ntaps = Rb.c2[c].k; q.in_shift_step = 1;
        q.tc_w = Rb.c2[c].tc.hi; q.tc_kpad = Rb.c2[c].tc.kpad; q.bias = Rb.c2[c].b;
        q.res = r; q.ldr = C; q.y = dst; q.ldy = C; q.accumulate = (c == 2 && j > 0) ? 1 : 0;
        q.in_off = S.off; q.out_off = S.off; q.B = S.B; q.M = S.maxT; q.M_out = S.maxT;
        launch_tc_pair_conv(q, Rb.c1[c].tc.hi, Rb.c1[c].b, Rb.c1[c].tc.kpad, dils[c], m.tc_err, m.stream);
        r = dst;
      }
      for (int c = 0; c < 3 && !fused; ++c) {
        ConvOpt a; a.dil = dils[c]; a.pre_slope = 0.1f;
        if (h16) { a.act = ACT_LRELU; a.act_slope = 0.1f; a.y16 = GA16; }
        run_conv(m, Rb.c1[c], r, C, GA, C, S, a);
        ConvOpt b2; b2.pre_slope = h16 ? 1.f : 0.1f; b2.res = r; b2.ldr = C;
        if (h16) b2.x16 = GA16;
        if (c < 2) {
          run_conv(m, Rb.c2[c], GA, C, pp[c], C, S, b2);
          r = pp[c];
        } else {
          b2.accumulate = j > 0;                         // xs = r0 + r1 + r2 (the /3 lives in the next weights)
          run_conv(m, Rb.c2[c], GA, C, GX, C, S, b2);
        }
      }
    }
    if (m.keep) {
      keep_tensor(m, ("g_up" + std::to_string(i)).c_str(), UP, (long long)S.rows * C);
      keep_tensor(m, ("g_s" + std::to_string(i)).c_str(), GX, (long long)S.rows * C);
    }
  }
  float* AUD = ws.get<float>("v.audio", (size_t)R2 * 640);
  launch_conv_post_tanh(GX, m.c_last, m.conv_post, AUD, sg[5].off, B, sg[5].maxT, s);
  GENIE_CUDA(cudaEventRecord(ev2, s));
  const unsigned long long gen_launches = g_launches - launches_before_gen;
  if (audio)
    GENIE_CUDA(cudaMemcpyAsync(audio, AUD, (size_t)R2 * 640 * 4, io_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  GENIE_CUDA(cudaStreamSynchronize(s));
  check_tc_error(m);
  if (audio_len) for (int b = 0; b < B; ++b) audio_len[b] = sem_len[b] * 1280;
  float t01 = 0.f, t12 = 0.f;
  cudaEventElapsedTime(&t01, ev0, ev1); cudaEventElapsedTime(&t12, ev1, ev2);
  m.timing[4] = t01 + t12; m.timing[5] = t12; m.timing[6] = (float)gen_launches; m.timing[7] = (float)R2;
  m.timing[10] = 0.f; m.timing[11] = 0.f;
  if (ev_narrow) {
    // narrow generator stages incl. conv_post (reads the last stage output once, writes the waveform)
    narrow_bytes += 4.0 * ((double)sg[5].rows * m.c_last + (double)sg[5].rows);
    cudaEventElapsedTime(&m.timing[10], ev_narrow, ev2);
    m.timing[11] = (float)(narrow_bytes / 1e6);
    cudaEventDestroy(ev_narrow);
  }
  cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
}

}  // namespace genie
