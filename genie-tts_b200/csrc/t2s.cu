// T2S GPT stage: text encode (t2s_encoder#[49-83]) -> prefill
// (t2s_first_stage_decoder) -> decode loop (t2s_stage_decoder x <=500,
// reference src/genie_tts/Core/Inference.py:76-106) for a ragged batch of
// independent utterances.  KV cache: fp32, head-major, appended in place.
#include "model.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace genie {
namespace {

constexpr int D = 512, NL = 24, H = 16, V = 1025;

struct Batch {
  int B = 0;
  std::vector<int> Lr, Lt, Ly, Lx, S, row_off, txt_off;
  int rows = 0, txt_rows = 0, maxS = 0, cap = 0, hist_ld = 0;
};

__global__ void fill_rows_kernel(float* x, const float* bias, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * 512) return;
  x[i] = bias ? bias[i & 511] : 0.f;
}
__global__ void init_hist_kernel(int* hist, int hist_ld, const int* const* prompt_ptrs, const int* ly, int B) {
  int b = blockIdx.x;
  for (int i = threadIdx.x; i < ly[b]; i += blockDim.x) hist[(long long)b * hist_ld + i] = prompt_ptrs[b][i];
}
__global__ void hist_to_i64_kernel(const int* hist, int hist_ld, const int* hist_len, long long* y, int y_ld, int B) {
  int b = blockIdx.x;
  int n = hist_len[b];
  for (int i = threadIdx.x; i < y_ld; i += blockDim.x)
    y[(long long)b * y_ld + i] = i < n ? hist[(long long)b * hist_ld + i] : 0;
}

struct StepBufs {
  float *h, *qkv, *att, *tmp, *h1, *ff, *logits, *part, *part2, *ppart;
  int *hist, *hist_len, *kv_len, *active, *stop_step;
  float* kv; long long utt_stride, layer_stride, v_off; int cap, hist_ld; long long part_stride;
  int utt_base;   // first utterance of this branch within the batch
};

// State of one batch between genie_t2s_prefill and genie_t2s_read: everything the decode loop and the result
// read-out need.  The buffers live in the model's workspace, so a model has one live session at a time.
struct T2SSession {
  Batch bt; SamplingCfg cfg{}; int B = 0, max_steps = 0;
  float* LOGITS = nullptr; int* HIST = nullptr; long long* Y64 = nullptr;
  int *d_histlen = nullptr, *d_kvlen = nullptr, *d_active = nullptr, *d_stop = nullptr;
  StepBufs w{};
  bool can_graph = false;
  int steps_done = 0;
  bool all_stopped = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;                       // prefill
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> decode_spans;  // one pair per decode_steps call
  ~T2SSession() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto& e : decode_spans) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  }
};

T2SSession& session_of(Model& m) {
  GENIE_CHECK(m.t2s_session != nullptr, "no T2S batch in flight: call genie_t2s_prefill first");
  return *static_cast<T2SSession*>(m.t2s_session.get());
}

void record_step_logits(Model& m, const T2SSession& S) {
  if (!m.record_logits) return;
  std::vector<float>& rec = m.logits_host;
  const size_t o = rec.size();
  rec.resize(o + (size_t)S.B * V);
  GENIE_CUDA(cudaMemcpyAsync(rec.data() + o, S.LOGITS, (size_t)S.B * V * 4, cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaStreamSynchronize(m.stream));
}

// pointer table / barrier words of the persistent step (allocated outside any stream capture)
void ensure_persistent_step(Model& m) {
  if (!m.step_layers_dev) {
    std::vector<StepLayerPtrs> hl(NL);
    for (int l = 0; l < NL; ++l) {
      const T2SLayer& L = m.layers[l];
      hl[l] = StepLayerPtrs{reinterpret_cast<const __half*>(L.qkv.w), reinterpret_cast<const __half*>(L.out.w),
                            reinterpret_cast<const __half*>(L.ff1.w), reinterpret_cast<const __half*>(L.ff2.w),
                            L.qkv.b, L.out.b, L.ff1.b, L.ff2.b, L.ln1_g, L.ln1_b, L.ln2_g, L.ln2_b};
    }
    StepLayerPtrs* d = nullptr;
    GENIE_CUDA(cudaMalloc(&d, sizeof(StepLayerPtrs) * NL));
    GENIE_CUDA(cudaMemcpy(d, hl.data(), sizeof(StepLayerPtrs) * NL, cudaMemcpyHostToDevice));
    m.owned.push_back(d);
    m.step_layers_dev = d;
    GENIE_CUDA(cudaMalloc(&m.step_sync, 2 * sizeof(unsigned)));
    GENIE_CUDA(cudaMemset(m.step_sync, 0, 2 * sizeof(unsigned)));
    m.owned.push_back(m.step_sync);
    GENIE_CUDA(cudaDeviceGetAttribute(&m.num_sms, cudaDevAttrMultiProcessorCount, m.device));
  }
}

// one decode step for every active utterance (stage#[12-1821])
void decode_step(Model& m, const StepBufs& w, int B, const SamplingCfg& cfg) {
  cudaStream_t s = m.stream;
  struct PdlScope { int prev; explicit PdlScope(int on) : prev(g_pdl_now) { g_pdl_now = on; } ~PdlScope() { g_pdl_now = prev; } };
  const PdlScope pdl_scope(B > m.skinny_max_rows ? 1 : 0);
  launch_decode_embed(w.h, w.hist, w.hist_ld, w.hist_len, w.active, m.audio_emb, m.audio_alpha, m.div_term, B, s);
  const float scale = 1.0f / std::sqrt(32.0f);
  // measured (90 steps): batch 1 / 2 / 4 / 8 = 37 / 43 / 53 / 77 ms persistent vs 58 / - / 60 / 70 ms kernel chain
  const bool persistent = m.persistent_step && B <= m.persistent_step && B <= m.skinny_max_rows && m.layers[0].qkv.w_f16 &&
                          m.predict.w_f16 && w.ppart != nullptr && m.step_layers_dev != nullptr;
  if (persistent) {
    // batch <= 8: all 24 layers + logits in one resident kernel (t2s_persistent.cu)
    PersistentStep a;
    a.layers = reinterpret_cast<const StepLayerPtrs*>(m.step_layers_dev); a.n_layers = NL;
    a.wpredict = reinterpret_cast<const __half*>(m.predict.w); a.bpredict = m.predict.b; a.vocab = V;
    a.h = w.h; a.qkv = w.qkv; a.part = w.ppart; a.lnin = w.tmp; a.lnin2 = w.part2; a.h1 = w.h1; a.ff = w.ff;
    a.logits = w.logits; a.ld_logits = V;
    a.kv = w.kv; a.utt_stride = w.utt_stride; a.layer_stride = w.layer_stride; a.v_off = w.v_off; a.cap = w.cap;
    a.kv_len = w.kv_len; a.active = w.active; a.sync = m.step_sync;
    a.B = B; a.nch = persistent_step_chunks(B, m.num_sms); a.scale = scale;
    launch_t2s_step_persistent(a, m.num_sms, s);
  }
  for (int l = 0; l < (persistent ? 0 : NL); ++l) {
    const T2SLayer& L = m.layers[l];
    const long long ps = w.part_stride;   // rows of the WHOLE batch
    const bool small_tc = m.use_tc && B > m.skinny_max_rows && B <= 128;
    const int r0 = 0;   // partial-sum buffers belong to the branch (see the capture code): rows start at 0
    if (small_tc) {
      // single-shot tcgen05 GEMMs (tc_small_gemm.cu): every output is a split-K partial; bias, residual,
      // activation and the KV append are applied by the consumer kernels
      auto gemm = [&](const Linear& Lw, const float* x, int ldx, int a_ns, const float* a_bias, int a_relu, int nt) {
        SmallGemm g;
        g.x = x; g.ldx = ldx; g.a_nsplit = a_ns; g.a_stride = ps * (ldx / D); g.a_bias = a_bias; g.a_relu = a_relu;
        g.w = reinterpret_cast<const __half*>(Lw.w); g.ldw = Lw.K; g.N = Lw.N; g.K = Lw.K; g.M = B;
        g.y = w.part + (size_t)r0 * Lw.N; g.ldy = Lw.N; g.split_stride = ps * (Lw.N / D);
        launch_tc_small_gemm(g, nt, m.tc_err, s);
      };
      gemm(L.qkv, w.h, D, 1, nullptr, 0, 32);                                   // 2 partials [B,1536]
      launch_decode_attention_fused(w.part + (size_t)r0 * 3 * D, 2, ps * 3, L.qkv.b, w.att, w.kv, w.utt_stride, l * w.layer_stride, w.v_off,
                                    w.kv_len, w.active, B, w.cap, scale, s);
      gemm(L.out, w.att, D, 1, nullptr, 0, 32);                                 // 2 partials [B,512]
      launch_layernorm(w.part + (size_t)r0 * D, w.h, L.ln1_g, L.ln1_b, w.h1, B, D, s, 2, ps, L.out.b);
      gemm(L.ff1, w.h1, D, 1, nullptr, 0, 32);                                  // 2 partials [B,2048]
      // FFN2 reads relu(sum FFN1 partials + bias) as its A operand: partials live in w.part, so its own
      // output partials go to w.part2
      {
        SmallGemm g;
        g.x = w.part + (size_t)r0 * 4 * D; g.ldx = 4 * D; g.a_nsplit = 2; g.a_stride = ps * 4; g.a_bias = L.ff1.b; g.a_relu = 1;
        g.w = reinterpret_cast<const __half*>(L.ff2.w); g.ldw = L.ff2.K; g.N = D; g.K = 4 * D; g.M = B;
        g.y = w.part2 + (size_t)r0 * D; g.ldy = D; g.split_stride = ps;
        launch_tc_small_gemm(g, 32, m.tc_err, s);                               // 8 partials [B,512]
      }
      launch_layernorm(w.part2 + (size_t)r0 * D, w.h1, L.ln2_g, L.ln2_b, w.h, B, D, s, 8, ps, L.ff2.b);
      continue;
    }
    const bool sk = true;   // B <= skinny_max_rows (or tcgen05 disabled): weight-streaming SIMT GEMV path
    const bool tc = B <= 128;
    const int nt_w = 0, nt_s = 0, ks_out = 2;
    (void)sk;
    run_linear(m, L.qkv, w.h, D, w.qkv, 3 * D, B, ACT_NONE, nullptr, 0, tc ? nt_w : 0);
    // q / k_new / v_new straight from the finished QKV rows: cache append + attention in one kernel
    launch_decode_attention_fused(w.qkv, 1, 0, nullptr, w.att, w.kv, w.utt_stride, l * w.layer_stride, w.v_off,
                                  w.kv_len, w.active, B, w.cap, scale, s);
    if (tc) {
      run_linear(m, L.out, w.att, D, w.part, D, B, ACT_NONE, nullptr, 0, nt_s, ks_out, ps);
      launch_layernorm(w.part, w.h, L.ln1_g, L.ln1_b, w.h1, B, D, s, ks_out, ps, L.out.b);
      run_linear(m, L.ff1, w.h1, D, w.ff, 4 * D, B, ACT_RELU, nullptr, 0, nt_w);
      run_linear(m, L.ff2, w.ff, 4 * D, w.part, D, B, ACT_NONE, nullptr, 0, nt_s, 8, ps);
      launch_layernorm(w.part, w.h1, L.ln2_g, L.ln2_b, w.h, B, D, s, 8, ps, L.ff2.b);
    } else {
      run_linear(m, L.out, w.att, D, w.tmp, D, B, ACT_NONE, w.h, D);
      launch_layernorm(w.tmp, nullptr, L.ln1_g, L.ln1_b, w.h1, B, D, s);
      run_linear(m, L.ff1, w.h1, D, w.ff, 4 * D, B, ACT_RELU);
      run_linear(m, L.ff2, w.ff, 4 * D, w.tmp, D, B, ACT_NONE, w.h1, D);
      launch_layernorm(w.tmp, nullptr, L.ln2_g, L.ln2_b, w.h, B, D, s);
    }
  }
  if (!persistent) run_linear(m, m.predict, w.h, D, w.logits, V, B);
  SamplerArgs a{};
  a.logits = w.logits; a.ld = V; a.hist = w.hist; a.hist_ld = w.hist_ld; a.hist_len = w.hist_len;
  a.kv_len = w.kv_len; a.active = w.active; a.stop_step = w.stop_step; a.B = B; a.utt_base = w.utt_base;
  a.top_k = cfg.top_k; a.temperature = cfg.temperature; a.penalty = cfg.penalty; a.greedy = cfg.greedy;
  a.seed = cfg.seed; a.step = 0; a.honour_stop = cfg.fixed_steps > 0 ? 0 : 1; a.advance_kv = 1; a.check_stop = 1;
  a.dbg_noise = nullptr;
  launch_sampler(a, s);
}

}  // namespace

// encoder + first-stage graph for a batch: embeddings, 24 prefill layers into the KV cache, first sampled token;
// leaves the decode-step buffers and graph ready (Inference.py:76-93)
void t2s_prefill(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len_in,
                 const float* text_bert, const SamplingCfg& cfg, int io_dev) {
  GENIE_CHECK(m.finalized, "model not finalized");
  GENIE_CHECK(B > 0, "empty batch");
  cudaStream_t s = m.stream;
  GENIE_CUDA(cudaSetDevice(m.device));
  const int max_steps = cfg.fixed_steps > 0 ? cfg.fixed_steps : cfg.max_steps;
  m.t2s_session.reset();
  std::shared_ptr<T2SSession> sess_ptr = std::make_shared<T2SSession>();
  T2SSession& S = *sess_ptr;
  S.cfg = cfg; S.B = B; S.max_steps = max_steps;

  // ---- text lengths (device-resident payload: lengths are still host metadata)
  std::vector<int> text_len(text_len_in, text_len_in + B);
  Batch& bt = S.bt; bt.B = B;
  bt.row_off.push_back(0); bt.txt_off.push_back(0);
  bool any_bert = text_bert != nullptr;
  for (int b = 0; b < B; ++b) {
    Prompt* p = prompts[b];
    GENIE_CHECK(p && p->model == &m, "prompt does not belong to this model");
    GENIE_CHECK(text_len[b] > 0, "empty text_seq");
    bt.Lr.push_back(p->Lr); bt.Lt.push_back(text_len[b]); bt.Ly.push_back(p->Ly);
    bt.Lx.push_back(p->Lr + text_len[b]); bt.S.push_back(p->Lr + text_len[b] + p->Ly);
    bt.row_off.push_back(bt.row_off.back() + bt.S.back());
    bt.txt_off.push_back(bt.txt_off.back() + bt.Lx.back());
    bt.maxS = std::max(bt.maxS, bt.S.back());
    any_bert = any_bert || p->has_bert;
  }
  bt.rows = bt.row_off.back(); bt.txt_rows = bt.txt_off.back();
  bt.cap = ((bt.maxS + max_steps + 1 + 15) / 16) * 16;
  int maxLy = *std::max_element(bt.Ly.begin(), bt.Ly.end());
  bt.hist_ld = maxLy + max_steps + 2;

  Workspace& ws = m.ws;
  const int R = bt.rows;
  float* X = ws.get<float>("t2s.x", (size_t)R * D);
  float* QKV = ws.get<float>("t2s.qkv", (size_t)R * 3 * D);
  float* ATT = ws.get<float>("t2s.att", (size_t)R * D);
  float* TMP = ws.get<float>("t2s.tmp", (size_t)R * D);
  float* H1 = ws.get<float>("t2s.h1", (size_t)R * D);
  float* FF = ws.get<float>("t2s.ff", (size_t)R * 4 * D);
  // the same bytes viewed as two fp16 matrices [R, 2048] (hi | lo) when the FFN pair hands over in fp16
  __half* FF16 = reinterpret_cast<__half*>(FF);
  static const bool ffn16_env = [] { const char* e = getenv("GENIE_FFN16"); return !(e && e[0] == '0'); }();
  const bool ffn16 = ffn16_env && m.use_tc && R >= m.tc_min_rows && R > m.skinny_max_rows && m.layers[0].ff1.tc.hi &&
                     m.layers[0].ff2.tc.hi && !m.layers[0].ff2.tc.lo;
  float* XT = ws.get<float>("t2s.xtext", (size_t)bt.txt_rows * D);
  float* BERT = any_bert ? ws.get<float>("t2s.bert", (size_t)bt.txt_rows * 1024) : nullptr;
  long long* SEQ = ws.get<long long>("t2s.seq", bt.txt_rows);
  int* IMETA = ws.get<int>("t2s.imeta", (size_t)R * 2 + bt.txt_rows + 8 * (B + 1));
  float* LAST = ws.get<float>("t2s.last", (size_t)B * D);
  float* LOGITS = ws.get<float>("t2s.logits", (size_t)B * V);
  long long* LASTIDX = ws.get<long long>("t2s.lastidx", B);
  const int** PPTR = ws.get<const int*>("t2s.pptr", B);
  int* HIST = ws.get<int>("t2s.hist", (size_t)B * bt.hist_ld);
  long long* Y64 = ws.get<long long>("t2s.y64", (size_t)B * bt.hist_ld);
  // KV cache [B][layer][K|V][H][cap][32] fp32
  const long long head_sz = (long long)bt.cap * 32, v_off = H * head_sz, layer_stride = 2 * v_off,
                  utt_stride = NL * layer_stride;
  float* KV = ws.get<float>("t2s.kv", (size_t)B * utt_stride);

  // ---- host-built index metadata, one upload
  //   txt_pos[txt_rows] | row2utt[R] | aud_pos[R] (per audio row; text rows unused) | per-utt arrays
  std::vector<int> meta((size_t)R * 2 + bt.txt_rows + 8 * (B + 1), 0);
  // per-utterance state first: its device addresses (baked into the decode-step graph) then
  // depend only on the buffer base and B, not on this call's row counts
  int* h_row_off = meta.data();                 // [B+1]
  int* h_lx = h_row_off + (B + 1);              // [B]
  int* h_zero = h_lx + (B + 1);                 // [B] zeros (dst_pos0 for prefill scatter)
  int* h_kvlen = h_zero + (B + 1);              // [B]
  int* h_histlen = h_kvlen + (B + 1);           // [B]
  int* h_active = h_histlen + (B + 1);          // [B]
  int* h_stop = h_active + (B + 1);             // [B]
  int* h_ly = h_stop + (B + 1);                 // [B]
  int* h_txt_pos = h_ly + (B + 1);
  int* h_row2utt = h_txt_pos + bt.txt_rows;
  int* h_aud_tok_pos = h_row2utt + R;           // position (1-based) of audio rows
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < bt.Lx[b]; ++i) h_txt_pos[bt.txt_off[b] + i] = i + 1;
    for (int i = 0; i < bt.S[b]; ++i) h_row2utt[bt.row_off[b] + i] = b;
    for (int i = 0; i < bt.Ly[b]; ++i) h_aud_tok_pos[bt.row_off[b] + bt.Lx[b] + i] = i + 1;
    h_row_off[b] = bt.row_off[b]; h_lx[b] = bt.Lx[b]; h_kvlen[b] = bt.S[b]; h_histlen[b] = bt.Ly[b];
    h_active[b] = 1; h_stop[b] = -1; h_ly[b] = bt.Ly[b];
  }
  h_row_off[B] = R;
  GENIE_CUDA(cudaMemcpyAsync(IMETA, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  int* d_row_off = IMETA; int* d_lx = d_row_off + (B + 1); int* d_zero = d_lx + (B + 1);
  int* d_kvlen = d_zero + (B + 1); int* d_histlen = d_kvlen + (B + 1); int* d_active = d_histlen + (B + 1);
  int* d_stop = d_active + (B + 1); int* d_ly = d_stop + (B + 1);
  int* d_txt_pos = d_ly + (B + 1); int* d_row2utt = d_txt_pos + bt.txt_rows; int* d_aud_pos = d_row2utt + R;

  GENIE_CUDA(cudaEventCreate(&S.ev0)); GENIE_CUDA(cudaEventCreate(&S.ev1));
  cudaEvent_t ev0 = S.ev0, ev1 = S.ev1;
  GENIE_CUDA(cudaEventRecord(ev0, s));

  // ---- K1: x = Emb_text[ref||text] + bert_proj(bert) + alpha*PE(1..Lx)   (t2s_encoder#[49-83])
  const cudaMemcpyKind in_kind = io_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  {
    long long toff = 0;
    for (int b = 0; b < B; ++b) {
      Prompt* p = prompts[b];
      GENIE_CUDA(cudaMemcpyAsync(SEQ + bt.txt_off[b], p->ref_seq, p->Lr * sizeof(long long), cudaMemcpyDeviceToDevice, s));
      GENIE_CUDA(cudaMemcpyAsync(SEQ + bt.txt_off[b] + p->Lr, text_seq + toff, text_len[b] * sizeof(long long), in_kind, s));
      if (any_bert) {
        float* dst = BERT + (long long)bt.txt_off[b] * 1024;
        if (p->has_bert) GENIE_CUDA(cudaMemcpyAsync(dst, p->ref_bert, (size_t)p->Lr * 1024 * 4, cudaMemcpyDeviceToDevice, s));
        else GENIE_CUDA(cudaMemsetAsync(dst, 0, (size_t)p->Lr * 1024 * 4, s));
        dst += (long long)p->Lr * 1024;
        if (text_bert) GENIE_CUDA(cudaMemcpyAsync(dst, text_bert + toff * 1024, (size_t)text_len[b] * 1024 * 4, in_kind, s));
        else GENIE_CUDA(cudaMemsetAsync(dst, 0, (size_t)text_len[b] * 1024 * 4, s));
      }
      toff += text_len[b];
    }
  }
  if (any_bert) {
    run_linear(m, m.bert_proj, BERT, 1024, XT, D, bt.txt_rows);
  } else {
    // bert features are zeros for ja/en (reference GetPhonesAndBert.py:60,80): bert_proj reduces to its bias
    fill_rows_kernel<<<(unsigned)(((long long)bt.txt_rows * 512 + 255) / 256), 256, 0, s>>>(XT, m.bert_proj.b, bt.txt_rows);
    GENIE_LAUNCHED("fill_rows");
  }
  launch_text_embed_pe(XT, SEQ, d_txt_pos, m.text_emb, m.text_alpha, m.div_term, bt.txt_rows, s);
  keep_tensor(m, "x", XT, (long long)bt.txt_rows * D);

  // ---- K3: xy = [x || Emb_audio[prompts] + alpha*PE(1..Ly)]  (first_stage#[5-25])
  {
    std::vector<const int*> pp(B);
    for (int b = 0; b < B; ++b) pp[b] = prompts[b]->prompts;
    GENIE_CUDA(cudaMemcpyAsync(PPTR, pp.data(), B * sizeof(int*), cudaMemcpyHostToDevice, s));
    GENIE_CUDA(cudaMemsetAsync(HIST, 0, (size_t)B * bt.hist_ld * sizeof(int), s));
    init_hist_kernel<<<B, 128, 0, s>>>(HIST, bt.hist_ld, PPTR, d_ly, B);
    GENIE_LAUNCHED("init_hist");
    for (int b = 0; b < B; ++b) {
      GENIE_CUDA(cudaMemcpyAsync(X + (long long)bt.row_off[b] * D, XT + (long long)bt.txt_off[b] * D,
                                 (size_t)bt.Lx[b] * D * 4, cudaMemcpyDeviceToDevice, s));
      float* dst = X + (long long)(bt.row_off[b] + bt.Lx[b]) * D;
      launch_audio_embed_pe(dst, prompts[b]->prompts, d_aud_pos + bt.row_off[b] + bt.Lx[b], m.audio_emb,
                            m.audio_alpha, m.div_term, bt.Ly[b], s);
    }
  }

  // ---- K4: 24 prefill layers over all rows of all utterances
  const float scale = 1.0f / std::sqrt(32.0f);
  float* Hcur = X;
  for (int l = 0; l < NL; ++l) {
    const T2SLayer& L = m.layers[l];
    run_linear(m, L.qkv, Hcur, D, QKV, 3 * D, R);
    launch_kv_scatter(QKV, 3 * D, KV, utt_stride, l * layer_stride, v_off, bt.cap, d_row_off, d_zero, d_row2utt, R,
                      nullptr, s);
    Attn a;
    a.q = QKV; a.ldq = 3 * D; a.k = QKV + D; a.ldk = 3 * D; a.v = QKV + 2 * D; a.ldv = 3 * D;
    a.o = ATT; a.ldo = D; a.q_off = d_row_off; a.kv_off = d_row_off; a.B = B; a.H = H; a.d = 32; a.max_q = bt.maxS;
    a.scale = scale; a.mask_mode = 1; a.lx = d_lx;
    launch_attention(a, s);
    run_linear(m, L.out, ATT, D, TMP, D, R, ACT_NONE, Hcur, D);
    launch_layernorm(TMP, nullptr, L.ln1_g, L.ln1_b, H1, R, D, s);
    if (ffn16) {
      // FFN1 writes relu(.) straight as the fp16 hi / lo operand rows FFN2 needs (the values its loader would
      // have derived from the fp32 tensor, so nothing changes numerically; same bytes, no conversion pass in
      // FFN2, whose 2 N tiles used to convert every row twice)
      Half2Part ffh{FF16, FF16 + (size_t)R * 4 * D};
      run_linear(m, L.ff1, H1, D, FF, 4 * D, R, ACT_RELU, nullptr, 0, 0, 1, 0, nullptr, &ffh);
      run_linear(m, L.ff2, FF, 4 * D, TMP, D, R, ACT_NONE, H1, D, 0, 1, 0, &ffh, nullptr);
    } else {
      run_linear(m, L.ff1, H1, D, FF, 4 * D, R, ACT_RELU);
      run_linear(m, L.ff2, FF, 4 * D, TMP, D, R, ACT_NONE, H1, D);
    }
    launch_layernorm(TMP, nullptr, L.ln2_g, L.ln2_b, X, R, D, s);
    Hcur = X;
    if (l == 0 && m.keep) {
      keep_tensor(m, "qkv0", QKV, (long long)R * 3 * D);
      keep_tensor(m, "h0", X, (long long)R * D);
    }
  }
  // logits for the last row of each utterance (first_stage#[1785-1788])
  {
    std::vector<long long> li(B);
    for (int b = 0; b < B; ++b) li[b] = bt.row_off[b + 1] - 1;
    GENIE_CUDA(cudaMemcpyAsync(LASTIDX, li.data(), B * sizeof(long long), cudaMemcpyHostToDevice, s));
    launch_gather_rows(LAST, D, X, D, LASTIDX, B, 1, s);
    run_linear(m, m.predict, LAST, D, LOGITS, V, B);
  }
  S.LOGITS = LOGITS; S.HIST = HIST; S.Y64 = Y64;
  S.d_histlen = d_histlen; S.d_kvlen = d_kvlen; S.d_active = d_active; S.d_stop = d_stop;
  m.logits_host.clear();
  record_step_logits(m, S);
  {
    SamplerArgs a{};
    a.logits = LOGITS; a.ld = V; a.hist = HIST; a.hist_ld = bt.hist_ld; a.hist_len = d_histlen; a.kv_len = d_kvlen;
    a.active = d_active; a.stop_step = d_stop; a.B = B; a.top_k = cfg.top_k; a.temperature = cfg.temperature;
    a.penalty = cfg.penalty; a.greedy = cfg.greedy; a.seed = cfg.seed; a.honour_stop = 0; a.advance_kv = 0;
    a.check_stop = 0;   // the first-stage graph has no stop output
    launch_sampler(a, s);
  }
  GENIE_CUDA(cudaEventRecord(ev1, s));

  // ---- decode-step buffers and graph (Inference.py:95-106 runs in t2s_decode_steps)
  StepBufs& w = S.w;
  w = StepBufs{};
  w.h = ws.get<float>("t2s.step.h", (size_t)B * D); w.qkv = ws.get<float>("t2s.step.qkv", (size_t)B * 3 * D);
  w.att = ws.get<float>("t2s.step.att", (size_t)B * D); w.tmp = ws.get<float>("t2s.step.tmp", (size_t)B * D);
  w.h1 = ws.get<float>("t2s.step.h1", (size_t)B * D); w.ff = ws.get<float>("t2s.step.ff", (size_t)B * 4 * D);
  w.part = ws.get<float>("t2s.step.part", (size_t)8 * B * D);
  w.part2 = ws.get<float>("t2s.step.part2", (size_t)8 * B * D);
  w.ppart = ws.get<float>("t2s.step.ppart", (size_t)B * 16 * 8 * 36);
  w.part_stride = (long long)B * D;
  w.logits = LOGITS; w.hist = HIST; w.hist_len = d_histlen; w.kv_len = d_kvlen; w.active = d_active;
  w.stop_step = d_stop; w.kv = KV; w.utt_stride = utt_stride; w.layer_stride = layer_stride; w.v_off = v_off;
  w.cap = bt.cap; w.hist_ld = bt.hist_ld;

  if (m.persistent_step && B <= m.persistent_step) ensure_persistent_step(m);
  // the graph bakes pointers and scalar args; re-capture when any of them changes
  const int flags = (cfg.greedy ? 1 : 0) | (cfg.fixed_steps > 0 ? 2 : 0) | (m.use_tc ? 4 : 0) | (cfg.top_k << 4) |
                    (m.tc_min_rows << 16);
  const bool can_graph = m.use_graph && !m.record_logits && !g_sync_debug;
  S.can_graph = can_graph;
  if (can_graph) {
    bool stale = !m.step_graph || m.step_graph_B != B || m.step_graph_gen != ws.generation ||
                 m.step_graph_cap != bt.cap || m.step_graph_flags != flags || m.step_graph_hist_ld != bt.hist_ld;
    // seed / temperature / penalty are baked too: fold them into staleness via a cheap hash
    if (m.step_graph_seed != cfg.seed || m.step_graph_temp != cfg.temperature || m.step_graph_pen != cfg.penalty)
      stale = true;
    if (stale) {
      if (m.step_graph) { cudaGraphExecDestroy(m.step_graph); m.step_graph = nullptr; }
      cudaGraph_t g = nullptr;
      unsigned long long launches_before = g_launches;
      GENIE_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      try {
        int nb = (B >= m.decode_split_min && m.stream2) ? m.decode_branches : 1;
        if (nb > 1 && (B + nb - 1) / nb > 128) nb = (B + 127) / 128;   // keep every branch on the <= 128-row GEMM
        if (nb > 4) nb = 4;
        while (nb > 1 && B / nb < 16) --nb;              // every branch stays on the tensor-core path
        if (nb > 1) {
          // independent utterance ranges as parallel graph branches: every decode kernel is latency-bound
          // (or, the attention, bandwidth-bound with few resources), so the branches overlap
          cudaStream_t bs[4] = {s, m.stream2, m.stream3, m.stream4};
          cudaEvent_t bj[4] = {nullptr, m.ev_join, m.ev_join3, m.ev_join4};
          GENIE_CUDA(cudaEventRecord(m.ev_fork, s));
          cudaStream_t keep = m.stream;
          try {
            for (int k = 0; k < nb; ++k) {
              const int b0 = (int)((long long)B * k / nb), b1 = (int)((long long)B * (k + 1) / nb);
              StepBufs w2 = w;
              w2.h += (size_t)b0 * D; w2.qkv += (size_t)b0 * 3 * D; w2.att += (size_t)b0 * D; w2.tmp += (size_t)b0 * D;
              w2.h1 += (size_t)b0 * D; w2.ff += (size_t)b0 * 4 * D; w2.logits += (size_t)b0 * V;
              // private slice of the partial-sum buffers ([split][rows][N] views with N changing per use must
              // not overlap between branches that run in different phases): 8 * rows * D floats per branch
              w2.part += (size_t)8 * b0 * D; w2.part2 += (size_t)8 * b0 * D; w2.part_stride = (long long)(b1 - b0) * D;
              w2.hist += (size_t)b0 * w.hist_ld;
              w2.hist_len += b0; w2.kv_len += b0; w2.active += b0; w2.stop_step += b0;
              w2.kv += (size_t)b0 * w.utt_stride; w2.utt_base = b0;
              if (k > 0) GENIE_CUDA(cudaStreamWaitEvent(bs[k], m.ev_fork, 0));
              m.stream = bs[k];
              decode_step(m, w2, b1 - b0, cfg);
              if (k > 0) GENIE_CUDA(cudaEventRecord(bj[k], bs[k]));
            }
          } catch (...) { m.stream = keep; throw; }
          m.stream = keep;
          for (int k = 1; k < nb; ++k) GENIE_CUDA(cudaStreamWaitEvent(s, bj[k], 0));
        } else {
          decode_step(m, w, B, cfg);
        }
      } catch (...) {
        cudaStreamEndCapture(s, &g);
        if (g) cudaGraphDestroy(g);
        throw;
      }
      GENIE_CUDA(cudaStreamEndCapture(s, &g));
      m.step_graph_launches = g_launches - launches_before;
      g_launches = launches_before;            // capture does not launch
      GENIE_CUDA(cudaGraphInstantiate(&m.step_graph, g, 0));
      cudaGraphDestroy(g);
      m.step_graph_B = B; m.step_graph_gen = ws.generation; m.step_graph_cap = bt.cap; m.step_graph_flags = flags; m.step_graph_hist_ld = bt.hist_ld;
      m.step_graph_seed = cfg.seed; m.step_graph_temp = cfg.temperature; m.step_graph_pen = cfg.penalty;
    }
  }
  m.t2s_session = sess_ptr;
}

// up to n_steps more decode steps for the batch in flight (stage graph + sampler per step, Inference.py:95-106);
// returns GENIE_CANCELLED (2) when the host flag fired, 0 otherwise
int t2s_decode_steps(Model& m, int n_steps, const volatile int* cancel, int* n_active_out, int* steps_done_out) {
  T2SSession& S = session_of(m);
  cudaStream_t s = m.stream;
  GENIE_CUDA(cudaSetDevice(m.device));
  const int B = S.B;
  const SamplingCfg& cfg = S.cfg;
  int rc = 0;
  std::vector<int> h_act(B, 1);
  const int poll_every = 8;
  cudaEvent_t ea, eb;
  GENIE_CUDA(cudaEventCreate(&ea)); GENIE_CUDA(cudaEventCreate(&eb));
  S.decode_spans.emplace_back(ea, eb);
  GENIE_CUDA(cudaEventRecord(ea, s));
  for (int k = 0; k < n_steps && S.steps_done < S.max_steps && !S.all_stopped; ++k) {
    if (cancel && *cancel) { rc = 2 /* GENIE_CANCELLED */; break; }
    if (S.can_graph) {
      GENIE_CUDA(cudaGraphLaunch(m.step_graph, s));
      g_launches += m.step_graph_launches;
    } else {
      decode_step(m, S.w, B, cfg);
    }
    ++S.steps_done;
    record_step_logits(m, S);
    if (cfg.fixed_steps <= 0 && (S.steps_done % poll_every == 0 || m.record_logits)) {
      GENIE_CUDA(cudaMemcpyAsync(h_act.data(), S.d_active, B * sizeof(int), cudaMemcpyDeviceToHost, s));
      GENIE_CUDA(cudaStreamSynchronize(s));
      bool any = false;
      for (int b = 0; b < B; ++b) any = any || h_act[b];
      if (!any) S.all_stopped = true;
    }
  }
  GENIE_CUDA(cudaEventRecord(eb, s));
  if (n_active_out) {
    int n = 0;
    if (cfg.fixed_steps > 0) n = S.steps_done < S.max_steps ? B : 0;
    else if (!S.all_stopped && S.steps_done < S.max_steps) {
      GENIE_CUDA(cudaMemcpyAsync(h_act.data(), S.d_active, B * sizeof(int), cudaMemcpyDeviceToHost, s));
      GENIE_CUDA(cudaStreamSynchronize(s));
      for (int b = 0; b < B; ++b) n += h_act[b] ? 1 : 0;
      if (n == 0) S.all_stopped = true;
    }
    *n_active_out = n;
  }
  if (steps_done_out) *steps_done_out = S.steps_done;
  return rc;
}

// tokens generated so far (prompt tokens + generated, per utterance) and the reference's loop index
void t2s_read(Model& m, int io_dev, int64_t* y_out, int y_ld, int* y_len_out, int* idx_out) {
  T2SSession& S = session_of(m);
  cudaStream_t s = m.stream;
  GENIE_CUDA(cudaSetDevice(m.device));
  const int B = S.B;
  const Batch& bt = S.bt;
  const SamplingCfg& cfg = S.cfg;
  const StepBufs& w = S.w;
  const int steps_done = S.steps_done;
  GENIE_CHECK(y_ld >= bt.hist_ld || y_out == nullptr, "y_ld too small: need >= " + std::to_string(bt.hist_ld));

  // ---- results
  std::vector<int> h_len(B), h_stopv(B), h_kvlen_final(B);
  GENIE_CUDA(cudaMemcpyAsync(h_len.data(), S.d_histlen, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  GENIE_CUDA(cudaMemcpyAsync(h_kvlen_final.data(), S.d_kvlen, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  GENIE_CUDA(cudaMemcpyAsync(h_stopv.data(), S.d_stop, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (y_out) {
    hist_to_i64_kernel<<<B, 256, 0, s>>>(S.HIST, bt.hist_ld, S.d_histlen, S.Y64, bt.hist_ld, B);
    GENIE_LAUNCHED("hist_to_i64");
    GENIE_CUDA(cudaMemcpy2DAsync(y_out, (size_t)y_ld * 8, S.Y64, (size_t)bt.hist_ld * 8, (size_t)bt.hist_ld * 8, B,
                                 io_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  GENIE_CUDA(cudaStreamSynchronize(s));
  check_tc_error(m);
  if (m.step_sync) {
    unsigned flag[2] = {0, 0};
    GENIE_CUDA(cudaMemcpy(flag, m.step_sync, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag[1]) {
      cudaMemset(m.step_sync, 0, sizeof(flag));
      GENIE_CHECK(false, "persistent decode step: device-wide barrier timed out");
    }
  }
  for (int b = 0; b < B; ++b) {
    // reference loop variable at exit: index of the step whose stop flag fired, else last index
    int idx;
    if (cfg.fixed_steps > 0 || h_stopv[b] < 0) idx = steps_done - 1;
    else idx = h_stopv[b] - (bt.Ly[b] + 1);     // stop_step stores the history length before that step's token
    if (y_len_out) y_len_out[b] = h_len[b];
    if (idx_out) idx_out[b] = idx < 0 ? 0 : idx;
  }
  // measurement aid (bench.py roofline): the dominant decode kernel sits inside a CUDA graph, so it is
  // replayed here on the final cache state, all layers back to back (2.8 GB of KV >> L2), between events
  if (m.time_attention > 0 && B > m.skinny_max_rows && B <= 128) {
    cudaEvent_t ea, eb;
    cudaEventCreate(&ea); cudaEventCreate(&eb);
    const float scale = 1.0f / std::sqrt(32.0f);
    GENIE_CUDA(cudaEventRecord(ea, s));
    for (int r = 0; r < m.time_attention; ++r)
      for (int l = 0; l < 24; ++l)
        launch_decode_attention_fused(w.part, 2, w.part_stride * 3, m.layers[l].qkv.b, w.att, w.kv, w.utt_stride,
                                      l * w.layer_stride, w.v_off, w.kv_len, nullptr, B, w.cap, scale, s);
    GENIE_CUDA(cudaEventRecord(eb, s));
    GENIE_CUDA(cudaEventSynchronize(eb));
    float ms = 0;
    cudaEventElapsedTime(&ms, ea, eb);
    double kv_tokens = 0;
    for (int b = 0; b < B; ++b) kv_tokens += h_kvlen_final[b];
    m.timing[8] = 1000.f * ms / (float)(m.time_attention * 24);
    m.timing[9] = (float)(kv_tokens * 2 * 512 * 4 / 1e6);
    cudaEventDestroy(ea); cudaEventDestroy(eb);
  }
  float t01 = 0, t12 = 0;
  cudaEventElapsedTime(&t01, S.ev0, S.ev1);
  for (auto& e : S.decode_spans) {
    float t = 0;
    cudaEventElapsedTime(&t, e.first, e.second);
    t12 += t;
  }
  m.timing[0] = t01; m.timing[1] = t12; m.timing[2] = t01 + t12; m.timing[3] = (float)steps_done;
}

// Inference.py:63-106 in one call: prefill, decode to the stop condition / budget, read the tokens
int t2s_generate(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len,
                 const float* text_bert, const SamplingCfg& cfg, const volatile int* cancel, int io_dev,
                 int64_t* y_out, int y_ld, int* y_len_out, int* idx_out) {
  t2s_prefill(m, prompts, B, text_seq, text_len, text_bert, cfg, io_dev);
  const int max_steps = cfg.fixed_steps > 0 ? cfg.fixed_steps : cfg.max_steps;
  const int rc = t2s_decode_steps(m, max_steps, cancel, nullptr, nullptr);
  t2s_read(m, io_dev, y_out, y_ld, y_len_out, idx_out);
  return rc;
}

}  // namespace genie
