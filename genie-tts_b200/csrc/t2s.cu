// T2S GPT stage: text encode (t2s_encoder#[49-83]) -> prefill (t2s_first_stage_decoder) -> decode loop
// (t2s_stage_decoder x <=500, reference src/genie_tts/Core/Inference.py:76-106) for independent utterances.
//
// The stage is organised as a POOL OF DECODE SLOTS.  A slot owns a KV slab (fp32, head-major, appended in place),
// a token-history row and its decode parameters (SlotParams, in device memory).  Utterances are ADMITTED into free
// slots (the admission runs their prefill as one ragged batch and writes K/V straight into the slot slabs), every
// decode step advances all active slots at once (one CUDA-graph replay over the first `rows` slots; stopped / free
// slots exit early in every kernel), finished slots are read and released.  The batch API of the reference-facing
// path (genie_t2s_prefill / decode_steps / read / generate) is the special case "pool sized for this batch, all
// utterances admitted at once"; the continuous-batching server admits and releases slots while others decode.
#include "model.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace genie {
namespace {

constexpr int D = 512, NL = 24, H = 16, V = 1025;

__global__ void hist_to_i64_kernel(const int* hist, int hist_ld, const int* hist_len, long long* y, int y_ld, int B) {
  int b = blockIdx.x;
  int n = hist_len[b];
  for (int i = threadIdx.x; i < y_ld; i += blockDim.x)
    y[(long long)b * y_ld + i] = i < n ? hist[(long long)b * hist_ld + i] : 0;
}

struct StepBufs {
  unsigned long long key = 0;   // mix of the buffer addresses below (what a captured step bakes)
  float *h, *qkv, *att, *tmp, *h1, *ff, *logits, *part, *part2, *ppart;
  int *hist, *hist_len, *kv_len, *active, *stop_step;
  const SlotParams* params;
  void* kv; int kv_f16;   // cache of floats, or of halves (option kv_fp16); strides below are in elements
  long long utt_stride, layer_stride, v_off; int cap, hist_ld; long long part_stride;
};

struct SlotHost { int in_use = 0, Ly = 0, S = 0, max_steps = 0, honour = 1; };

// captured decode steps, keyed by everything the graph bakes: row count, slab geometry, buffer addresses
// (workspace generation) and the path-selection options
struct StepGraph {
  int rows = 0, cap = 0, hist_ld = 0; unsigned long long gen = 0, opt = 0;
  cudaGraphExec_t exec = nullptr; unsigned long long launches = 0; unsigned long long last_use = 0;
};
struct GraphCache {
  std::vector<StepGraph> e; unsigned long long tick = 0;
  ~GraphCache() { for (auto& g : e) if (g.exec) cudaGraphExecDestroy(g.exec); }
};

// The slot pool of one model handle.  Buffers live in the handle's workspace (grow-only, named), so a handle has
// one pool at a time: creating a new one (any batch-API prefill does) drops the previous one.
struct T2SSession {
  int n_slots = 0, cap = 0, hist_ld = 0;
  bool pool_mode = false;
  float* LOGITS = nullptr; float* LOGITS_PRE = nullptr; int* HIST = nullptr; long long* Y64 = nullptr;
  int *d_histlen = nullptr, *d_kvlen = nullptr, *d_active = nullptr, *d_stop = nullptr;
  SlotParams* d_params = nullptr;
  StepBufs w{};
  std::vector<SlotHost> slots;
  // batch API bookkeeping
  int B = 0, max_steps = 0, steps_done = 0, fixed = 0;
  bool all_stopped = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;                       // last admission (prefill)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> decode_spans;  // one pair per decode_steps call
  ~T2SSession() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto& e : decode_spans) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  }
};

T2SSession& session_of(Model& m) {
  GENIE_CHECK(m.t2s_session != nullptr, "no T2S batch in flight: call genie_t2s_prefill / genie_t2s_pool_create first");
  return *static_cast<T2SSession*>(m.t2s_session.get());
}
GraphCache& graphs_of(Model& m) {
  if (!m.t2s_graphs) m.t2s_graphs = std::make_shared<GraphCache>();
  return *static_cast<GraphCache*>(m.t2s_graphs.get());
}

void record_logits_rows(Model& m, const float* dev, int rows) {
  if (!m.record_logits) return;
  std::vector<float>& rec = m.logits_host;
  const size_t o = rec.size();
  rec.resize(o + (size_t)rows * V);
  GENIE_CUDA(cudaMemcpyAsync(rec.data() + o, dev, (size_t)rows * V * 4, cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaStreamSynchronize(m.stream));
}

// barrier words of the persistent step (per handle; allocated outside any stream capture) and the residency check
void ensure_persistent_step(Model& m) {
  if (!m.step_sync) {
    GENIE_CUDA(cudaMalloc(&m.step_sync, 2 * sizeof(unsigned)));
    GENIE_CUDA(cudaMemset(m.step_sync, 0, 2 * sizeof(unsigned)));
    m.ctx_owned.push_back(m.step_sync);
  }
  if (m.persistent_ok < 0) {
    // one CTA per SM must be co-resident for the device-wide barrier; the launch itself is cooperative
    // (t2s_persistent.cu), this check only decides whether the path is offered at all on this device
    m.persistent_ok = persistent_step_fits(8, m.num_sms) ? 1 : 0;
  }
}

// one decode step for every active utterance (stage#[12-1821])
void decode_step(Model& m, const StepBufs& w, int B) {
  cudaStream_t s = m.stream;
  struct PdlScope { int prev; explicit PdlScope(int on) : prev(g_pdl_now) { g_pdl_now = on; } ~PdlScope() { g_pdl_now = prev; } };
  const PdlScope pdl_scope(B > m.skinny_max_rows ? 1 : 0);
  launch_decode_embed(w.h, w.hist, w.hist_ld, w.hist_len, w.active, m.audio_emb, m.audio_alpha, m.div_term, B, s);
  const float scale = 1.0f / std::sqrt(32.0f);
  // measured (90 steps): batch 1 / 2 / 4 / 8 = 37 / 43 / 53 / 77 ms persistent vs 58 / - / 60 / 70 ms kernel chain
  const bool persistent = m.persistent_step && B <= m.persistent_step && B <= m.skinny_max_rows && m.layers[0].qkv.w_f16 &&
                          m.predict.w_f16 && w.ppart != nullptr && m.step_layers_dev != nullptr && m.step_sync != nullptr &&
                          m.persistent_ok == 1;
  if (persistent) {
    // batch <= 8: all 24 layers + logits in one resident kernel (t2s_persistent.cu)
    PersistentStep a;
    a.layers = reinterpret_cast<const StepLayerPtrs*>(m.step_layers_dev); a.n_layers = NL;
    a.wpredict = reinterpret_cast<const __half*>(m.predict.w); a.bpredict = m.predict.b; a.vocab = V;
    a.h = w.h; a.qkv = w.qkv; a.part = w.ppart; a.lnin = w.tmp; a.lnin2 = w.part2; a.h1 = w.h1; a.ff = w.ff;
    a.logits = w.logits; a.ld_logits = V;
    a.kv = w.kv; a.kv_f16 = w.kv_f16; a.utt_stride = w.utt_stride; a.layer_stride = w.layer_stride; a.v_off = w.v_off; a.cap = w.cap;
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
      launch_decode_attention_fused(w.part + (size_t)r0 * 3 * D, 2, ps * 3, L.qkv.b, w.att, w.kv, w.kv_f16, w.utt_stride,
                                    l * w.layer_stride, w.v_off, w.kv_len, w.active, B, w.cap, scale, s);
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
    launch_decode_attention_fused(w.qkv, 1, 0, nullptr, w.att, w.kv, w.kv_f16, w.utt_stride, l * w.layer_stride, w.v_off,
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
  a.kv_len = w.kv_len; a.active = w.active; a.stop_step = w.stop_step; a.params = w.params; a.slot_map = nullptr;
  a.B = B; a.advance_kv = 1; a.check_stop = 1;
  launch_sampler(a, s);
}


SlotParams slot_params(const SamplingCfg& c, int Ly, int utt) {
  SlotParams p{};
  p.top_k = c.top_k; p.greedy = c.greedy; p.honour_stop = c.fixed_steps > 0 ? 0 : 1;
  const int steps = c.fixed_steps > 0 ? c.fixed_steps : c.max_steps;
  p.hist_max = Ly + 1 + steps;
  p.temperature = c.temperature; p.penalty = c.penalty; p.top_p = c.top_p; p.utt = utt; p.seed = c.seed;
  return p;
}

// (re)build the pool of a handle: per-slot state, KV slabs and decode-step buffers for n_slots slots of kv_cap tokens
T2SSession& pool_build(Model& m, int n_slots, int kv_cap, int hist_ld, bool pool_mode) {
  GENIE_CHECK(m.finalized, "model not finalized");
  GENIE_CHECK(n_slots > 0 && kv_cap > 0 && hist_ld > 0, "bad pool geometry");
  GENIE_CUDA(cudaSetDevice(m.device));
  m.t2s_session.reset();
  std::shared_ptr<T2SSession> sp = std::make_shared<T2SSession>();
  T2SSession& S = *sp;
  S.n_slots = n_slots; S.cap = (kv_cap + 15) / 16 * 16; S.hist_ld = hist_ld; S.pool_mode = pool_mode;
  S.slots.assign(n_slots, SlotHost{});
  Workspace& ws = m.ws;
  const int B = n_slots;
  S.LOGITS = ws.get<float>("t2s.logits", (size_t)B * V);
  S.HIST = ws.get<int>("t2s.hist", (size_t)B * hist_ld);
  S.Y64 = ws.get<long long>("t2s.y64", (size_t)B * hist_ld);
  S.d_histlen = ws.get<int>("t2s.slot.histlen", B); S.d_kvlen = ws.get<int>("t2s.slot.kvlen", B);
  S.d_active = ws.get<int>("t2s.slot.active", B); S.d_stop = ws.get<int>("t2s.slot.stop", B);
  S.d_params = ws.get<SlotParams>("t2s.slot.params", B);
  // KV cache [slot][layer][K|V][H][cap][32] fp32
  const long long head_sz = (long long)S.cap * 32, v_off = H * head_sz, layer_stride = 2 * v_off,
                  utt_stride = NL * layer_stride;
  StepBufs& w = S.w;
  w = StepBufs{};
  // fp16 rows by default for pools that decode on the batched kernels; option kv_fp16 = 0 keeps the reference's
  // fp32 rows (measured greedy-token agreement of both against the reference graphs)
  // Batches that decode on the persistent step kernel (<= persistent_step rows, the first-audio path) keep fp32 rows:
  // that kernel is latency-bound, cache bytes do not matter there.
  w.kv_f16 = (m.kv_fp16 && (pool_mode || n_slots > m.persistent_step)) ? 1 : 0;
  w.kv = w.kv_f16 ? static_cast<void*>(ws.get<__half>("t2s.kv", (size_t)B * utt_stride))
                  : static_cast<void*>(ws.get<float>("t2s.kv", (size_t)B * utt_stride));
  w.h = ws.get<float>("t2s.step.h", (size_t)B * D); w.qkv = ws.get<float>("t2s.step.qkv", (size_t)B * 3 * D);
  w.att = ws.get<float>("t2s.step.att", (size_t)B * D); w.tmp = ws.get<float>("t2s.step.tmp", (size_t)B * D);
  w.h1 = ws.get<float>("t2s.step.h1", (size_t)B * D); w.ff = ws.get<float>("t2s.step.ff", (size_t)B * 4 * D);
  w.part = ws.get<float>("t2s.step.part", (size_t)8 * B * D);
  w.part2 = ws.get<float>("t2s.step.part2", (size_t)8 * B * D);
  w.ppart = ws.get<float>("t2s.step.ppart", (size_t)std::max(B, 8) * 16 * 8 * 36);
  w.part_stride = (long long)B * D;
  w.logits = S.LOGITS; w.hist = S.HIST; w.hist_len = S.d_histlen; w.kv_len = S.d_kvlen; w.active = S.d_active;
  w.stop_step = S.d_stop; w.params = S.d_params;
  w.utt_stride = utt_stride; w.layer_stride = layer_stride; w.v_off = v_off; w.cap = S.cap; w.hist_ld = hist_ld;
  {
    const void* ptrs[] = {w.h, w.qkv, w.att, w.tmp, w.h1, w.ff, w.logits, w.part, w.part2, w.ppart, w.hist, w.hist_len,
                          w.kv_len, w.active, w.stop_step, w.params, w.kv};
    unsigned long long k = 1469598103934665603ull ^ (unsigned long long)w.kv_f16;
    for (const void* q : ptrs) { k ^= (unsigned long long)reinterpret_cast<uintptr_t>(q); k *= 1099511628211ull; }
    w.key = k;
  }
  cudaStream_t s = m.stream;
  // free slots must read as inactive; rows of free slots flow through the GEMMs, so start them finite
  GENIE_CUDA(cudaMemsetAsync(S.d_active, 0, B * sizeof(int), s));
  GENIE_CUDA(cudaMemsetAsync(S.d_histlen, 0, B * sizeof(int), s));
  GENIE_CUDA(cudaMemsetAsync(S.d_kvlen, 0, B * sizeof(int), s));
  if (pool_mode) {
    GENIE_CUDA(cudaMemsetAsync(w.h, 0, (size_t)B * D * 4, s));
    GENIE_CUDA(cudaMemsetAsync(w.att, 0, (size_t)B * D * 4, s));
    GENIE_CUDA(cudaMemsetAsync(w.h1, 0, (size_t)B * D * 4, s));
    GENIE_CUDA(cudaMemsetAsync(w.part, 0, (size_t)8 * B * D * 4, s));
    GENIE_CUDA(cudaMemsetAsync(w.part2, 0, (size_t)8 * B * D * 4, s));
  }
  if (m.persistent_step > 0) ensure_persistent_step(m);
  m.t2s_session = sp;
  return S;
}

// Admission = encoder + first-stage graph (Inference.py:76-93) for n new utterances as one ragged batch: embeddings,
// 24 prefill layers writing K/V into the slots' slabs, first sampled token appended to the slots' histories.
void admit(Model& m, T2SSession& S, int n, const int* slots_in, Prompt* const* prompts, const int64_t* text_seq,
           const int* text_len, const float* text_bert, const SamplingCfg* cfgs, int n_cfg, int io_dev) {
  GENIE_CHECK(n > 0, "empty batch");
  GENIE_CUDA(cudaSetDevice(m.device));
  const BulkStreamScope bulk(m);             // prefill is throughput-bound: the handle's low-priority stream; the
  cudaStream_t s = m.stream;                 // decode steps (main stream) are ordered after it when the scope ends
  // ---- geometry
  std::vector<int> Lr(n), Lt(n), Ly(n), Lx(n), Sx(n), slot(n), row_off(n + 1, 0), txt_off(n + 1, 0), txt_in_off(n + 1, 0);
  bool any_bert = text_bert != nullptr;
  int maxS = 0;
  for (int b = 0; b < n; ++b) {
    Prompt* p = prompts[b];
    const SamplingCfg& c = cfgs[n_cfg == 1 ? 0 : b];
    GENIE_CHECK(p && p->model_uid == m.owner->uid, "prompt does not belong to this model");
    GENIE_CHECK(text_len[b] > 0, "empty text_seq");
    const int steps = c.fixed_steps > 0 ? c.fixed_steps : c.max_steps;
    slot[b] = slots_in ? slots_in[b] : b;
    GENIE_CHECK(slot[b] >= 0 && slot[b] < S.n_slots, "slot index out of range");
    GENIE_CHECK(!S.slots[slot[b]].in_use, "slot " + std::to_string(slot[b]) + " is still in use");
    Lr[b] = p->Lr; Lt[b] = text_len[b]; Ly[b] = p->Ly; Lx[b] = Lr[b] + Lt[b]; Sx[b] = Lx[b] + Ly[b];
    GENIE_CHECK(Sx[b] + steps + 1 <= S.cap, "utterance needs " + std::to_string(Sx[b] + steps + 1) +
                                                " KV rows, the pool's slots hold " + std::to_string(S.cap));
    GENIE_CHECK(Ly[b] + steps + 2 <= S.hist_ld, "utterance needs a longer token history than the pool was built for");
    row_off[b + 1] = row_off[b] + Sx[b]; txt_off[b + 1] = txt_off[b] + Lx[b]; txt_in_off[b + 1] = txt_in_off[b] + Lt[b];
    maxS = std::max(maxS, Sx[b]);
    any_bert = any_bert || p->has_bert;
  }
  for (int b = 0; b < n; ++b)
    for (int c = 0; c < b; ++c) GENIE_CHECK(slot[b] != slot[c], "slot listed twice in one admission");
  const int R = row_off[n], txt_rows = txt_off[n], n_text_in = txt_in_off[n];
  if (!io_dev)
    for (int i = 0; i < n_text_in; ++i)
      GENIE_CHECK(text_seq[i] >= 0 && text_seq[i] < m.text_vocab, "text_seq id out of range (phoneme table has " +
                                                                      std::to_string(m.text_vocab) + " rows)");

  Workspace& ws = m.ws;
  float* X = ws.get<float>("t2s.x", (size_t)R * D);
  float* QKV = ws.get<float>("t2s.qkv", (size_t)R * 3 * D);
  float* ATT = ws.get<float>("t2s.att", (size_t)R * D);
  float* TMP = ws.get<float>("t2s.tmp", (size_t)R * D);
  float* H1 = ws.get<float>("t2s.h1", (size_t)R * D);
  float* FF = ws.get<float>("t2s.ff", (size_t)R * 4 * D);
  // the same bytes viewed as two fp16 matrices [R, 2048] (hi | lo) when the FFN pair hands over in fp16
  __half* FF16 = reinterpret_cast<__half*>(FF);
  static const bool ffn16_env = [] { const char* e = getenv("GENIE_FFN16"); return !(e && e[0] == '0'); }();
  const bool ffn16 = ffn16_env && !m.prefill_single && m.use_tc && R >= m.tc_min_rows && R > m.skinny_max_rows && m.layers[0].ff1.tc.hi &&
                     m.layers[0].ff2.tc.hi && !m.layers[0].ff2.tc.lo;
  float* XT = any_bert ? ws.get<float>("t2s.xtext", (size_t)txt_rows * D) : nullptr;
  float* BERT = any_bert ? ws.get<float>("t2s.bert", (size_t)txt_rows * 1024) : nullptr;
  long long* SEQ_IN = io_dev ? nullptr : ws.get<long long>("t2s.seq_in", n_text_in);
  float* BERT_IN = (text_bert && !io_dev) ? ws.get<float>("t2s.bert_in", (size_t)n_text_in * 1024) : nullptr;
  float* LAST = ws.get<float>("t2s.last", (size_t)n * D);
  float* LOGITS_PRE = ws.get<float>("t2s.logits_pre", (size_t)n * V);
  long long* LASTIDX = ws.get<long long>("t2s.lastidx", n);
  // one int block + one pointer block + the slot-init records per admission
  //   ints: row_off[n+1] | txt_off[n+1] | txt_in_off[n+1] | lr[n] | lx[n] | slot[n] | zero[n] | row2utt[R]
  const size_t n_int = (size_t)3 * (n + 1) + 4 * n + R;
  int* IMETA = ws.get<int>("t2s.imeta", n_int);
  const void** PMETA = ws.get<const void*>("t2s.pmeta", (size_t)3 * n);
  SlotInit* SINIT = ws.get<SlotInit>("t2s.sinit", n);

  std::vector<int> meta(n_int, 0);
  int* h_row_off = meta.data(); int* h_txt_off = h_row_off + (n + 1); int* h_txt_in = h_txt_off + (n + 1);
  int* h_lr = h_txt_in + (n + 1); int* h_lx = h_lr + n; int* h_slot = h_lx + n; int* h_zero = h_slot + n;
  int* h_row2utt = h_zero + n;
  std::vector<const void*> pmeta((size_t)3 * n);
  std::vector<SlotInit> sinit(n);
  std::vector<long long> li(n);
  for (int b = 0; b < n; ++b) {
    h_row_off[b] = row_off[b]; h_txt_off[b] = txt_off[b]; h_txt_in[b] = txt_in_off[b];
    h_lr[b] = Lr[b]; h_lx[b] = Lx[b]; h_slot[b] = slot[b];
    for (int i = 0; i < Sx[b]; ++i) h_row2utt[row_off[b] + i] = b;
    pmeta[b] = prompts[b]->ref_seq; pmeta[n + b] = prompts[b]->ref_bert; pmeta[2 * n + b] = prompts[b]->prompts;
    const SamplingCfg& c = cfgs[n_cfg == 1 ? 0 : b];
    sinit[b] = SlotInit{slot[b], Sx[b], Ly[b], 0, slot_params(c, Ly[b], b)};
    li[b] = row_off[b + 1] - 1;
  }
  h_row_off[n] = R; h_txt_off[n] = txt_rows; h_txt_in[n] = n_text_in;
  GENIE_CUDA(cudaMemcpyAsync(IMETA, meta.data(), n_int * sizeof(int), cudaMemcpyHostToDevice, s));
  GENIE_CUDA(cudaMemcpyAsync(PMETA, pmeta.data(), pmeta.size() * sizeof(void*), cudaMemcpyHostToDevice, s));
  GENIE_CUDA(cudaMemcpyAsync(SINIT, sinit.data(), sinit.size() * sizeof(SlotInit), cudaMemcpyHostToDevice, s));
  GENIE_CUDA(cudaMemcpyAsync(LASTIDX, li.data(), n * sizeof(long long), cudaMemcpyHostToDevice, s));
  const long long* d_text = reinterpret_cast<const long long*>(text_seq);
  if (!io_dev) {
    GENIE_CUDA(cudaMemcpyAsync(SEQ_IN, text_seq, (size_t)n_text_in * 8, cudaMemcpyHostToDevice, s));
    d_text = SEQ_IN;
  }
  const float* d_text_bert = text_bert;
  if (text_bert && !io_dev) {
    GENIE_CUDA(cudaMemcpyAsync(BERT_IN, text_bert, (size_t)n_text_in * 1024 * 4, cudaMemcpyHostToDevice, s));
    d_text_bert = BERT_IN;
  }
  // the host vectors above are pageable: the copies are staged by the runtime before the call returns
  int* d_row_off = IMETA; int* d_txt_off = d_row_off + (n + 1); int* d_txt_in = d_txt_off + (n + 1);
  int* d_lr = d_txt_in + (n + 1); int* d_lx = d_lr + n; int* d_slot = d_lx + n; int* d_zero = d_slot + n;
  int* d_row2utt = d_zero + n;
  PrefillMeta pm{};
  pm.row2utt = d_row2utt; pm.row_off = d_row_off; pm.txt_off = d_txt_off; pm.txt_in_off = d_txt_in; pm.lr = d_lr; pm.lx = d_lx;
  pm.ref_seq = reinterpret_cast<const long long* const*>(PMETA);
  pm.ref_bert = reinterpret_cast<const float* const*>(PMETA + n);
  pm.prompt_tok = reinterpret_cast<const int* const*>(PMETA + 2 * n);

  if (!S.ev0) { GENIE_CUDA(cudaEventCreate(&S.ev0)); GENIE_CUDA(cudaEventCreate(&S.ev1)); }
  GENIE_CUDA(cudaEventRecord(S.ev0, s));

  // ---- K1 + K3: xy = [Emb_text[ref||text] + bert_proj(bert) + alpha PE(1..Lx)  ||  Emb_audio[prompts] + alpha PE(1..Ly)]
  // (t2s_encoder#[49-83], first_stage#[5-25]); bert features are zeros for ja/en (GetPhonesAndBert.py:60,80), where
  // bert_proj reduces to its bias
  if (any_bert) {
    launch_prefill_bert_gather(BERT, pm, d_text_bert, R, s);
    run_linear(m, m.bert_proj, BERT, 1024, XT, D, txt_rows);
  }
  launch_prefill_embed(X, pm, d_text, XT, m.bert_proj.b, m.text_emb, m.text_alpha, m.text_vocab, m.audio_emb,
                       m.audio_alpha, m.div_term, R, m.tc_err, s);
  launch_slot_init(SINIT, n, pm.prompt_tok, S.HIST, S.hist_ld, S.d_histlen, S.d_kvlen, S.d_active, S.d_stop, S.d_params, s);
  if (m.keep) {   // debug: the encoder graph's output x = the text rows
    float* XK = ws.get<float>("t2s.xkeep", (size_t)txt_rows * D);
    for (int b = 0; b < n; ++b)
      GENIE_CUDA(cudaMemcpyAsync(XK + (size_t)txt_off[b] * D, X + (size_t)row_off[b] * D, (size_t)Lx[b] * D * 4,
                                 cudaMemcpyDeviceToDevice, s));
    keep_tensor(m, "x", XK, (long long)txt_rows * D);
  }

  // ---- K4: 24 prefill layers over all rows of all admitted utterances
  const StepBufs& w = S.w;
  const float scale = 1.0f / std::sqrt(32.0f);
  static const bool prefill_mma = [] { const char* e = getenv("GENIE_PREFILL_MMA"); return !(e && e[0] == '0'); }();
  float* Hcur = X;
  m.lin_single_now = m.prefill_single;
  for (int l = 0; l < NL; ++l) {
    const T2SLayer& L = m.layers[l];
    run_linear(m, L.qkv, Hcur, D, QKV, 3 * D, R);
    launch_kv_scatter(QKV, 3 * D, w.kv, w.kv_f16, w.utt_stride, l * w.layer_stride, w.v_off, S.cap, d_row_off, d_zero,
                      d_row2utt, R, nullptr, s, d_slot);
    Attn a;
    a.q = QKV; a.ldq = 3 * D; a.k = QKV + D; a.ldk = 3 * D; a.v = QKV + 2 * D; a.ldv = 3 * D;
    a.o = ATT; a.ldo = D; a.q_off = d_row_off; a.kv_off = d_row_off; a.B = n; a.H = H; a.d = 32; a.max_q = maxS;
    a.scale = scale; a.mask_mode = 1; a.lx = d_lx;
    a.use_mma = prefill_mma ? 1 : 0;                   // tensor-core kernel, 3-product split precision
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
  // logits for the last row of each utterance (first_stage#[1785-1788]) and the first sampled token (#[1789-1820];
  // the first-stage graph has no stop output)
  m.lin_single_now = 0;
  launch_gather_rows(LAST, D, X, D, LASTIDX, n, 1, s);
  run_linear(m, m.predict, LAST, D, LOGITS_PRE, V, n);
  S.LOGITS_PRE = LOGITS_PRE;
  record_logits_rows(m, LOGITS_PRE, n);
  {
    SamplerArgs a{};
    a.logits = LOGITS_PRE; a.ld = V; a.hist = S.HIST; a.hist_ld = S.hist_ld; a.hist_len = S.d_histlen; a.kv_len = S.d_kvlen;
    a.active = S.d_active; a.stop_step = S.d_stop; a.params = S.d_params; a.slot_map = d_slot; a.B = n;
    a.advance_kv = 0; a.check_stop = 0;
    launch_sampler(a, s);
  }
  GENIE_CUDA(cudaEventRecord(S.ev1, s));
  for (int b = 0; b < n; ++b) {
    const SamplingCfg& c = cfgs[n_cfg == 1 ? 0 : b];
    SlotHost& h = S.slots[slot[b]];
    h.in_use = 1; h.Ly = Ly[b]; h.S = Sx[b]; h.honour = c.fixed_steps > 0 ? 0 : 1;
    h.max_steps = c.fixed_steps > 0 ? c.fixed_steps : c.max_steps;
  }
}

// the captured decode step over the first `rows` slots (captured on first use, then replayed)
StepGraph& step_graph(Model& m, T2SSession& S, int rows) {
  GraphCache& gc = graphs_of(m);
  const unsigned long long opt = m.options_gen;
  for (auto& g : gc.e)
    if (g.rows == rows && g.cap == S.cap && g.hist_ld == S.hist_ld && g.gen == S.w.key && g.opt == opt) {
      g.last_use = ++gc.tick;
      return g;
    }
  // drop entries whose buffers have moved or whose options changed, keep the cache small
  for (size_t i = 0; i < gc.e.size();) {
    if (gc.e[i].gen != S.w.key || gc.e[i].opt != opt) {
      cudaGraphExecDestroy(gc.e[i].exec);
      gc.e.erase(gc.e.begin() + i);
    } else ++i;
  }
  if (gc.e.size() >= 12) {
    size_t old = 0;
    for (size_t i = 1; i < gc.e.size(); ++i) if (gc.e[i].last_use < gc.e[old].last_use) old = i;
    cudaGraphExecDestroy(gc.e[old].exec);
    gc.e.erase(gc.e.begin() + old);
  }
  cudaStream_t s = m.stream;
  const StepBufs& w = S.w;
  const int B = rows;
  cudaGraph_t g = nullptr;
  unsigned long long captured = 0;
  GENIE_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  t_capture_counter = &captured;
  try {
    int nb = (B >= m.decode_split_min && m.stream2) ? m.decode_branches : 1;
    if (nb > 1 && (B + nb - 1) / nb > 128) nb = (B + 127) / 128;   // keep every branch on the <= 128-row GEMM
    if (nb > 4) nb = 4;
    while (nb > 1 && B / nb < 16) --nb;              // every branch stays on the tensor-core path
    if (nb > 1) {
      // independent slot ranges as parallel graph branches: every decode kernel is latency-bound
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
          w2.hist_len += b0; w2.kv_len += b0; w2.active += b0; w2.stop_step += b0; w2.params += b0;
          w2.kv = static_cast<char*>(w.kv) + (size_t)b0 * w.utt_stride * (w.kv_f16 ? 2 : 4);   // strides are in elements
          if (k > 0) GENIE_CUDA(cudaStreamWaitEvent(bs[k], m.ev_fork, 0));
          m.stream = bs[k];
          decode_step(m, w2, b1 - b0);
          if (k > 0) GENIE_CUDA(cudaEventRecord(bj[k], bs[k]));
        }
      } catch (...) { m.stream = keep; throw; }
      m.stream = keep;
      for (int k = 1; k < nb; ++k) GENIE_CUDA(cudaStreamWaitEvent(s, bj[k], 0));
    } else {
      StepBufs w1 = w;
      w1.part_stride = (long long)B * D;
      decode_step(m, w1, B);
    }
  } catch (...) {
    t_capture_counter = nullptr;
    cudaStreamEndCapture(s, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  t_capture_counter = nullptr;
  GENIE_CUDA(cudaStreamEndCapture(s, &g));
  StepGraph e;
  e.rows = rows; e.cap = S.cap; e.hist_ld = S.hist_ld; e.gen = S.w.key; e.opt = opt; e.launches = captured;
  cudaError_t ie = cudaGraphInstantiate(&e.exec, g, 0);
  cudaGraphDestroy(g);
  GENIE_CUDA(ie);
  e.last_use = ++gc.tick;
  gc.e.push_back(e);
  return gc.e.back();
}

// one decode step over the first `rows` slots
void run_step(Model& m, T2SSession& S, int rows) {
  const bool can_graph = m.use_graph && !m.record_logits && !g_sync_debug;
  if (can_graph) {
    StepGraph& g = step_graph(m, S, rows);
    GENIE_CUDA(cudaGraphLaunch(g.exec, m.stream));
    g_launches += g.launches;
  } else {
    StepBufs w1 = S.w;
    w1.part_stride = (long long)rows * D;
    decode_step(m, w1, rows);
  }
  record_logits_rows(m, S.LOGITS, rows);
}

void check_step_errors(Model& m) {
  check_tc_error(m);
  if (m.step_sync) {
    unsigned flag[2] = {0, 0};
    GENIE_CUDA(cudaMemcpy(flag, m.step_sync, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag[1]) {
      cudaMemset(m.step_sync, 0, sizeof(flag));
      GENIE_CHECK(false, "persistent decode step: device-wide barrier timed out");
    }
  }
}

int slot_idx(const SlotHost& h, int hist_len, int stop_step) {
  // the reference's loop variable at exit (Inference.py:95-106): index of the step whose stop flag fired, else the
  // last index run
  const int steps_run = hist_len - h.Ly - 1;
  int idx = (!h.honour || stop_step < 0) ? steps_run - 1 : stop_step - (h.Ly + 1);
  return idx < 0 ? 0 : idx;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// batch API (reference-facing path: Inference.py:63-109 for B utterances at once)
// ---------------------------------------------------------------------------------------------------------------
void t2s_prefill(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len,
                 const float* text_bert, const SamplingCfg& cfg, int io_dev) {
  GENIE_CHECK(m.finalized, "model not finalized");
  GENIE_CHECK(B > 0, "empty batch");
  const int max_steps = cfg.fixed_steps > 0 ? cfg.fixed_steps : cfg.max_steps;
  int maxS = 0, maxLy = 0;
  for (int b = 0; b < B; ++b) {
    GENIE_CHECK(prompts[b] != nullptr, "null prompt");
    GENIE_CHECK(text_len[b] > 0, "empty text_seq");
    maxS = std::max(maxS, prompts[b]->Lr + text_len[b] + prompts[b]->Ly);
    maxLy = std::max(maxLy, prompts[b]->Ly);
  }
  T2SSession& S = pool_build(m, B, maxS + max_steps + 1, maxLy + max_steps + 2, false);
  S.B = B; S.max_steps = max_steps; S.fixed = cfg.fixed_steps > 0;
  m.logits_host.clear();
  admit(m, S, B, nullptr, prompts, text_seq, text_len, text_bert, &cfg, 1, io_dev);
  // capture (or find) the decode graph now: the first decode_steps call then only replays
  if (m.use_graph && !m.record_logits && !g_sync_debug) step_graph(m, S, B);
}

// up to n_steps more decode steps for the batch in flight (stage graph + sampler per step, Inference.py:95-106);
// returns GENIE_CANCELLED (2) when the host flag fired, 0 otherwise
int t2s_decode_steps(Model& m, int n_steps, const volatile int* cancel, int* n_active_out, int* steps_done_out) {
  T2SSession& S = session_of(m);
  GENIE_CHECK(!S.pool_mode, "this handle runs a slot pool: use genie_t2s_pool_step");
  const DecodeTokenScope token(m);           // partitioned device: one handle decodes at a time
  cudaStream_t s = m.stream;
  GENIE_CUDA(cudaSetDevice(m.device));
  const int B = S.B;
  int rc = 0;
  std::vector<int> h_act(B, 1);
  const int poll_every = 8;
  cudaEvent_t ea, eb;
  GENIE_CUDA(cudaEventCreate(&ea)); GENIE_CUDA(cudaEventCreate(&eb));
  S.decode_spans.emplace_back(ea, eb);
  GENIE_CUDA(cudaEventRecord(ea, s));
  for (int k = 0; k < n_steps && S.steps_done < S.max_steps && !S.all_stopped; ++k) {
    if (cancel && *cancel) { rc = 2 /* GENIE_CANCELLED */; break; }
    run_step(m, S, B);
    ++S.steps_done;
    if (!S.fixed && (S.steps_done % poll_every == 0 || m.record_logits)) {
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
    if (S.fixed) n = S.steps_done < S.max_steps ? B : 0;
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
  GENIE_CHECK(!S.pool_mode, "this handle runs a slot pool: use genie_t2s_pool_read");
  cudaStream_t s = m.stream;
  GENIE_CUDA(cudaSetDevice(m.device));
  const int B = S.B;
  const StepBufs& w = S.w;
  GENIE_CHECK(y_ld >= S.hist_ld || y_out == nullptr, "y_ld too small: need >= " + std::to_string(S.hist_ld));

  std::vector<int> h_len(B), h_stopv(B), h_kvlen_final(B);
  GENIE_CUDA(cudaMemcpyAsync(h_len.data(), S.d_histlen, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  GENIE_CUDA(cudaMemcpyAsync(h_kvlen_final.data(), S.d_kvlen, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  GENIE_CUDA(cudaMemcpyAsync(h_stopv.data(), S.d_stop, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (y_out) {
    hist_to_i64_kernel<<<B, 256, 0, s>>>(S.HIST, S.hist_ld, S.d_histlen, S.Y64, S.hist_ld, B);
    GENIE_LAUNCHED("hist_to_i64");
    GENIE_CUDA(cudaMemcpy2DAsync(y_out, (size_t)y_ld * 8, S.Y64, (size_t)S.hist_ld * 8, (size_t)S.hist_ld * 8, B,
                                 io_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  GENIE_CUDA(cudaStreamSynchronize(s));
  check_step_errors(m);
  for (int b = 0; b < B; ++b) {
    if (y_len_out) y_len_out[b] = h_len[b];
    if (idx_out) idx_out[b] = slot_idx(S.slots[b], h_len[b], h_stopv[b]);
  }
  // measurement aid (bench.py roofline): the dominant decode kernel sits inside a CUDA graph, so it is
  // replayed here on the final cache state, all layers back to back (GBs of KV >> L2), between events
  if (m.time_attention > 0 && B > m.skinny_max_rows) {
    cudaEvent_t ea, eb;
    cudaEventCreate(&ea); cudaEventCreate(&eb);
    const float scale = 1.0f / std::sqrt(32.0f);
    GENIE_CUDA(cudaEventRecord(ea, s));
    for (int r = 0; r < m.time_attention; ++r)
      for (int l = 0; l < 24; ++l)
        launch_decode_attention_fused(w.part, 2, (long long)B * D * 3, m.layers[l].qkv.b, w.att, w.kv, w.kv_f16,
                                      w.utt_stride, l * w.layer_stride, w.v_off, w.kv_len, nullptr, B, w.cap, scale, s);
    GENIE_CUDA(cudaEventRecord(eb, s));
    GENIE_CUDA(cudaEventSynchronize(eb));
    float ms = 0;
    cudaEventElapsedTime(&ms, ea, eb);
    double kv_tokens = 0;
    for (int b = 0; b < B; ++b) kv_tokens += h_kvlen_final[b];
    m.timing[8] = 1000.f * ms / (float)(m.time_attention * 24);
    m.timing[9] = (float)(kv_tokens * 2 * 512 * (w.kv_f16 ? 2 : 4) / 1e6);
    cudaEventDestroy(ea); cudaEventDestroy(eb);
  }
  float t01 = 0, t12 = 0;
  cudaEventElapsedTime(&t01, S.ev0, S.ev1);
  for (auto& e : S.decode_spans) {
    float t = 0;
    cudaEventElapsedTime(&t, e.first, e.second);
    t12 += t;
  }
  m.timing[0] = t01; m.timing[1] = t12; m.timing[2] = t01 + t12; m.timing[3] = (float)S.steps_done;
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

// ---------------------------------------------------------------------------------------------------------------
// continuous batching (SURVEY 8f-1 / BASELINE config 5): slots admitted and released while others decode
// ---------------------------------------------------------------------------------------------------------------
void t2s_pool_create(Model& m, int n_slots, int kv_cap, int max_prompt_tokens, int max_steps) {
  GENIE_CHECK(n_slots >= 1 && n_slots <= 1024, "pool: n_slots must be 1..1024");
  GENIE_CHECK(max_steps >= 1 && max_prompt_tokens >= 1, "pool: bad limits");
  pool_build(m, n_slots, kv_cap, max_prompt_tokens + max_steps + 2, true);
  GENIE_CUDA(cudaStreamSynchronize(m.stream));
}

void t2s_pool_admit(Model& m, int n, const int* slots, Prompt* const* prompts, const int64_t* text_seq,
                    const int* text_len, const float* text_bert, const SamplingCfg* cfgs) {
  T2SSession& S = session_of(m);
  GENIE_CHECK(S.pool_mode, "no slot pool on this handle: call genie_t2s_pool_create first");
  GENIE_CHECK(slots != nullptr, "null slots");
  admit(m, S, n, slots, prompts, text_seq, text_len, text_bert, cfgs, n, 0);
}

// up to n_steps decode steps over every active slot; stops early once no slot is active.  The step graph covers
// the slots [0, rows) with rows = the smallest bucket holding the highest slot in use (free slots exit early)
int t2s_pool_step(Model& m, int n_steps, int* n_active_out) {
  T2SSession& S = session_of(m);
  GENIE_CHECK(S.pool_mode, "no slot pool on this handle");
  const DecodeTokenScope token(m);
  GENIE_CUDA(cudaSetDevice(m.device));
  cudaStream_t s = m.stream;
  int hi = -1;
  for (int i = 0; i < S.n_slots; ++i) if (S.slots[i].in_use) hi = i;
  int n_act = 0;
  if (hi >= 0) {
    static const int buckets[] = {1, 2, 4, 8, 16, 32, 64, 96, 128, 192, 256, 384, 512, 768, 1024};
    int rows = S.n_slots;
    for (int bkt : buckets) if (bkt >= hi + 1) { rows = std::min(bkt, S.n_slots); break; }
    std::vector<int> h_act(rows, 0);
    for (int k = 0; k < n_steps; ++k) {
      run_step(m, S, rows);
      if ((k & 7) == 7 || k + 1 == n_steps) {
        GENIE_CUDA(cudaMemcpyAsync(h_act.data(), S.d_active, rows * sizeof(int), cudaMemcpyDeviceToHost, s));
        GENIE_CUDA(cudaStreamSynchronize(s));
        n_act = 0;
        for (int b = 0; b < rows; ++b) n_act += h_act[b] ? 1 : 0;
        if (n_act == 0) break;
      }
    }
    check_step_errors(m);
  }
  if (n_active_out) *n_active_out = n_act;
  return 0;
}

// state[i]: 0 free, 1 decoding, 2 finished (stopped or budget used; read + release it); n_generated[i] = tokens
// generated so far (first-stage token included)
void t2s_pool_poll(Model& m, int* state, int* n_generated, int n) {
  T2SSession& S = session_of(m);
  GENIE_CHECK(S.pool_mode, "no slot pool on this handle");
  GENIE_CHECK(n >= S.n_slots, "poll: arrays shorter than the pool");
  GENIE_CUDA(cudaSetDevice(m.device));
  std::vector<int> act(S.n_slots), len(S.n_slots);
  GENIE_CUDA(cudaMemcpyAsync(act.data(), S.d_active, S.n_slots * sizeof(int), cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaMemcpyAsync(len.data(), S.d_histlen, S.n_slots * sizeof(int), cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaStreamSynchronize(m.stream));
  for (int i = 0; i < S.n_slots; ++i) {
    const SlotHost& h = S.slots[i];
    if (state) state[i] = !h.in_use ? 0 : (act[i] ? 1 : 2);
    if (n_generated) n_generated[i] = h.in_use ? len[i] - h.Ly : 0;
  }
}

void t2s_pool_read(Model& m, int slot, int64_t* y, int y_cap, int* y_len, int* idx) {
  T2SSession& S = session_of(m);
  GENIE_CHECK(S.pool_mode, "no slot pool on this handle");
  GENIE_CHECK(slot >= 0 && slot < S.n_slots && S.slots[slot].in_use, "read: slot not in use");
  GENIE_CUDA(cudaSetDevice(m.device));
  int len = 0, stop = -1;
  std::vector<int> row(S.hist_ld);
  GENIE_CUDA(cudaMemcpyAsync(&len, S.d_histlen + slot, sizeof(int), cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaMemcpyAsync(&stop, S.d_stop + slot, sizeof(int), cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaMemcpyAsync(row.data(), S.HIST + (size_t)slot * S.hist_ld, S.hist_ld * sizeof(int),
                             cudaMemcpyDeviceToHost, m.stream));
  GENIE_CUDA(cudaStreamSynchronize(m.stream));
  GENIE_CHECK(y == nullptr || y_cap >= len, "read: y too short, need " + std::to_string(len));
  if (y) for (int i = 0; i < len; ++i) y[i] = row[i];
  if (y_len) *y_len = len;
  if (idx) *idx = slot_idx(S.slots[slot], len, stop);
}

void t2s_pool_release(Model& m, int slot) {
  T2SSession& S = session_of(m);
  GENIE_CHECK(S.pool_mode, "no slot pool on this handle");
  GENIE_CHECK(slot >= 0 && slot < S.n_slots, "release: slot index out of range");
  GENIE_CUDA(cudaSetDevice(m.device));
  if (S.slots[slot].in_use) {   // a slot released while still decoding (cancelled request) stops here
    GENIE_CUDA(cudaMemsetAsync(S.d_active + slot, 0, sizeof(int), m.stream));
    S.slots[slot] = SlotHost{};
  }
}

void t2s_pool_info(Model& m, int* n_slots, int* kv_cap, int* hist_ld) {
  T2SSession& S = session_of(m);
  if (n_slots) *n_slots = S.n_slots;
  if (kv_cap) *kv_cap = S.cap;
  if (hist_ld) *hist_ld = S.hist_ld;
}

}  // namespace genie
