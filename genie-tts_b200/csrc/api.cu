// C-ABI (include/genie_b200.h).  Every entry point catches, records the message
// for genie_last_error() and returns a status; nothing here aborts.
#include "../../include/genie_b200.h"
#include "model.h"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <cuda_profiler_api.h>

using namespace genie;
namespace genie {
void tc_selftest(int M, int Cin, int Cout, int ntaps, int dil, int mode, int exact_w, float* max_err, float* ref_max);
}

struct genie_model { Model m; };
struct genie_prompt { Prompt p; };

namespace {
std::atomic<unsigned long long> g_model_uid{0};
thread_local std::string g_err;
template <typename F> int guarded(F&& f) {
  try {
    return f();
  } catch (const Error& e) {
    g_err = e.msg;
  } catch (const std::exception& e) {
    g_err = e.what();
  } catch (...) {
    g_err = "unknown error";
  }
  cudaGetLastError();   // clear sticky launch-config errors so the next call starts clean
  return 1;
}
}  // namespace

void genie::set_error(const std::string& msg) { g_err = msg; }

namespace {
// streams / events of one handle; `user_stream` non-null binds the handle to a caller-owned stream
void init_exec_state(Model& m, void* user_stream) {
  int least = 0, greatest = 0;
  GENIE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  static const bool prio = [] { const char* e = getenv("GENIE_STREAM_PRIO"); return !(e && e[0] == '0'); }();
  const int hi = prio ? greatest : least;
  if (user_stream) { m.stream = reinterpret_cast<cudaStream_t>(user_stream); m.stream_owned = false; }
  else {
    GENIE_CUDA(cudaStreamCreateWithPriority(&m.stream, cudaStreamNonBlocking, hi));
    if (prio) {
      GENIE_CUDA(cudaStreamCreateWithPriority(&m.stream_bulk, cudaStreamNonBlocking, least));
      GENIE_CUDA(cudaEventCreateWithFlags(&m.ev_bulk, cudaEventDisableTiming));
    }
  }
  GENIE_CUDA(cudaStreamCreateWithPriority(&m.stream2, cudaStreamNonBlocking, hi));
  GENIE_CUDA(cudaStreamCreateWithPriority(&m.stream3, cudaStreamNonBlocking, hi));
  GENIE_CUDA(cudaStreamCreateWithPriority(&m.stream4, cudaStreamNonBlocking, hi));
  GENIE_CUDA(cudaEventCreateWithFlags(&m.ev_join3, cudaEventDisableTiming));
  GENIE_CUDA(cudaEventCreateWithFlags(&m.ev_join4, cudaEventDisableTiming));
  GENIE_CUDA(cudaEventCreateWithFlags(&m.ev_fork, cudaEventDisableTiming));
  GENIE_CUDA(cudaEventCreateWithFlags(&m.ev_join, cudaEventDisableTiming));
}
}  // namespace

extern "C" {

const char* genie_last_error(void) { return g_err.c_str(); }
int genie_version(void) { return 200; }
unsigned long long genie_launch_count(void) { return g_launches.load(); }
int genie_host_alloc(size_t bytes, void** out) {
  return guarded([&] {
    GENIE_CHECK(out != nullptr && bytes > 0, "bad argument");
    GENIE_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return 0;
  });
}
int genie_host_free(void* p) {
  return guarded([&] { if (p) GENIE_CUDA(cudaFreeHost(p)); return 0; });
}
int genie_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int genie_model_create(int device, genie_model** out) {
  return guarded([&] {
    GENIE_CHECK(out != nullptr, "null out");
    int n = 0;
    GENIE_CUDA(cudaGetDeviceCount(&n));
    GENIE_CHECK(n > 0, "no CUDA device: genie_b200 has no CPU fallback");
    GENIE_CHECK(device >= 0 && device < n, "bad device index");
    GENIE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GENIE_CUDA(cudaGetDeviceProperties(&prop, device));
    GENIE_CHECK(prop.major == 10, "genie_b200 is built for sm_100a (B200) only; found sm_" +
                                      std::to_string(prop.major) + std::to_string(prop.minor));
    std::unique_ptr<genie_model> h(new genie_model());
    h->m.device = device;
    h->m.owner = std::make_shared<WeightOwner>();
    h->m.owner->device = device;
    h->m.owner->uid = ++g_model_uid;
    init_exec_state(h->m, nullptr);
    *out = h.release();
    return 0;
  });
}

int genie_context_create(genie_model* src, void* cuda_stream, genie_model** out) {
  return guarded([&] {
    GENIE_CHECK(src && out, "null argument");
    std::lock_guard<std::mutex> lock(src->m.mu);
    GENIE_CHECK(src->m.finalized, "context: the model must be finalized first");
    GENIE_CUDA(cudaSetDevice(src->m.device));
    std::unique_ptr<genie_model> h(new genie_model());
    static_cast<ModelWeights&>(h->m) = static_cast<const ModelWeights&>(src->m);   // shares the WeightOwner
    // options that select kernels travel with the clone
    h->m.use_graph = src->m.use_graph; h->m.use_tc = src->m.use_tc; h->m.tc_vits = src->m.tc_vits;
    h->m.tc_min_rows = src->m.tc_min_rows; h->m.skinny_max_rows = src->m.skinny_max_rows;
    h->m.fuse_pairs = src->m.fuse_pairs; h->m.persistent_step = src->m.persistent_step; h->m.kv_fp16 = src->m.kv_fp16; h->m.prefill_single = src->m.prefill_single;
    h->m.decode_split_min = src->m.decode_split_min; h->m.decode_branches = src->m.decode_branches;
    init_exec_state(h->m, cuda_stream);
    h->m.tc_err = dev_alloc<int>(h->m.ctx_owned, 1);
    GENIE_CUDA(cudaMemset(h->m.tc_err, 0, sizeof(int)));
    *out = h.release();
    return 0;
  });
}

int genie_set_stream(genie_model* h, void* cuda_stream) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    Model& m = h->m;
    GENIE_CUDA(cudaSetDevice(m.device));
    GENIE_CUDA(cudaStreamSynchronize(m.stream));
    m.t2s_graphs.reset();                       // captured steps were recorded on the old stream's branches
    if (m.stream && m.stream_owned) GENIE_CUDA(cudaStreamDestroy(m.stream));
    if (m.stream_bulk) { cudaStreamSynchronize(m.stream_bulk); cudaStreamDestroy(m.stream_bulk); m.stream_bulk = nullptr; }
    if (cuda_stream) { m.stream = reinterpret_cast<cudaStream_t>(cuda_stream); m.stream_owned = false; }
    else { GENIE_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking)); m.stream_owned = true; }
    return 0;
  });
}

int genie_model_add_tensor(genie_model* h, int graph, const char* name, const void* host_data, int dtype,
                           const int64_t* dims, int ndim) {
  return guarded([&] {
    GENIE_CHECK(h && name && host_data, "null argument");
    GENIE_CHECK(graph >= 0 && graph < 4, "bad graph id");
    GENIE_CHECK(!h->m.finalized, "model already finalized");
    Model& m = h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    GENIE_CUDA(cudaSetDevice(m.device));
    RawTensor t;
    t.f16 = dtype == GENIE_F16;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) { t.dims.push_back(dims[i]); t.numel *= dims[i]; }
    const size_t bytes = (size_t)t.numel * (t.f16 ? 2 : 4);
    GENIE_CUDA(cudaMalloc(&t.d, std::max<size_t>(bytes, 16)));
    m.owner->owned.push_back(t.d);
    m.weight_bytes += bytes;
    GENIE_CUDA(cudaMemcpyAsync(t.d, host_data, bytes, cudaMemcpyHostToDevice, m.stream));
    // the host buffer may be a transient view: finish the copy before returning
    GENIE_CUDA(cudaStreamSynchronize(m.stream));
    m.raw[graph][name] = t;
    return 0;
  });
}

int genie_model_set_constants(genie_model* h, const float* div, int top_k, float penalty, float temperature,
                              float noise_scale) {
  return guarded([&] {
    GENIE_CHECK(h && div, "null argument");
    Model& m = h->m;
    GENIE_CUDA(cudaSetDevice(m.device));
    if (!m.div_term) m.div_term = dev_alloc<float>(m.owner->owned, 256);
    GENIE_CUDA(cudaMemcpy(m.div_term, div, 256 * sizeof(float), cudaMemcpyHostToDevice));
    m.top_k = top_k; m.penalty = penalty; m.temperature = temperature; m.noise_scale = noise_scale;
    return 0;
  });
}

int genie_model_finalize(genie_model* h) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    GENIE_CUDA(cudaSetDevice(h->m.device));
    model_finalize(h->m);
    return 0;
  });
}

int genie_model_info(const genie_model* h, int* is_v2pp, long long* weight_bytes, long long* workspace_bytes) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    if (is_v2pp) *is_v2pp = h->m.v2pp;
    if (weight_bytes) *weight_bytes = (long long)h->m.weight_bytes;
    if (workspace_bytes) *workspace_bytes = (long long)h->m.ws.total();
    return 0;
  });
}

void genie_model_destroy(genie_model* h) {
  if (!h) return;
  {
    // a call still running on another thread finishes first; prompts built for this model stay valid objects
    // (they never dereference the model) and can be destroyed in any order
    std::lock_guard<std::mutex> lock(h->m.mu);
    cudaSetDevice(h->m.device);
    cudaStreamSynchronize(h->m.stream);
    cudaDeviceSynchronize();
  }
  delete h;
}

static int prompt_create_impl(genie_model* h, const int64_t* ref_seq, int Lr, const float* ref_bert, const float* ssl,
                              int Ts, const float* ref_audio, int n_audio, const float* sv_emb, const float* ge,
                              int ge_dim, const float* ge_adv, genie_prompt** out) {
  return guarded([&] {
    GENIE_CHECK(h && ref_seq && ssl && out, "null argument");
    std::lock_guard<std::mutex> lock(h->m.mu);
    genie_prompt* p = new genie_prompt();
    try {
      prompt_build(h->m, p->p, ref_seq, Lr, ref_bert, ssl, Ts, ref_audio, n_audio, sv_emb, ge, ge_dim, ge_adv);
    } catch (...) {
      delete p;
      throw;
    }
    *out = p;
    return 0;
  });
}

int genie_prompt_create(genie_model* h, const int64_t* ref_seq, int Lr, const float* ref_bert, const float* ssl,
                        int Ts, const float* ref_audio, int n_audio, const float* sv_emb, genie_prompt** out) {
  return prompt_create_impl(h, ref_seq, Lr, ref_bert, ssl, Ts, ref_audio, n_audio, sv_emb, nullptr, 0, nullptr, out);
}
int genie_prompt_create_with_ge(genie_model* h, const int64_t* ref_seq, int Lr, const float* ref_bert,
                                const float* ssl, int Ts, const float* ge, int ge_dim, const float* ge_advanced,
                                genie_prompt** out) {
  if (!ge) { g_err = "null ge"; return 1; }
  return prompt_create_impl(h, ref_seq, Lr, ref_bert, ssl, Ts, nullptr, 0, nullptr, ge, ge_dim, ge_advanced, out);
}

int genie_prompt_info(const genie_prompt* p, int* n_prompt_tokens, int* ge_dim, int* ref_len) {
  return guarded([&] {
    GENIE_CHECK(p, "null prompt");
    if (n_prompt_tokens) *n_prompt_tokens = p->p.Ly;
    if (ge_dim) *ge_dim = p->p.ge_dim;
    if (ref_len) *ref_len = p->p.Lr;
    return 0;
  });
}

int genie_prompt_read(const genie_prompt* p, int64_t* prompts, float* ge, float* ge_advanced) {
  return guarded([&] {
    GENIE_CHECK(p, "null prompt");
    GENIE_CUDA(cudaSetDevice(p->p.device));
    if (prompts) std::memcpy(prompts, p->p.prompts_host.data(), p->p.prompts_host.size() * sizeof(int64_t));
    if (ge) GENIE_CUDA(cudaMemcpy(ge, p->p.ge, p->p.ge_dim * 4, cudaMemcpyDeviceToHost));
    if (ge_advanced) GENIE_CUDA(cudaMemcpy(ge_advanced, p->p.ge_mrte, 512 * 4, cudaMemcpyDeviceToHost));
    return 0;
  });
}

void genie_prompt_destroy(genie_prompt* p) {
  if (!p) return;
  // no model access: the model may already be gone.  cudaFree (in ~Prompt) waits for work that may still read the
  // prompt's buffers.
  delete p;
}

namespace {
SamplingCfg sampling_cfg(const Model& m, const genie_sampling* sp) {
  SamplingCfg cfg;
  cfg.top_k = (sp && sp->top_k > 0) ? sp->top_k : m.top_k;
  cfg.temperature = (sp && sp->temperature > 0.f) ? sp->temperature : m.temperature;
  cfg.penalty = (sp && sp->repetition_penalty > 0.f) ? sp->repetition_penalty : m.penalty;
  cfg.top_p = (sp && sp->top_p > 0.f && sp->top_p < 1.f) ? sp->top_p : 1.0f;
  cfg.greedy = sp ? sp->greedy : 0;
  cfg.seed = sp ? sp->seed : 0;
  cfg.max_steps = (sp && sp->max_steps > 0) ? sp->max_steps : 500;
  cfg.fixed_steps = sp ? sp->fixed_steps : 0;
  return cfg;
}
}  // namespace

int genie_t2s_generate(genie_model* h, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                       const int* text_len, const float* text_bert, const genie_sampling* sp,
                       const volatile int* cancel, int io_on_device, int64_t* y, int y_ld, int* y_len, int* idx) {
  return guarded([&] {
    GENIE_CHECK(h && prompts && text_seq && text_len, "null argument");
    Model& m = h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    std::vector<Prompt*> ps(B);
    for (int b = 0; b < B; ++b) { GENIE_CHECK(prompts[b], "null prompt"); ps[b] = &prompts[b]->p; }
    return t2s_generate(m, ps.data(), B, text_seq, text_len, text_bert, sampling_cfg(m, sp), cancel, io_on_device, y,
                        y_ld, y_len, idx);
  });
}

int genie_t2s_prefill(genie_model* h, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                      const int* text_len, const float* text_bert, const genie_sampling* sp, int io_on_device) {
  return guarded([&] {
    GENIE_CHECK(h && prompts && text_seq && text_len, "null argument");
    Model& m = h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    std::vector<Prompt*> ps(B);
    for (int b = 0; b < B; ++b) { GENIE_CHECK(prompts[b], "null prompt"); ps[b] = &prompts[b]->p; }
    t2s_prefill(m, ps.data(), B, text_seq, text_len, text_bert, sampling_cfg(m, sp), io_on_device);
    return 0;
  });
}

int genie_t2s_decode_steps(genie_model* h, int n_steps, const volatile int* cancel, int* n_active, int* steps_done) {
  return guarded([&] {
    GENIE_CHECK(h && n_steps >= 0, "bad argument");
    std::lock_guard<std::mutex> lock(h->m.mu);
    return t2s_decode_steps(h->m, n_steps, cancel, n_active, steps_done);
  });
}

int genie_t2s_read(genie_model* h, int io_on_device, int64_t* y, int y_ld, int* y_len, int* idx) {
  return guarded([&] {
    GENIE_CHECK(h, "null argument");
    std::lock_guard<std::mutex> lock(h->m.mu);
    t2s_read(h->m, io_on_device, y, y_ld, y_len, idx);
    return 0;
  });
}

int genie_vits_decode(genie_model* h, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                      const int* text_len, const int64_t* sem, const int* sem_len, const float* zp_noise,
                      unsigned long long seed, float noise_scale, int io_on_device, float* audio, int* audio_len,
                      const int* noise_ids) {
  return guarded([&] {
    GENIE_CHECK(h && prompts && text_seq && text_len && sem && sem_len, "null argument");
    std::lock_guard<std::mutex> lock(h->m.mu);
    std::vector<Prompt*> ps(B);
    for (int b = 0; b < B; ++b) { GENIE_CHECK(prompts[b], "null prompt"); ps[b] = &prompts[b]->p; }
    vits_decode(h->m, ps.data(), B, text_seq, text_len, sem, sem_len, zp_noise, seed, noise_scale, io_on_device,
                audio, audio_len, noise_ids);
    return 0;
  });
}

int genie_debug_record_logits(genie_model* h, int enable) {
  if (!h) return 1;
  h->m.record_logits = enable != 0;
  return 0;
}
int genie_debug_read_logits(genie_model* h, float* out, int max_floats, int* n_floats) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    const std::vector<float>& v = h->m.logits_host;
    if (n_floats) *n_floats = (int)v.size();
    if (out) std::memcpy(out, v.data(), std::min<size_t>(v.size(), (size_t)max_floats) * sizeof(float));
    return 0;
  });
}
int genie_debug_keep(genie_model* h, int enable) {
  if (!h) return 1;
  h->m.keep = enable != 0;
  if (!enable) h->m.kept.clear();
  return 0;
}
int genie_debug_read(genie_model* h, const char* what, float* out, long long max_floats, long long* n_floats) {
  return guarded([&] {
    GENIE_CHECK(h && what, "null argument");
    auto it = h->m.kept.find(what);
    GENIE_CHECK(it != h->m.kept.end(), std::string("no kept tensor named ") + what);
    if (n_floats) *n_floats = (long long)it->second.size();
    if (out) std::memcpy(out, it->second.data(), std::min<size_t>(it->second.size(), (size_t)max_floats) * sizeof(float));
    return 0;
  });
}
int genie_debug_tc_selftest(int M, int Cin, int Cout, int ntaps, int dil, int mode, int exact_w, float* max_err,
                            float* ref_max) {
  return guarded([&] {
    GENIE_CHECK(max_err && ref_max, "null argument");
    tc_selftest(M, Cin, Cout, ntaps, dil, mode, exact_w, max_err, ref_max);
    return 0;
  });
}
int genie_profiler_range(int on) {   // cudaProfilerStart/Stop: `ncu --profile-from-start off` captures only this range
  cudaError_t e = on ? cudaProfilerStart() : cudaProfilerStop();
  return e == cudaSuccess ? 0 : 1;
}
int genie_last_timing(genie_model* h, float* ms, int n) {
  if (!h || !ms) return 1;
  for (int i = 0; i < n && i < 12; ++i) ms[i] = h->m.timing[i];
  return 0;
}
int genie_set_option(genie_model* h, const char* key, int value) {
  if (!h || !key) return 1;
  std::lock_guard<std::mutex> lock(h->m.mu);
  Model& m = h->m;
  struct Opt { const char* name; int* field; bool baked; };   // baked: changes what a captured decode step contains
  const Opt opts[] = {{"use_graph", &m.use_graph, false}, {"time_attention", &m.time_attention, false},
                      {"persistent_step", &m.persistent_step, true}, {"use_tc", &m.use_tc, true},
                      {"tc_vits", &m.tc_vits, false}, {"fuse_pairs", &m.fuse_pairs, false},
                      {"kv_fp16", &m.kv_fp16, false}, {"prefill_single", &m.prefill_single, false},
                      {"decode_split_min", &m.decode_split_min, true}, {"decode_branches", &m.decode_branches, true},
                      {"skinny_max_rows", &m.skinny_max_rows, true}, {"tc_min_rows", &m.tc_min_rows, true}};
  if (std::strcmp(key, "sm_partition") == 0) {
    return guarded([&] { model_enable_partition(m, value); return 0; });
  }
  for (const Opt& o : opts)
    if (std::strcmp(key, o.name) == 0) {
      *o.field = value;
      if (o.baked) ++m.options_gen;
      return 0;
    }
  g_err = std::string("unknown option ") + key;
  return 1;
}

// ---- continuous batching -------------------------------------------------------------------------------------
int genie_t2s_pool_create(genie_model* h, int n_slots, int kv_capacity, int max_prompt_tokens, int max_steps) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    t2s_pool_create(h->m, n_slots, kv_capacity, max_prompt_tokens, max_steps);
    return 0;
  });
}
int genie_t2s_admit(genie_model* h, int n, const int* slots, genie_prompt* const* prompts, const int64_t* text_seq,
                    const int* text_len, const float* text_bert, const genie_sampling* sampling) {
  return guarded([&] {
    GENIE_CHECK(h && slots && prompts && text_seq && text_len && n > 0, "bad argument");
    Model& m = h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    std::vector<Prompt*> ps(n);
    std::vector<SamplingCfg> cfgs(n);
    for (int b = 0; b < n; ++b) {
      GENIE_CHECK(prompts[b], "null prompt");
      ps[b] = &prompts[b]->p;
      cfgs[b] = sampling_cfg(m, sampling ? sampling + b : nullptr);
    }
    t2s_pool_admit(m, n, slots, ps.data(), text_seq, text_len, text_bert, cfgs.data());
    return 0;
  });
}
int genie_t2s_pool_step(genie_model* h, int n_steps, int* n_active) {
  return guarded([&] {
    GENIE_CHECK(h && n_steps >= 0, "bad argument");
    std::lock_guard<std::mutex> lock(h->m.mu);
    return t2s_pool_step(h->m, n_steps, n_active);
  });
}
int genie_t2s_pool_poll(genie_model* h, int* state, int* n_generated, int n) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    t2s_pool_poll(h->m, state, n_generated, n);
    return 0;
  });
}
int genie_t2s_pool_read(genie_model* h, int slot, int64_t* y, int y_capacity, int* y_len, int* idx) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    t2s_pool_read(h->m, slot, y, y_capacity, y_len, idx);
    return 0;
  });
}
int genie_t2s_release(genie_model* h, int slot) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    t2s_pool_release(h->m, slot);
    return 0;
  });
}
int genie_t2s_pool_info(genie_model* h, int* n_slots, int* kv_capacity, int* hist_ld) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
    std::lock_guard<std::mutex> lock(h->m.mu);
    t2s_pool_info(h->m, n_slots, kv_capacity, hist_ld);
    return 0;
  });
}

// ---- sampler in isolation (parity tests): rows of host logits + histories -> tokens ----------------------------
int genie_debug_sample(genie_model* h, const float* logits, int rows, const int64_t* hist, int hist_ld,
                       const int* hist_len, const genie_sampling* sampling, const float* noise, int n_draws,
                       int64_t* tokens, int* stop) {
  return guarded([&] {
    GENIE_CHECK(h && logits && hist && hist_len && tokens && rows > 0 && n_draws > 0 && hist_ld > 0, "bad argument");
    Model& m = h->m;
    std::lock_guard<std::mutex> lock(m.mu);
    GENIE_CUDA(cudaSetDevice(m.device));
    cudaStream_t s = m.stream;
    const SamplingCfg cfg = sampling_cfg(m, sampling);
    Workspace& ws = m.ws;
    float* d_log = ws.get<float>("dbg.logits", (size_t)rows * 1025);
    int* d_hist = ws.get<int>("dbg.hist", (size_t)rows * hist_ld);
    int* d_len = ws.get<int>("dbg.len", rows);
    int* d_misc = ws.get<int>("dbg.misc", (size_t)4 * rows);          // kv_len | active | stop_step | scratch
    int* d_tok = ws.get<int>("dbg.tok", (size_t)rows * n_draws);
    int* d_stop = ws.get<int>("dbg.stopf", (size_t)rows * n_draws);
    SlotParams* d_par = ws.get<SlotParams>("dbg.params", rows);
    float* d_noise = noise ? ws.get<float>("dbg.noise", (size_t)rows * 1025 * n_draws) : nullptr;
    std::vector<int> hist32((size_t)rows * hist_ld);
    for (size_t i = 0; i < hist32.size(); ++i) {
      GENIE_CHECK(hist[i] >= 0 && hist[i] <= 1024, "debug_sample: history id out of range");
      hist32[i] = (int)hist[i];
    }
    std::vector<int> misc((size_t)4 * rows, 0);
    for (int b = 0; b < rows; ++b) {
      GENIE_CHECK(hist_len[b] >= 0 && hist_len[b] <= hist_ld, "debug_sample: bad hist_len");
      misc[rows + b] = 1; misc[2 * rows + b] = -1;
    }
    GENIE_CUDA(cudaMemcpyAsync(d_log, logits, (size_t)rows * 1025 * 4, cudaMemcpyHostToDevice, s));
    GENIE_CUDA(cudaMemcpyAsync(d_hist, hist32.data(), hist32.size() * 4, cudaMemcpyHostToDevice, s));
    GENIE_CUDA(cudaMemcpyAsync(d_len, hist_len, rows * 4, cudaMemcpyHostToDevice, s));
    GENIE_CUDA(cudaMemcpyAsync(d_misc, misc.data(), misc.size() * 4, cudaMemcpyHostToDevice, s));
    if (noise) GENIE_CUDA(cudaMemcpyAsync(d_noise, noise, (size_t)rows * 1025 * n_draws * 4, cudaMemcpyHostToDevice, s));
    std::vector<SlotParams> par(rows);
    for (int d = 0; d < n_draws; ++d) {
      for (int b = 0; b < rows; ++b) {
        SlotParams p{};
        p.top_k = cfg.top_k; p.greedy = cfg.greedy; p.honour_stop = 1; p.hist_max = 1 << 30;
        p.temperature = cfg.temperature; p.penalty = cfg.penalty; p.top_p = cfg.top_p; p.utt = b;
        p.seed = cfg.seed + (unsigned long long)d;            // draw d of row b: Philox key (seed + d, b, hist_len)
        par[b] = p;
      }
      GENIE_CUDA(cudaMemcpyAsync(d_par, par.data(), rows * sizeof(SlotParams), cudaMemcpyHostToDevice, s));
      SamplerArgs a{};
      a.logits = d_log; a.ld = 1025; a.hist = d_hist; a.hist_ld = hist_ld; a.hist_len = d_len; a.kv_len = d_misc;
      a.active = d_misc + rows; a.stop_step = d_misc + 2 * rows; a.params = d_par; a.slot_map = nullptr; a.B = rows;
      a.advance_kv = 0; a.check_stop = 1; a.dbg_noise = d_noise ? d_noise + (size_t)d * rows * 1025 : nullptr;
      a.dbg_no_append = 1; a.dbg_tokens = d_tok + (size_t)d * rows; a.dbg_stop = d_stop + (size_t)d * rows;
      launch_sampler(a, s);
      GENIE_CUDA(cudaStreamSynchronize(s));                   // `par` is re-filled for the next draw
    }
    std::vector<int> tok((size_t)rows * n_draws), st((size_t)rows * n_draws);
    GENIE_CUDA(cudaMemcpy(tok.data(), d_tok, tok.size() * 4, cudaMemcpyDeviceToHost));
    GENIE_CUDA(cudaMemcpy(st.data(), d_stop, st.size() * 4, cudaMemcpyDeviceToHost));
    for (int d = 0; d < n_draws; ++d)
      for (int b = 0; b < rows; ++b) {                        // out: [rows, n_draws]
        tokens[(size_t)b * n_draws + d] = tok[(size_t)d * rows + b];
        if (stop) stop[(size_t)b * n_draws + d] = st[(size_t)d * rows + b];
      }
    return 0;
  });
}

}  // extern "C"
