// C-ABI (include/genie_b200.h).  Every entry point catches, records the message
// for genie_last_error() and returns a status; nothing here aborts.
#include "../../include/genie_b200.h"
#include "model.h"
#include <cstring>
#include <cuda_profiler_api.h>

using namespace genie;
namespace genie {
void tc_selftest(int M, int Cin, int Cout, int ntaps, int dil, int mode, int exact_w, float* max_err, float* ref_max);
}

struct genie_model { Model m; };
struct genie_prompt { Prompt p; };

namespace {
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

extern "C" {

const char* genie_last_error(void) { return g_err.c_str(); }
int genie_version(void) { return 100; }
unsigned long long genie_launch_count(void) { return g_launches; }
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
    genie_model* h = new genie_model();
    h->m.device = device;
    GENIE_CUDA(cudaStreamCreateWithFlags(&h->m.stream, cudaStreamNonBlocking));
    GENIE_CUDA(cudaStreamCreateWithFlags(&h->m.stream2, cudaStreamNonBlocking));
    GENIE_CUDA(cudaStreamCreateWithFlags(&h->m.stream3, cudaStreamNonBlocking));
    GENIE_CUDA(cudaStreamCreateWithFlags(&h->m.stream4, cudaStreamNonBlocking));
    GENIE_CUDA(cudaEventCreateWithFlags(&h->m.ev_join3, cudaEventDisableTiming));
    GENIE_CUDA(cudaEventCreateWithFlags(&h->m.ev_join4, cudaEventDisableTiming));
    GENIE_CUDA(cudaEventCreateWithFlags(&h->m.ev_fork, cudaEventDisableTiming));
    GENIE_CUDA(cudaEventCreateWithFlags(&h->m.ev_join, cudaEventDisableTiming));
    *out = h;
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
    GENIE_CUDA(cudaSetDevice(m.device));
    RawTensor t;
    t.f16 = dtype == GENIE_F16;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) { t.dims.push_back(dims[i]); t.numel *= dims[i]; }
    const size_t bytes = (size_t)t.numel * (t.f16 ? 2 : 4);
    GENIE_CUDA(cudaMalloc(&t.d, std::max<size_t>(bytes, 16)));
    m.owned.push_back(t.d);
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
    if (!m.div_term) m.div_term = dev_alloc<float>(m.owned, 256);
    GENIE_CUDA(cudaMemcpy(m.div_term, div, 256 * sizeof(float), cudaMemcpyHostToDevice));
    m.top_k = top_k; m.penalty = penalty; m.temperature = temperature; m.noise_scale = noise_scale;
    return 0;
  });
}

int genie_model_finalize(genie_model* h) {
  return guarded([&] {
    GENIE_CHECK(h, "null model");
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
  cudaSetDevice(h->m.device);
  cudaDeviceSynchronize();
  delete h;
}

static int prompt_create_impl(genie_model* h, const int64_t* ref_seq, int Lr, const float* ref_bert, const float* ssl,
                              int Ts, const float* ref_audio, int n_audio, const float* sv_emb, const float* ge,
                              int ge_dim, const float* ge_adv, genie_prompt** out) {
  return guarded([&] {
    GENIE_CHECK(h && ref_seq && ssl && out, "null argument");
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
    GENIE_CUDA(cudaSetDevice(p->p.model->device));
    if (prompts) std::memcpy(prompts, p->p.prompts_host.data(), p->p.prompts_host.size() * sizeof(int64_t));
    if (ge) GENIE_CUDA(cudaMemcpy(ge, p->p.ge, p->p.ge_dim * 4, cudaMemcpyDeviceToHost));
    if (ge_advanced) GENIE_CUDA(cudaMemcpy(ge_advanced, p->p.ge_mrte, 512 * 4, cudaMemcpyDeviceToHost));
    return 0;
  });
}

void genie_prompt_destroy(genie_prompt* p) {
  if (!p) return;
  if (p->p.model) { cudaSetDevice(p->p.model->device); cudaStreamSynchronize(p->p.model->stream); }
  delete p;
}

namespace {
SamplingCfg sampling_cfg(const Model& m, const genie_sampling* sp) {
  SamplingCfg cfg;
  cfg.top_k = (sp && sp->top_k > 0) ? sp->top_k : m.top_k;
  cfg.temperature = (sp && sp->temperature > 0.f) ? sp->temperature : m.temperature;
  cfg.penalty = (sp && sp->repetition_penalty > 0.f) ? sp->repetition_penalty : m.penalty;
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
    std::vector<Prompt*> ps(B);
    for (int b = 0; b < B; ++b) { GENIE_CHECK(prompts[b], "null prompt"); ps[b] = &prompts[b]->p; }
    t2s_prefill(m, ps.data(), B, text_seq, text_len, text_bert, sampling_cfg(m, sp), io_on_device);
    return 0;
  });
}

int genie_t2s_decode_steps(genie_model* h, int n_steps, const volatile int* cancel, int* n_active, int* steps_done) {
  return guarded([&] {
    GENIE_CHECK(h && n_steps >= 0, "bad argument");
    return t2s_decode_steps(h->m, n_steps, cancel, n_active, steps_done);
  });
}

int genie_t2s_read(genie_model* h, int io_on_device, int64_t* y, int y_ld, int* y_len, int* idx) {
  return guarded([&] {
    GENIE_CHECK(h, "null argument");
    t2s_read(h->m, io_on_device, y, y_ld, y_len, idx);
    return 0;
  });
}

int genie_vits_decode(genie_model* h, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                      const int* text_len, const int64_t* sem, const int* sem_len, const float* zp_noise,
                      unsigned long long seed, float noise_scale, int io_on_device, float* audio, int* audio_len) {
  return guarded([&] {
    GENIE_CHECK(h && prompts && text_seq && text_len && sem && sem_len, "null argument");
    std::vector<Prompt*> ps(B);
    for (int b = 0; b < B; ++b) { GENIE_CHECK(prompts[b], "null prompt"); ps[b] = &prompts[b]->p; }
    vits_decode(h->m, ps.data(), B, text_seq, text_len, sem, sem_len, zp_noise, seed, noise_scale, io_on_device,
                audio, audio_len);
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
  if (std::strcmp(key, "use_graph") == 0) { h->m.use_graph = value; return 0; }
  if (std::strcmp(key, "time_attention") == 0) { h->m.time_attention = value; return 0; }
  if (std::strcmp(key, "persistent_step") == 0) { h->m.persistent_step = value; h->m.step_graph_flags = -1; return 0; }
  if (std::strcmp(key, "use_tc") == 0) { h->m.use_tc = value; h->m.step_graph_flags = -1; return 0; }
  if (std::strcmp(key, "tc_vits") == 0) { h->m.tc_vits = value; return 0; }
  if (std::strcmp(key, "fuse_pairs") == 0) { h->m.fuse_pairs = value; return 0; }
  if (std::strcmp(key, "decode_split_min") == 0) { h->m.decode_split_min = value; h->m.step_graph_flags = -1; return 0; }
  if (std::strcmp(key, "decode_branches") == 0) { h->m.decode_branches = value; h->m.step_graph_flags = -1; return 0; }
  if (std::strcmp(key, "skinny_max_rows") == 0) { h->m.skinny_max_rows = value; h->m.step_graph_flags = -1; return 0; }
  if (std::strcmp(key, "tc_min_rows") == 0) { h->m.tc_min_rows = value; h->m.step_graph_flags = -1; return 0; }
  g_err = std::string("unknown option ") + key;
  return 1;
}

}  // extern "C"
