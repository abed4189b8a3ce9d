// Internal model / prompt / workspace structures.
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "../../include/genie_b200.h"
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace genie {

struct DevBuf {
  void* p = nullptr; size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  // grow-only; contents are NOT preserved
  bool reserve(size_t n) {
    if (n <= bytes) return false;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    size_t want = n + n / 8 + 256;
    GENIE_CUDA(cudaMalloc(&p, want));
    bytes = want;
    return true;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct RawTensor {
  void* d = nullptr; int f16 = 0; std::vector<int64_t> dims; long long numel = 0;
};

// tcgen05 operand: fp16 [Cout][kpad] (K = tap*Cin + ci, zero padded to x64) + optional low part
struct TcW { const __half* hi = nullptr; const __half* lo = nullptr; int kpad = 0; const __half* tiles = nullptr; };
struct Linear {            // y = x W^T + b ; W [N,K] row-major
  const void* w = nullptr; int w_f16 = 0; const float* b = nullptr; int N = 0, K = 0; TcW tc;
};
struct Conv {              // repacked [Cout][k][Cin] fp32 (weight-norm folded)
  const float* w = nullptr; const float* b = nullptr; int Cout = 0, Cin = 0, k = 1; TcW tc;
};
struct ConvT {             // repacked [k][Cout][Cin] fp32
  const float* w = nullptr; const float* b = nullptr; int Cout = 0, Cin = 0, k = 0, stride = 0, pad = 0;
  TcW tc[10];              // one packed matrix per output phase r = (t_out + pad) mod stride
  // k == stride, pad == 0: the phases do not overlap, so the layer is ONE linear map [T, Cin] -> [T, k*Cout]
  // whose output rows are the k output frames side by side (= the channels-last output as it is)
  TcW tc_fused; const float* bias_fused = nullptr;
};

struct T2SLayer {
  Linear qkv, out, ff1, ff2;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct VitsEncLayer {
  Conv q, k, v, o, ff1, ff2;
  const float *rel_k, *rel_v, *g1, *b1, *g2, *b2;
};
struct WNLayer { Conv in, rs; };
struct FlowStep { Conv pre, post, cond; WNLayer wn[4]; };
struct ResBlock { Conv c1[3], c2[3]; int k = 3; };
struct MelStyle { Linear fc1, fc2, wq, wk, wv, fo, fc; Conv t0, t1; int out_dim = 512; };

struct Workspace {
  // named grow-only device buffers
  std::map<std::string, std::unique_ptr<DevBuf>> bufs;
  template <typename T> T* get(const std::string& name, size_t count) {
    auto& b = bufs[name];
    if (!b) b.reset(new DevBuf());
    if (b->reserve(count * sizeof(T))) ++generation;
    return b->as<T>();
  }
  size_t total() const { size_t t = 0; for (auto& kv : bufs) t += kv.second->bytes; return t; }
  unsigned long long generation = 0;   // bumped whenever any buffer moves (invalidates captured graphs)
};

// SM partition of one device (partition.cu): a decode partition and a bulk partition (CUDA green contexts) shared
// by every handle that opts in, plus the two tokens that make the handles alternate between them.
struct DevicePartition {
  int device = 0, decode_sms_requested = 0, decode_sms = 0, bulk_sms = 0;
  void* green_decode = nullptr; void* green_bulk = nullptr;      // CUgreenCtx
  std::mutex tok_decode, tok_bulk;
};
DevicePartition* device_partition(int device, int decode_sms);
cudaStream_t partition_stream(DevicePartition* p, bool decode, int priority);

// Device allocations holding a model's weights.  Shared (ref-counted) by the model handle and every execution
// context cloned from it (genie_context_create): the weights are freed when the last of them is destroyed.
struct WeightOwner {
  int device = 0;
  unsigned long long uid = 0;          // identity of the weight set (prompts remember it; never reused)
  std::vector<void*> owned;
  ~WeightOwner() { cudaSetDevice(device); for (void* p : owned) cudaFree(p); }
};

// Everything that is immutable after genie_model_finalize: plain pointers into the WeightOwner's allocations and
// architecture constants.  Copyable: a context is a copy of this plus its own execution state.
struct ModelWeights {
  int device = 0;
  bool finalized = false, v2pp = false;
  std::shared_ptr<WeightOwner> owner;
  std::unordered_map<std::string, RawTensor> raw[4];
  size_t weight_bytes = 0;
  int text_vocab = 732;               // rows of the phoneme embedding tables (input-id validation)
  int vits_text_vocab = 732;

  // constants
  float* div_term = nullptr;          // [256]
  int top_k = 15; float penalty = 1.35f, temperature = 1.0f, noise_scale = 0.5f;

  // T2S
  const float *text_emb = nullptr, *text_alpha = nullptr, *audio_emb = nullptr, *audio_alpha = nullptr;
  Linear bert_proj, predict;
  T2SLayer layers[24];
  // prompt-time VQ
  Conv ssl_vq;                        // k=2 stride-2 conv as a linear over row pairs
  const float *codebook_enc = nullptr, *codebook_enc_sq = nullptr;

  // VITS
  int gin = 512;                      // global-embedding width of flow/dec conditioning (512 V2, 1024 V2ProPlus)
  const float *codebook = nullptr, *vits_text_emb = nullptr;
  Conv ssl_proj, enc_proj;
  VitsEncLayer enc_ssl[3], enc_text[6], enc2[3];
  Conv mrte_c_pre, mrte_text_pre, mrte_q, mrte_k, mrte_v, mrte_o, mrte_c_post;
  FlowStep flow[4];                   // in execution order (flows.6, .4, .2, .0)
  Conv dec_pre, dec_cond;
  int n_up = 5;
  ConvT ups[5];
  ResBlock res[15];
  const float* conv_post = nullptr;   // [7][C_last]
  int c_last = 16;
  MelStyle ref_enc;                   // V2: vits ref_enc.*; V2ProPlus: prompt_encoder ref_enc.*
  Linear sv_emb, ge_to512; const float* prelu = nullptr;
  float* dft = nullptr;               // [1408, 2048]
  // persistent decode step (batch <= skinny_max_rows): per-layer pointer table
  void* step_layers_dev = nullptr; int num_sms = 0;
};

struct Model : ModelWeights {
  // ---- execution state: one per handle (a context cloned from a model has its own)
  cudaStream_t stream = nullptr;
  bool stream_owned = true;           // false: bound to a caller-owned stream (genie_set_stream / genie_context_create)
  // the decode-step graph runs the batch as up to 4 independent branches (contiguous utterance ranges) on
  // their own streams: every decode kernel is latency-bound, so the branches overlap
  cudaStream_t stream2 = nullptr, stream3 = nullptr, stream4 = nullptr;
  // Bulk work (prefill, SoVITS) runs on a LOWER-priority stream than the decode step: when two handles share a GPU
  // (GENIE.tts_batch_stream, server contexts) the tiny latency-bound decode kernels of one batch are scheduled ahead
  // of the remaining CTAs of the other batch's vocoder kernels.  Null when the handle is bound to a caller's stream.
  cudaStream_t stream_bulk = nullptr;
  cudaEvent_t ev_bulk = nullptr;
  // option sm_partition = N: the decode streams live in an N-SM partition of the device, the bulk stream in the
  // rest (partition.cu); decode_sms is what the decode kernels may size their grids for
  DevicePartition* partition = nullptr;
  int decode_sms = 0;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr, ev_join4 = nullptr;
  int decode_split_min = 64;          // batches >= this are split; decode_branches = number of branches
  int decode_branches = 2;
  std::vector<void*> ctx_owned;       // device allocations of this handle (error flag, barrier words)
  std::mutex mu;                      // entry points of one handle are serialised (one scheduler thread per GPU
                                      // is the intended use; a second thread waits instead of corrupting state)

  Workspace ws;
  unsigned long long options_gen = 0; // bumped by genie_set_option for anything baked into a captured graph
  int use_graph = 1;
  // tcgen05 path: 0 = exact SIMT everywhere; T2S always runs x_hi+x_lo against fp16-exact weights;
  // tc_vits: 1 = x_hi . w_hi, 2 = (x_hi+x_lo) . w_hi, 3 = (x_hi+x_lo) . w_hi + x_hi . w_lo
  int use_tc = 1, tc_vits = 1, tc_min_rows = 9, skinny_max_rows = 8;
  int fuse_pairs = 1;                  // narrow resblock pairs as one kernel (tc_pair_conv.cu)
  int prefill_single = 0;              // experiment: prefill linears as ONE fp16 product (activations rounded to fp16)
  int lin_single_now = 0;              // set around the prefill layer loop when prefill_single is on
  int kv_fp16 = 1;                     // KV cache rows stored as fp16 (q, scores, accumulators fp32); 0 = fp32 rows
  int* tc_err = nullptr;               // device flag: 1 = tcgen05 mbarrier timeout, 2 = input id out of range
  unsigned* step_sync = nullptr;       // persistent step: [0] barrier counter, [1] barrier-timeout flag
  int persistent_step = 4;             // largest batch that takes the persistent step (0 = off, <= 8)
  int persistent_ok = -1;              // -1 unknown, 0 the device cannot hold the grid co-resident, 1 ok
  // debug
  bool record_logits = false, keep = false;
  std::vector<float> logits_host;
  std::map<std::string, std::vector<float>> kept;
  float timing[12] = {0};      // [8] decode-attention us / launch, [9] its KV MB / launch (time_attention option)
  std::shared_ptr<void> t2s_session;   // slot pool of the T2S stage (t2s.cu): batch in flight / continuous batching
  std::shared_ptr<void> t2s_graphs;    // captured decode steps (t2s.cu)
  int time_attention = 0;      // > 0: after t2s_generate replay the fused decode attention this many times per layer

  Model() = default;
  Model(const Model&) = delete; Model& operator=(const Model&) = delete;
  ~Model();
};

struct Prompt {
  int device = 0;
  unsigned long long model_uid = 0;   // WeightOwner::uid of the model it was built for (contexts share it)
  int Lr = 0, Ly = 0, ge_dim = 0;
  bool has_bert = false;
  long long* ref_seq = nullptr;       // device int64 [Lr]
  float* ref_bert = nullptr;          // device [Lr,1024] or null
  int* prompts = nullptr;             // device int32 [Ly]
  std::vector<int64_t> prompts_host;
  float* ge = nullptr;                // device [ge_dim]
  float* ge_mrte = nullptr;           // device [512] (V2: == ge; V2ProPlus: ge_advanced)
  float* flow_cond = nullptr;         // device [4][1536]
  float* dec_cond = nullptr;          // device [C0]
  std::vector<void*> owned;
  // never touches the model: a prompt may outlive it (the handle's owner decides the order)
  ~Prompt() { if (!owned.empty()) cudaSetDevice(device); for (void* p : owned) cudaFree(p); }
};

// weights.cu
void model_finalize(Model& m);
// stages
void prompt_build(Model& m, Prompt& p, const int64_t* ref_seq, int Lr, const float* ref_bert, const float* ssl, int Ts,
                  const float* ref_audio, int n_audio, const float* sv_emb, const float* ge_in, int ge_dim,
                  const float* ge_adv_in);
struct SamplingCfg { int top_k; float temperature, penalty, top_p; int greedy; unsigned long long seed; int max_steps, fixed_steps; };
// legacy batch API: one batch admitted at once into a pool sized for it
void t2s_prefill(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len,
                 const float* text_bert, const SamplingCfg& cfg, int io_dev);
int t2s_decode_steps(Model& m, int n_steps, const volatile int* cancel, int* n_active, int* steps_done);
void t2s_read(Model& m, int io_dev, int64_t* y, int y_ld, int* y_len, int* idx);
int t2s_generate(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len,
                 const float* text_bert, const SamplingCfg& cfg, const volatile int* cancel, int io_dev,
                 int64_t* y, int y_ld, int* y_len, int* idx);
// continuous batching: a pool of decode slots with fixed per-slot KV capacity
void t2s_pool_create(Model& m, int n_slots, int kv_cap, int max_prompt_tokens, int max_steps);
void t2s_pool_admit(Model& m, int n, const int* slots, Prompt* const* prompts, const int64_t* text_seq,
                    const int* text_len, const float* text_bert, const SamplingCfg* cfgs);
int t2s_pool_step(Model& m, int n_steps, int* n_active);
void t2s_pool_poll(Model& m, int* state, int* n_generated, int n);
void t2s_pool_read(Model& m, int slot, int64_t* y, int y_cap, int* y_len, int* idx);
void t2s_pool_release(Model& m, int slot);
void t2s_pool_info(Model& m, int* n_slots, int* kv_cap, int* hist_ld);
void vits_decode(Model& m, Prompt* const* prompts, int B, const int64_t* text_seq, const int* text_len,
                 const int64_t* sem, const int* sem_len, const float* zp_noise, unsigned long long seed,
                 float noise_scale, int io_dev, float* audio, int* audio_len, const int* noise_ids = nullptr);

// helpers shared by the stage files
// fp16 hi/lo pair: v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (operand hand-over between linears)
struct Half2Part { __half* hi; __half* lo; };
void run_linear(Model& m, const Linear& L, const float* x, int ldx, float* y, int ldy, int M, int act = ACT_NONE,
                const float* res = nullptr, int ldr = 0, int nt = 0, int ksplit = 1, long long split_stride = 0, const Half2Part* a16 = nullptr,
                const Half2Part* y16 = nullptr);
inline bool tc_linear_ok(const Model& m, const Linear& L, int M) { return m.use_tc && L.tc.hi && M >= m.tc_min_rows; }
void keep_tensor(Model& m, const char* name, const float* dev, long long n);
// RAII: route the launches of a bulk stage to the handle's low-priority stream (ordered after what is already queued
// on the main stream); on exit the main stream waits for the stage's work
struct BulkStreamScope {
  Model& m; cudaStream_t keep; bool locked = false;
  explicit BulkStreamScope(Model& mm);
  ~BulkStreamScope();
};
// RAII: the decode token of a partitioned device (no-op otherwise); released after the decode stream has drained
struct DecodeTokenScope {
  Model& m; bool locked = false;
  explicit DecodeTokenScope(Model& mm);
  ~DecodeTokenScope();
};
void model_enable_partition(Model& m, int decode_sms);
void check_tc_error(Model& m);
template <typename T> T* dev_alloc(std::vector<void*>& owned, size_t count) {
  void* p = nullptr;
  GENIE_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
  owned.push_back(p);
  return reinterpret_cast<T*>(p);
}

}  // namespace genie
