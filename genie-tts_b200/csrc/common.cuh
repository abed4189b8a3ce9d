// Shared declarations for the genie_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>

namespace genie {

// ---- error plumbing: the C-ABI never aborts; it returns a status and keeps
// the message for genie_last_error() (reference convention: everything inside
// the worker is caught and logged, src/genie_tts/Core/TTSPlayer.py:109-114).
void set_error(const std::string& msg);
struct Error { std::string msg; };

#define GENIE_CUDA(expr)                                                            \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      throw ::genie::Error{std::string(#expr) + " failed: " + cudaGetErrorString(_e) +  \
                           " (" __FILE__ ":" + std::to_string(__LINE__) + ")"};     \
  } while (0)

#define GENIE_CHECK(cond, msg)                                                      \
  do {                                                                              \
    if (!(cond)) throw ::genie::Error{std::string(msg) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"}; \
  } while (0)

extern int g_sync_debug;   // GENIE_SYNC_DEBUG=1: synchronise after every launch and name the failing kernel
inline void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error{std::string(what) + " launch failed: " + cudaGetErrorString(e)};
  if (g_sync_debug) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess && e2 != cudaErrorStreamCaptureUnsupported)
      throw Error{std::string(what) + " failed at run time: " + cudaGetErrorString(e2)};
    (void)st;
    cudaGetLastError();
  }
}

// global kernel-launch counter (bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launches;
// during stream capture nothing launches: the capturing thread points this at its own counter (kernels per replay)
extern thread_local unsigned long long* t_capture_counter;
#define GENIE_LAUNCHED(name)                                                         \
  do {                                                                               \
    if (::genie::t_capture_counter) ++*::genie::t_capture_counter; else ++::genie::g_launches; \
    ::genie::check_launch(name);                                                     \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: one process may drive models
// on several GPUs (ReplicaPool: one scheduler thread per GPU), so every launch site keeps one of these per kernel
// instantiation and raises the limit on the CURRENT device the first time (or when a larger size is needed).
struct DynSmemAttr {
  static constexpr int MAX_DEV = 64;
  std::atomic<size_t> set[MAX_DEV] = {};
  std::mutex mu;
  template <typename K> void ensure(K kernel, size_t smem) {
    int dev = 0;
    GENIE_CUDA(cudaGetDevice(&dev));
    const bool tracked = dev >= 0 && dev < MAX_DEV;
    if (tracked && smem <= set[dev].load(std::memory_order_acquire)) return;
    std::lock_guard<std::mutex> lock(mu);
    if (tracked && smem <= set[dev].load(std::memory_order_relaxed)) return;
    GENIE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (tracked) set[dev].store(smem, std::memory_order_release);
  }
};

// ---------------------------------------------------------------------------
// Programmatic dependent launch (decode step chain).  A kernel launched through launch_pdl may start while
// its predecessor in the stream is still running (once every CTA of the predecessor has executed
// pdl_trigger or exited); it must call pdl_wait() before it reads anything an earlier kernel wrote or
// writes anything an earlier kernel reads.  Everything before pdl_wait (TMEM allocation, barrier init,
// loads of constant weights) overlaps the predecessor's tail.  Both are no-ops in a normal launch.
// ---------------------------------------------------------------------------
extern int g_pdl;          // GENIE_PDL=0 disables
extern thread_local int g_pdl_now;   // cleared by the caller for launch sequences where it does not pay (batch <= 8
                                    // decode: measured 58 -> 68 ms per 90 steps with it on); per host thread, as one
                                    // scheduler thread per GPU may be issuing launch sequences at the same time
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (g_pdl && g_pdl_now) ? 1 : 0;
  GENIE_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_MISH = 3, ACT_TANH = 4 };

// ---------------------------------------------------------------------------
// Implicit-GEMM 1-D convolution / linear layer on channels-last activations.
//   y[seg, out_mul*q + out_add, co] (op)= out_scale * act( sum_{m<ntaps} sum_{ci}
//        pre(x[seg, q + in_shift0 + m*in_shift_step, ci]) * W[co, m, ci] + bias[co] + bias2[seg, co] ) + res[...]
// Rows outside [0, T_seg) read as zero (conv zero padding at utterance edges).
// A linear layer is ntaps=1, in_shift0=0 over one segment of M rows.
// ---------------------------------------------------------------------------
struct ConvGemm {
  const float* x = nullptr;   int ldx = 0;     // input  rows x Cin   (fp32)
  const void*  w = nullptr;   int w_f16 = 0;   // weights: element (co, m, ci) at co*w_co_stride + m*w_tap_stride + ci
  long long w_co_stride = 0, w_tap_stride = 0;
  const float* bias = nullptr;                 // [Cout]
  const float* bias2 = nullptr; int ldb2 = 0;  // [B, ldb2] per-segment bias (conditioning)
  const float* res = nullptr; int ldr = 0;     // residual rows x Cout (output row indexing)
  float* y = nullptr;         int ldy = 0;
  int Cin = 0, Cout = 0;
  int ntaps = 1, in_shift0 = 0, in_shift_step = 1;
  int out_mul = 1, out_add = 0;
  float pre_slope = 1.f;                       // leaky-relu slope applied to x on load (1 = identity)
  int act = ACT_NONE; float act_slope = 0.f;
  float out_scale = 1.f;
  int accumulate = 0;                          // y += result
  // segments (utterances): rows of segment b are [in_off[b], in_off[b+1]) in x and
  // [out_off[b], out_off[b+1]) in y/res.  Null => one segment [0, M) / [0, M_out).
  const int* in_off = nullptr; const int* out_off = nullptr;
  int B = 1; int M = 0; int M_out = 0;         // M: max #q per segment (grid sizing); single-segment row counts
  int q_extra = 0;                             // q ranges over [0, T_in + q_extra)
  // tcgen05 path (tc_gemm.cu): weights pre-packed as fp16 [Cout][tc_kpad], K index = tap*Cin + ci,
  // zero padded to a multiple of 64; optional low part (w - fp16(w)); tc_split_a: x = x_hi + x_lo
  const __half* tc_w = nullptr; const __half* tc_wlo = nullptr; int tc_kpad = 0; int tc_split_a = 0;
  int tc_nt = 0;                               // force the N tile (0 = by Cout)
  const __half* tc_tiles = nullptr;            // pre-tiled weights for tc_halo_bulk_kernel (wide k-tap convs)
  // fp16 hand-over between the two convs of a resblock pair (tc_halo_conv only): the producer writes
  // fp16(lrelu(out)) to y16 instead of fp32 y, the consumer reads x16 as its A operand as is (no pre-activation)
  const __half* x16 = nullptr; __half* y16 = nullptr;
  // hi/lo-split form of the same hand-over for the (x_hi + x_lo) . w linears (tc_gemm.cu, ntaps == 1): the
  // producer writes fp16(v) and fp16(v - fp16(v)), the consumer copies both parts into its operand tiles
  const __half* x16_lo = nullptr; __half* y16_lo = nullptr;
  // split-K: CTA (n-tile, ks) reduces k-blocks [ks*KB/ksplit, (ks+1)*KB/ksplit) and stores the raw partial
  // at y + ks*split_stride; bias / residual / activation are then applied by the consumer (layernorm)
  int ksplit = 1; long long split_stride = 0;
};
void launch_conv_gemm(const ConvGemm& p, cudaStream_t s);
void launch_tc_conv_gemm(const ConvGemm& p, int* err_flag, cudaStream_t s);
// k-tap convs with the activation halo staged once per CTA (tc_halo_conv.cu); false => not applicable
bool try_launch_tc_halo_conv(const ConvGemm& p, int* err_flag, cudaStream_t s);
bool tc_halo_fp16_pair_ok(int C, int ntaps);   // both convs of a C -> C pair take tc_halo_conv_kernel
// one resblock pair (conv1 dilated, conv2 dilation 1, C -> C, C = 16 / 32) in one kernel (tc_pair_conv.cu)
bool tc_pair_conv_supported(int C, int k);
void launch_tc_pair_conv(const ConvGemm& conv2, const __half* w1, const float* bias1, int kpad1, int d1, int* err_flag,
                         cudaStream_t s);
bool pretile_w128_supported(int Cin, int Cout, int ntaps);
long long pretile_w128_halves(int Cin, int Cout, int ntaps);   // size of the pre-tiled copy (last N tile zero padded)
void launch_pretile_w128(const __half* hi, int Cout, int kpad, int Cin, int ntaps, __half* tiles, cudaStream_t s);
bool skinny_gemm_supported(const ConvGemm& p);
void launch_skinny_gemm(const ConvGemm& p, cudaStream_t s);

// ---------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------
struct Attn {
  const float* q = nullptr; int ldq = 0;       // [rows, H*d] (row-major, heads side by side)
  const float* k = nullptr; int ldk = 0;
  const float* v = nullptr; int ldv = 0;
  float* o = nullptr;       int ldo = 0;
  const int* q_off = nullptr;                  // [B+1] query segment offsets
  const int* kv_off = nullptr;                 // [B+1] key segment offsets
  int B = 1, H = 1, d = 32, max_q = 0;
  float scale = 1.f;                           // applied to q before q.k (and q.rel_k)
  int mask_mode = 0;                           // 0 none, 1 = T2S prefill (lx[b] text rows)
  const int* lx = nullptr;
  const float* rel_k = nullptr;                // [2*window+1, d] or null
  const float* rel_v = nullptr;
  int window = 0;
  int use_mma = 0;                             // head dim 32, no rel-pos: tensor-core kernel (hi/lo split operands)
};
void launch_attention(const Attn& p, cudaStream_t s);

}  // namespace genie
