// T2S sampler (stage#[1775-1821], first_stage#[1789-1820]): repetition penalty
// over the whole token history, temperature, top-k (ties kept), softmax,
// argmax(probs / noise), EOS stop flag; appends the token to the history.
// noise == 1 (greedy) is the bit-comparable mode; otherwise noise ~ N(0,1) from
// a counter-based Philox4x32-10 stream keyed by (seed, utterance, step).
#include "kernels.cuh"
#include "philox.cuh"
#include <math_constants.h>

namespace genie {
namespace {

constexpr int V = 1025;
constexpr int EOS = 1024;

struct ArgVal { float v; int i; };
__device__ __forceinline__ ArgVal better(ArgVal a, ArgVal b) {   // max value, first index on ties
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ ArgVal block_argmax(ArgVal x, ArgVal* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgVal y; y.v = __shfl_xor_sync(0xffffffffu, x.v, o); y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  __syncthreads();
  if (lane == 0) sh[warp] = x;
  __syncthreads();
  ArgVal r = sh[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = better(r, sh[w]);
  return r;
}
__device__ float block_sum(float x, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  __syncthreads();
  if (lane == 0) sh[warp] = x;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) r += sh[w];
  return r;
}

__global__ void __launch_bounds__(256) sampler_kernel(SamplerArgs a) {
  pdl_trigger();
  pdl_wait();      // logits come from the predecessor GEMM
  __shared__ float raw[V];
  __shared__ float lg[V];
  __shared__ float pr[V];
  __shared__ unsigned char taken[V];
  __shared__ ArgVal sh_av[8];
  __shared__ float sh_f[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int slot = a.slot_map ? a.slot_map[b] : b;
  if (a.active && !a.active[slot]) return;
  const SlotParams P = a.params[slot];
  const float* lrow = a.logits + (long long)b * a.ld;
  int* hist = a.hist + (long long)slot * a.hist_ld;
  const int n = a.hist_len[slot];

  for (int i = tid; i < V; i += 256) { float v = lrow[i]; raw[i] = v; lg[i] = v; taken[i] = 0; }
  __syncthreads();
  // stop test part 1: argmax of the raw logits (stage#[1807-1809])
  ArgVal loc; loc.v = -CUDART_INF_F; loc.i = V;
  for (int i = tid; i < V; i += 256) { ArgVal c; c.v = raw[i]; c.i = i; loc = better(loc, c); }
  const ArgVal raw_best = block_argmax(loc, sh_av);

  // repetition penalty: every token of the history, from the raw logit (duplicates idempotent)
  for (int i = tid; i < n; i += 256) {
    int t = hist[i];
    float s = raw[t];
    lg[t] = (s < 0.f) ? s * P.penalty : s / P.penalty;
  }
  __syncthreads();
  // top-p (extension; absent from the reference graphs, off at 1.0): in descending order of the penalised logits
  // drop every token whose inclusive cumulative probability exceeds top_p, except the first
  if (P.top_p > 0.f && P.top_p < 1.f) {
    loc.v = -CUDART_INF_F; loc.i = V;
    for (int i = tid; i < V; i += 256) { ArgVal c; c.v = lg[i]; c.i = i; loc = better(loc, c); }
    const ArgVal top = block_argmax(loc, sh_av);
    float part = 0.f;
    for (int i = tid; i < V; i += 256) { float e = expf(lg[i] - top.v); pr[i] = e; part += e; }
    const float denom = block_sum(part, sh_f);
    __syncthreads();
    for (int i = tid; i < V; i += 256) {
      const float li = lg[i];
      float cum = 0.f;
      for (int j = 0; j < V; ++j) {
        const float lj = lg[j];
        if (lj > li || (lj == li && j <= i)) cum += pr[j];
      }
      taken[i] = (cum / denom > P.top_p && i != top.i) ? 2 : 0;
    }
    __syncthreads();
    for (int i = tid; i < V; i += 256) if (taken[i]) { lg[i] = -CUDART_INF_F; taken[i] = 0; }
    __syncthreads();
  }
  for (int i = tid; i < V; i += 256) lg[i] = lg[i] / P.temperature;
  __syncthreads();
  // k-th largest value counting duplicates: peel the maximum top_k times
  float kth = 0.f;
  const int top_k = P.top_k < 1 ? 1 : (P.top_k > V ? V : P.top_k);
  for (int r = 0; r < top_k; ++r) {
    loc.v = -CUDART_INF_F; loc.i = V;
    for (int i = tid; i < V; i += 256)
      if (!taken[i]) { ArgVal c; c.v = lg[i]; c.i = i; loc = better(loc, c); }
    ArgVal m = block_argmax(loc, sh_av);
    if (tid == 0 && m.i < V) taken[m.i] = 1;
    kth = m.v;
    __syncthreads();
  }
  // softmax over {lg >= kth}
  loc.v = -CUDART_INF_F; loc.i = V;
  for (int i = tid; i < V; i += 256) { ArgVal c; c.v = lg[i]; c.i = i; loc = better(loc, c); }
  const float mx = block_argmax(loc, sh_av).v;
  float part = 0.f;
  for (int i = tid; i < V; i += 256) {
    float e = (lg[i] < kth) ? 0.f : expf(lg[i] - mx);
    lg[i] = e; part += e;
  }
  const float denom = block_sum(part, sh_f);
  // token = argmax(probs / noise), first index on ties (ArgMax select_last_index=0)
  loc.v = -CUDART_INF_F; loc.i = V;
  for (int i = tid; i < V; i += 256) {
    float p = lg[i] / denom;
    float q = 1.f;
    if (!P.greedy) q = a.dbg_noise ? a.dbg_noise[(long long)b * V + i]
                                   : philox_normal(P.seed, (uint32_t)P.utt, (uint32_t)n, (uint32_t)i);
    ArgVal c; c.v = p / q; c.i = i;
    loc = better(loc, c);
  }
  const ArgVal tokv = block_argmax(loc, sh_av);
  if (tid == 0) {
    const int tok = tokv.i;
    const bool stop = (raw_best.i == EOS) || (tok == EOS);
    if (a.dbg_no_append) {
      a.dbg_tokens[b] = tok;
      if (a.dbg_stop) a.dbg_stop[b] = stop ? 1 : 0;
      return;
    }
    hist[n] = tok;
    a.hist_len[slot] = n + 1;
    if (a.advance_kv) a.kv_len[slot] += 1;
    if (a.check_stop && stop && a.stop_step[slot] < 0) {
      a.stop_step[slot] = n;   // history length before this token (graph-replay safe step id)
      if (P.honour_stop) a.active[slot] = 0;
    }
    if (n + 1 >= P.hist_max) a.active[slot] = 0;   // loop bound (Inference.py:95) / fixed token budget
  }
}

}  // namespace

void launch_sampler(const SamplerArgs& a, cudaStream_t s) {
  if (a.B <= 0) return;
  GENIE_CHECK(a.params != nullptr, "sampler: no per-slot parameters");
  launch_pdl(sampler_kernel, dim3(a.B), dim3(256), 0, s, a);
  GENIE_LAUNCHED("sampler");
}

}  // namespace genie
