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
  __shared__ unsigned char taken[V];
  __shared__ ArgVal sh_av[8];
  __shared__ float sh_f[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (a.active && !a.active[b]) return;
  const float* lrow = a.logits + (long long)b * a.ld;
  int* hist = a.hist + (long long)b * a.hist_ld;
  const int n = a.hist_len[b];

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
    lg[t] = (s < 0.f) ? s * a.penalty : s / a.penalty;
  }
  __syncthreads();
  for (int i = tid; i < V; i += 256) lg[i] = lg[i] / a.temperature;
  __syncthreads();
  // k-th largest value counting duplicates: peel the maximum top_k times
  float kth = 0.f;
  for (int r = 0; r < a.top_k; ++r) {
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
    float pr = lg[i] / denom;
    float q = 1.f;
    if (!a.greedy) q = a.dbg_noise ? a.dbg_noise[(long long)b * V + i]
                                   : philox_normal(a.seed, (uint32_t)(b + a.utt_base), (uint32_t)n, (uint32_t)i);
    ArgVal c; c.v = pr / q; c.i = i;
    loc = better(loc, c);
  }
  const ArgVal tokv = block_argmax(loc, sh_av);
  if (tid == 0) {
    const int tok = tokv.i;
    hist[n] = tok;
    a.hist_len[b] = n + 1;
    if (a.advance_kv) a.kv_len[b] += 1;
    if (a.check_stop) {
      const bool stop = (raw_best.i == EOS) || (tok == EOS);
      if (stop && a.stop_step[b] < 0) {
        a.stop_step[b] = n;   // history length before this token (graph-replay safe step id)
        if (a.honour_stop) a.active[b] = 0;
      }
    }
  }
}

}  // namespace

void launch_sampler(const SamplerArgs& a, cudaStream_t s) {
  if (a.B <= 0) return;
  GENIE_CHECK(a.top_k >= 1 && a.top_k <= V, "sampler: bad top_k");
  launch_pdl(sampler_kernel, dim3(a.B), dim3(256), 0, s, a);
  GENIE_LAUNCHED("sampler");
}

}  // namespace genie
