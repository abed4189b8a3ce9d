// Persistent decode step for small batches (rows <= 8): stage_decoder#[12-1821] of ALL 24 layers + logits in
// ONE kernel launch.
//
// At batch 1 a decode step is 152 MB of fp16 weights and ~170 dependent kernels; each kernel boundary costs
// ~3.8 us however small the kernel, so the step was launch-latency bound (0.64 ms per token).  Here one CTA
// per SM stays resident for the whole step and the phases of a layer are separated by a device-wide barrier
// (one atomic arrive + acquire spin, bounded) instead of kernel boundaries:
//
//   P1  x = LN2(prev layer) (recomputed by every CTA, 512 floats per row)  ->  QKV GEMV        | barrier
//   P2  attention partials per (utterance, head, key chunk), cache append of the new k / v      | barrier
//   P3  combine partials (every CTA)  ->  out-proj GEMV + bias + residual                       | barrier
//   P4  LN1 (every CTA)               ->  FFN1 GEMV + bias + ReLU                               | barrier
//   P5  FFN2 GEMV + bias + residual                                                             | barrier
//
// GEMV = weight streaming: each warp owns output columns, 16-byte fp16 weight loads (all chunks of up to two
// columns in flight), fp32 FMA against the activation rows held in shared memory, warp-shuffle reduction.
// Everything another CTA wrote is read through L2 (ld.global.cg); weights and old cache rows through the
// read-only path.
#include "kernels.cuh"

#include <math_constants.h>

namespace genie {
namespace {

constexpr int D = 512, FF = 2048, NH = 16, PART_LD = 36;   // partial: [0] m, [1] l, [4..35] acc
constexpr int NTHR = 256, NWARP = NTHR / 32;

__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& target, unsigned G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += G;
    // arrive: release-ordered reduction without a return value (no round trip before the poll starts); the
    // bar.sync above makes the other threads' writes part of what this release publishes
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    for (unsigned spins = 0;; ++spins) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (spins > (1u << 22)) { atomicExch(ctr + 1, 1u); break; }   // never hang the GPU: flag and fall through
    }
  }
  __syncthreads();
}

// GEMV in two halves so that the weight (and bias) loads of phase p+1 are issued BEFORE the device-wide
// barrier that ends phase p: weights are constants, so their HBM latency hides behind the barrier wait.
// Warp gw owns column gw (and gw + nwarps when TWO); requires N <= (TWO ? 2 : 1) * nwarps.
template <int KC, bool TWO>
struct WRegs { uint4 u0[KC]; uint4 u1[TWO ? KC : 1]; float b0, b1; };

template <int KC, bool TWO>
__device__ __forceinline__ void gemv_prefetch(WRegs<KC, TWO>& w, const __half* __restrict__ W,
                                              const float* __restrict__ bias, int N, int gwarp, int nwarps, int lane) {
  constexpr int K = KC * 256;
  const int n0 = gwarp, n1 = gwarp + nwarps;
  w.b0 = 0.f; w.b1 = 0.f;
#pragma unroll
  for (int c = 0; c < KC; ++c) w.u0[c] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int c = 0; c < (TWO ? KC : 1); ++c) w.u1[c] = make_uint4(0u, 0u, 0u, 0u);
  if (n0 < N) {
#pragma unroll
    for (int c = 0; c < KC; ++c) w.u0[c] = __ldg(reinterpret_cast<const uint4*>(W + (long long)n0 * K + c * 256 + lane * 8));
    if (bias) w.b0 = __ldg(bias + n0);
  }
  if (TWO && n1 < N) {
#pragma unroll
    for (int c = 0; c < KC; ++c) w.u1[c] = __ldg(reinterpret_cast<const uint4*>(W + (long long)n1 * K + c * 256 + lane * 8));
    if (bias) w.b1 = __ldg(bias + n1);
  }
}

// out[r, n] = act(W[n, :] . xs[r, :] + bias[n] + res[r, n])
template <int BR, int KC, bool TWO>
__device__ __forceinline__ void gemv_compute(const WRegs<KC, TWO>& w, int N, const float* xs, float* out, int ldo,
                                             const float* res, bool relu, int B, int gwarp, int nwarps, int lane) {
  constexpr int K = KC * 256;
  const int n0 = gwarp, n1 = gwarp + nwarps;
  if (n0 >= N) return;                                      // warp-uniform
  const bool two = TWO && n1 < N;
  float r0 = 0.f, r1 = 0.f;                                 // residual (producer data, through L2), issued early
  if (res && lane < B) {
    r0 = __ldcg(res + (long long)lane * ldo + n0);
    if (two) r1 = __ldcg(res + (long long)lane * ldo + n1);
  }
  float acc0[BR], acc1[BR];
#pragma unroll
  for (int r = 0; r < BR; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    const __half2* h0 = reinterpret_cast<const __half2*>(&w.u0[c]);
    const __half2* h1 = reinterpret_cast<const __half2*>(&w.u1[TWO ? c : 0]);
    float wa[8], wb[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(h0[e]), fb = __half22float2(h1[e]);
      wa[2 * e] = fa.x; wa[2 * e + 1] = fa.y; wb[2 * e] = fb.x; wb[2 * e + 1] = fb.y;
    }
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      const float4 xa = *reinterpret_cast<const float4*>(xs + r * K + c * 256 + lane * 8);
      const float4 xb = *reinterpret_cast<const float4*>(xs + r * K + c * 256 + lane * 8 + 4);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc0[r] = fmaf(wa[e], xv[e], acc0[r]);
        if (TWO) acc1[r] = fmaf(wb[e], xv[e], acc1[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < BR; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc0[r] += __shfl_xor_sync(0xffffffffu, acc0[r], o);
      if (TWO) acc1[r] += __shfl_xor_sync(0xffffffffu, acc1[r], o);
    }
  }
#pragma unroll
  for (int r = 0; r < BR; ++r) {
    if (lane == r && r < B) {
      float v0 = acc0[r] + w.b0 + r0;
      if (relu) v0 = fmaxf(v0, 0.f);
      out[(long long)r * ldo + n0] = v0;
      if (two) {
        float v1 = acc1[r] + w.b1 + r1;
        if (relu) v1 = fmaxf(v1, 0.f);
        out[(long long)r * ldo + n1] = v1;
      }
    }
  }
}

// xs[r, :] = LN(src[r, :]) * g + b for r < B (one warp per row), zero rows above; optional copy to global
template <int BR>
__device__ __forceinline__ void layernorm_rows(float* xs, const float* src, const float* __restrict__ g,
                                               const float* __restrict__ b, float* gout, int B, int warp, int lane) {
  for (int r = warp; r < BR; r += NWARP) {
    float v[16];
    if (r < B) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcg(src + (long long)r * D + lane + 32 * i);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) sum += v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / (float)D;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; var = fmaf(d, d, var); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
      const float rstd = 1.f / sqrtf(var / (float)D + 1e-5f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = lane + 32 * i;
        v[i] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
        if (gout) gout[(long long)r * D + c] = v[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) xs[r * D + lane + 32 * i] = v[i];
  }
}

template <int BR>
__device__ __forceinline__ void load_rows(float* xs, const float* src, int K, int B) {
  for (int i = threadIdx.x; i < BR * (K / 4); i += NTHR) {
    const int r = i / (K / 4), c4 = i - r * (K / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < B) v = __ldcg(reinterpret_cast<const float4*>(src + (long long)r * K + c4 * 4));
    *reinterpret_cast<float4*>(xs + r * K + c4 * 4) = v;
  }
}

// one (utterance, head, key chunk): partial softmax statistics over the chunk's keys -> part[36].
// The first pass of cached K / V rows (written by earlier steps) is fetched BEFORE the barrier that publishes
// this step's q: att_prefetch / att_run.
// cache rows are floats or halves (PersistentStep::kv_f16): element offsets + a uniform branch on load / store —
// this kernel is latency-bound, the byte count of the cache does not matter here, only that it reads the same
// cache the batched path appends to
__device__ __forceinline__ float4 ld_kv4(const void* base, long long idx, int f16) {
  if (f16) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + idx));
    const __half2* hp = reinterpret_cast<const __half2*>(&u);
    const float2 a = __half22float2(hp[0]), b = __half22float2(hp[1]);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
}
__device__ __forceinline__ void st_kv4(void* base, long long idx, const float4& v, int f16) {
  if (f16) {
    uint2 u;
    __half2* hp = reinterpret_cast<__half2*>(&u);
    hp[0] = __floats2half2_rn(v.x, v.y); hp[1] = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(base) + idx) = u;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
  }
}

struct AttItem {
  int b, h, c, T, lo, hi; long long K, V;   // element offsets of this head's K / V rows in the cache
  float4 k4[2], v4[2]; bool valid, has_new;
  uint4 wq[4][2], wk[4][2], wv[4][2];      // this warp's 4 q (k, v) columns of head h: rows of Wqkv, 2 x 256 k
  float bq[4], bk[4], bv[4];
};

__device__ __forceinline__ void att_prefetch(AttItem& it, const PersistentStep& a, int layer) {
  const int item = blockIdx.x;
  it.valid = item < a.B * NH * a.nch;
  if (!it.valid) return;
  it.c = item % a.nch;
  const int bh = item / a.nch;
  it.h = bh % NH; it.b = bh / NH;
  if (a.active && !a.active[it.b]) { it.valid = false; return; }
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  it.T = a.kv_len[it.b];                                     // cached tokens; the new token is key index T
  const int total = it.T + 1;
  const int cs = (total + a.nch - 1) / a.nch;
  it.lo = it.c * cs; it.hi = min(it.lo + cs, total);
  it.K = (long long)it.b * a.utt_stride + (long long)layer * a.layer_stride + (long long)it.h * a.cap * 32;
  it.V = it.K + a.v_off;
  it.has_new = it.lo <= it.T && it.T < it.hi;                // this chunk holds the step's own token
  {
    // q (and, for the chunk with the new token, k / v) columns of this head: warp w owns columns 4w..4w+3
    const StepLayerPtrs L = a.layers[layer];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = it.h * 32 + warp * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        it.wq[j][c] = __ldg(reinterpret_cast<const uint4*>(L.wqkv + (long long)(col0 + j) * D + c * 256 + lane * 8));
        if (it.has_new) {
          it.wk[j][c] = __ldg(reinterpret_cast<const uint4*>(L.wqkv + (long long)(D + col0 + j) * D + c * 256 + lane * 8));
          it.wv[j][c] = __ldg(reinterpret_cast<const uint4*>(L.wqkv + (long long)(2 * D + col0 + j) * D + c * 256 + lane * 8));
        }
      }
      it.bq[j] = __ldg(L.bqkv + col0 + j);
      it.bk[j] = __ldg(L.bqkv + D + col0 + j);
      it.bv[j] = __ldg(L.bqkv + 2 * D + col0 + j);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = it.lo + u * 32 + grp;
    it.k4[u] = make_float4(0.f, 0.f, 0.f, 0.f); it.v4[u] = it.k4[u];
    if (j < it.hi && j < it.T) {
      it.k4[u] = ld_kv4(a.kv, it.K + (long long)j * 32 + sub * 4, a.kv_f16);
      it.v4[u] = ld_kv4(a.kv, it.V + (long long)j * 32 + sub * 4, a.kv_f16);
    }
  }
}

// dot of 4 weight rows (2 x 256-k chunks per lane) with the activation row x[512] in shared memory
__device__ __forceinline__ void dot4(const uint4 (&w)[4][2], const float (&bias)[4], const float* x, int lane, float* out4) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float4 xa = *reinterpret_cast<const float4*>(x + c * 256 + lane * 8);
    const float4 xb = *reinterpret_cast<const float4*>(x + c * 256 + lane * 8 + 4);
    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2* hp = reinterpret_cast<const __half2*>(&w[j][c]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(hp[e]);
        acc[j] = fmaf(f.x, xv[2 * e], acc[j]);
        acc[j] = fmaf(f.y, xv[2 * e + 1], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    if (lane == j) out4[j] = acc[j] + bias[j];
  }
}

// x: this utterance's layer-input row in shared memory; sqkv: [96] scratch for q | k_new | v_new of the head
__device__ __forceinline__ void att_run(AttItem& it, const PersistentStep& a, const float* x, float* sqkv, float* sred) {
  if (!it.valid) return;                                    // CTA-uniform
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 3, sub = tid & 7;                  // 32 groups of 8 lanes, one key per group per pass
  const int b = it.b, h = it.h, c = it.c, T = it.T, lo = it.lo, hi = it.hi;
  const long long K = it.K, V = it.V;
  dot4(it.wq, it.bq, x, lane, sqkv + warp * 4);
  if (it.has_new) {
    dot4(it.wk, it.bk, x, lane, sqkv + 32 + warp * 4);
    dot4(it.wv, it.bv, x, lane, sqkv + 64 + warp * 4);
  }
  __syncthreads();
  float4 q4 = *reinterpret_cast<const float4*>(sqkv + sub * 4);
  q4.x *= a.scale; q4.y *= a.scale; q4.z *= a.scale; q4.w *= a.scale;
  (void)b;

  float m = -CUDART_INF_F, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = lo; j0 < hi; j0 += 64) {                    // two keys per group in flight
    float4 k4[2], v4[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + u * 32 + grp;
      ok[u] = j < hi;
      k4[u] = make_float4(0.f, 0.f, 0.f, 0.f); v4[u] = k4[u];
      if (ok[u]) {
        if (j < T) {
          if (j0 == lo) { k4[u] = it.k4[u]; v4[u] = it.v4[u]; }
          else {
            k4[u] = ld_kv4(a.kv, K + (long long)j * 32 + sub * 4, a.kv_f16);
            v4[u] = ld_kv4(a.kv, V + (long long)j * 32 + sub * 4, a.kv_f16);
          }
        } else {                                            // this step's token: from the QKV rows, and into the cache
          k4[u] = *reinterpret_cast<const float4*>(sqkv + 32 + sub * 4);
          v4[u] = *reinterpret_cast<const float4*>(sqkv + 64 + sub * 4);
          if (T < a.cap) {
            st_kv4(a.kv, K + (long long)T * 32 + sub * 4, k4[u], a.kv_f16);
            st_kv4(a.kv, V + (long long)T * 32 + sub * 4, v4[u], a.kv_f16);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float s = q4.x * k4[u].x + q4.y * k4[u].y + q4.z * k4[u].z + q4.w * k4[u].w;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (ok[u]) {
        const float m_new = fmaxf(m, s);
        const float corr = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
        const float pj = expf(s - m_new);
        l = l * corr + pj;
        acc.x = acc.x * corr + pj * v4[u].x; acc.y = acc.y * corr + pj * v4[u].y;
        acc.z = acc.z * corr + pj * v4[u].z; acc.w = acc.w * corr + pj * v4[u].w;
        m = m_new;
      }
    }
  }
  // combine the 32 groups: sred = [32] m | [32] l | [32][32] acc
  float* sm_m = sred; float* sm_l = sred + 32; float* sm_acc = sred + 64;
  __syncthreads();
  if (sub == 0) { sm_m[grp] = m; sm_l[grp] = l; }
  sm_acc[grp * 32 + sub * 4 + 0] = acc.x; sm_acc[grp * 32 + sub * 4 + 1] = acc.y;
  sm_acc[grp * 32 + sub * 4 + 2] = acc.z; sm_acc[grp * 32 + sub * 4 + 3] = acc.w;
  __syncthreads();
  if (warp == 0) {
    float M = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 32; ++i) M = fmaxf(M, sm_m[i]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float wgt = (sm_m[i] == -CUDART_INF_F) ? 0.f : expf(sm_m[i] - M);
      num = fmaf(sm_acc[i * 32 + lane], wgt, num);
      den = fmaf(sm_l[i], wgt, den);
    }
    float* p = a.part + ((long long)(b * NH + h) * a.nch + c) * PART_LD;
    if (lane == 0) { p[0] = M; p[1] = den; }
    p[4 + lane] = num;
  }
}

template <int BR>
__global__ void __launch_bounds__(NTHR, 1) t2s_step_persistent_kernel(PersistentStep a) {
  extern __shared__ __align__(16) float xs[];               // [BR][2048] activation rows
  __shared__ float sred[64 + 32 * 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int gwarp = blockIdx.x * NWARP + warp, nwarps = G * NWARP;
  const int B = a.B;
  unsigned target = 0;

  __shared__ __align__(16) float sqkv[96];
  WRegs<2, true> wq;                                        // FFN1 / logits: K = 512, up to two columns
  AttItem it;
  att_prefetch(it, a, 0);
  for (int l = 0; l < a.n_layers; ++l) {
    const StepLayerPtrs L = a.layers[l];
    // ---- PA: layer-input row of this CTA's utterance (embedding for layer 0, LN2 of the previous layer
    // otherwise) -> its own q (k_new, v_new) columns -> attention partials over its key chunk (+ cache append).
    // CTAs (h = 0, c = 0) publish the row: it is the residual of the out-projection.
    if (it.valid) {
      float* pub = (it.h == 0 && it.c == 0) ? a.h + (long long)it.b * D : nullptr;
      if (l == 0) {
        for (int i = tid; i < D; i += NTHR) xs[i] = __ldcg(a.h + (long long)it.b * D + i);
      } else if (warp == 0) {
        const StepLayerPtrs P = a.layers[l - 1];
        layernorm_rows<1>(xs, a.lnin2 + (long long)it.b * D, P.ln2g, P.ln2b, pub, 1, 0, lane);
      }
    }
    __syncthreads();
    att_run(it, a, xs, sqkv, sred);
    WRegs<2, false> wo;
    gemv_prefetch<2, false>(wo, L.wout, L.bout, D, gwarp, nwarps, lane);
    grid_sync(a.sync, target, G);
    // ---- P3: combine chunks -> attention rows; out-proj + bias + residual
    for (int i = tid; i < BR * D; i += NTHR) {
      const int r = i / D, col = i - r * D;
      float o = 0.f;
      if (r < B) {
        const float* p = a.part + (long long)(r * NH + (col >> 5)) * a.nch * PART_LD;
        float mc[8], lc[8], ac[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          mc[c] = -CUDART_INF_F; lc[c] = 0.f; ac[c] = 0.f;
          if (c < a.nch) {
            mc[c] = __ldcg(p + c * PART_LD); lc[c] = __ldcg(p + c * PART_LD + 1);
            ac[c] = __ldcg(p + c * PART_LD + 4 + (col & 31));
          }
        }
        float M = -CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < 8; ++c) M = fmaxf(M, mc[c]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float wgt = (mc[c] == -CUDART_INF_F) ? 0.f : expf(mc[c] - M);
          num = fmaf(ac[c], wgt, num);
          den = fmaf(lc[c], wgt, den);
        }
        o = num / den;
      }
      xs[i] = o;
    }
    __syncthreads();
    gemv_compute<BR, 2, false>(wo, D, xs, a.lnin, D, a.h, false, B, gwarp, nwarps, lane);
    gemv_prefetch<2, true>(wq, L.wff1, L.bff1, FF, gwarp, nwarps, lane);
    grid_sync(a.sync, target, G);
    // ---- P4: LN1 -> FFN1 (+ ReLU)
    layernorm_rows<BR>(xs, a.lnin, L.ln1g, L.ln1b, blockIdx.x == 0 ? a.h1 : nullptr, B, warp, lane);
    __syncthreads();
    gemv_compute<BR, 2, true>(wq, FF, xs, a.ff, FF, nullptr, true, B, gwarp, nwarps, lane);
    WRegs<8, false> w2;
    gemv_prefetch<8, false>(w2, L.wff2, L.bff2, D, gwarp, nwarps, lane);
    grid_sync(a.sync, target, G);
    // ---- P5: FFN2 + bias + residual
    load_rows<BR>(xs, a.ff, FF, B);
    __syncthreads();
    gemv_compute<BR, 8, false>(w2, D, xs, a.lnin2, D, a.h1, false, B, gwarp, nwarps, lane);
    if (l + 1 < a.n_layers) att_prefetch(it, a, l + 1);
    else gemv_prefetch<2, true>(wq, a.wpredict, a.bpredict, a.vocab, gwarp, nwarps, lane);
    grid_sync(a.sync, target, G);
  }
  // ---- logits from LN2 of the last layer
  {
    const StepLayerPtrs P = a.layers[a.n_layers - 1];
    layernorm_rows<BR>(xs, a.lnin2, P.ln2g, P.ln2b, nullptr, B, warp, lane);
    __syncthreads();
    gemv_compute<BR, 2, true>(wq, a.vocab, xs, a.logits, a.ld_logits, nullptr, false, B, gwarp, nwarps, lane);
  }
}

// Cooperative launch: the driver either makes all `grid` CTAs co-resident (what the device-wide barrier needs)
// or fails the launch with cudaErrorCooperativeLaunchTooLarge — never a silent partial residency that would turn
// every barrier into a 4M-iteration spin.  Kernel nodes keep the attribute under stream capture.
template <int BR>
void launch_br(const PersistentStep& a, int grid, cudaStream_t s) {
  constexpr size_t smem = (size_t)BR * FF * sizeof(float);
  static DynSmemAttr attr;
  attr.ensure(t2s_step_persistent_kernel<BR>, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  GENIE_CUDA(cudaLaunchKernelEx(&cfg, t2s_step_persistent_kernel<BR>, a));
}

template <int BR>
int max_resident_ctas() {
  constexpr size_t smem = (size_t)BR * FF * sizeof(float);
  static DynSmemAttr attr;
  attr.ensure(t2s_step_persistent_kernel<BR>, smem);
  int per_sm = 0, dev = 0, sms = 0;
  GENIE_CUDA(cudaGetDevice(&dev));
  GENIE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GENIE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, t2s_step_persistent_kernel<BR>, NTHR, smem));
  return per_sm * sms;
}

}  // namespace

int persistent_step_chunks(int B, int grid) {
  int n = grid / (B * NH);
  return n < 1 ? 1 : (n > 8 ? 8 : n);
}

// can `grid` CTAs of the batch-B instantiation be resident at once on the current device (an otherwise idle one)?
bool persistent_step_fits(int B, int grid) {
  if (B < 1 || B > 8) return false;
  const int cap = B == 1 ? max_resident_ctas<1>() : B == 2 ? max_resident_ctas<2>() : B <= 4 ? max_resident_ctas<4>()
                                                                                             : max_resident_ctas<8>();
  return cap >= grid;
}

void launch_t2s_step_persistent(const PersistentStep& a, int grid, cudaStream_t s) {
  GENIE_CHECK(a.B >= 1 && a.B <= 8, "persistent step: batch must be 1..8");
  GENIE_CHECK(grid * NWARP >= a.vocab && 2 * grid * NWARP >= FF && grid >= a.B * NH * a.nch,
              "persistent step: grid too small for one column pair / one attention item per warp / CTA");
  GENIE_CUDA(cudaMemsetAsync(a.sync, 0, sizeof(unsigned), s));
  if (a.B == 1) launch_br<1>(a, grid, s);
  else if (a.B == 2) launch_br<2>(a, grid, s);
  else if (a.B <= 4) launch_br<4>(a, grid, s);
  else launch_br<8>(a, grid, s);
  GENIE_LAUNCHED("t2s_step_persistent");
}

}  // namespace genie
