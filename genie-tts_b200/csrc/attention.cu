// fp32 attention kernels.
//  * attention_kernel: tiled online-softmax attention over ragged segments, with
//    the T2S prefill mask (first_stage#[26-56]) or the VITS windowed
//    relative-position terms (vits#[363-755], restated as a band |i-j|<=w).
//  * decode_attention_kernel: one query per (utterance, head) streamed over the
//    head-major fp32 KV cache (stage#[63-96] without the per-step Concat).
#include "common.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace genie {
namespace {

constexpr int QT = 16;   // query rows per CTA (4 per warp)
constexpr int KT = 32;   // keys per tile (one per lane)

template <int D>
__global__ void __launch_bounds__(128) attention_kernel(Attn p) {
  constexpr int DC = D / 32;   // value columns per lane
  __shared__ float Qs[QT][D];
  __shared__ float Ks[KT][D + 1];
  __shared__ float Vs[KT][D];
  __shared__ float QRel[QT][16];
  __shared__ float RelV[9][D];

  const int b = blockIdx.z, h = blockIdx.y;
  const int qs = p.q_off ? p.q_off[b] : 0;
  const int Tq = p.q_off ? p.q_off[b + 1] - qs : p.max_q;
  const int ks = p.kv_off ? p.kv_off[b] : 0;
  const int Tk = p.kv_off ? p.kv_off[b + 1] - ks : p.max_q;
  const int q0 = blockIdx.x * QT;
  if (q0 >= Tq) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = p.window;
  const bool rel = p.rel_k != nullptr;
  const int lx = (p.mask_mode == 1) ? p.lx[b] : 0;

  for (int i = tid; i < QT * D; i += 128) {
    int r = i / D, e = i % D;
    float v = 0.f;
    if (q0 + r < Tq) v = p.q[(long long)(qs + q0 + r) * p.ldq + h * D + e] * p.scale;
    Qs[r][e] = v;
  }
  if (rel)
    for (int i = tid; i < (2 * W + 1) * D; i += 128) RelV[i / D][i % D] = p.rel_v[i];
  __syncthreads();
  if (rel) {
    // QRel[r][i] = (q_r * scale) . rel_k[i]
    for (int i = tid; i < QT * (2 * W + 1); i += 128) {
      int r = i / (2 * W + 1), c = i % (2 * W + 1);
      float s = 0.f;
      for (int e = 0; e < D; ++e) s = fmaf(Qs[r][e], p.rel_k[c * D + e], s);
      QRel[r][c] = s;
    }
  }
  __syncthreads();

  float m_run[4], l_run[4], acc[4][DC];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -CUDART_INF_F; l_run[r] = 0.f;
#pragma unroll
    for (int c = 0; c < DC; ++c) acc[r][c] = 0.f;
  }
  const int qr0 = warp * 4;   // this warp's first row within the CTA tile

  // keys beyond the last row this CTA can see are skipped for the causal mask
  int k_end = Tk;
  if (p.mask_mode == 1) {
    int last_q = min(q0 + QT, Tq) - 1;
    k_end = (last_q < lx) ? lx : last_q + 1;
  }
  for (int k0 = 0; k0 < k_end; k0 += KT) {
    for (int i = tid; i < KT * D; i += 128) {
      int j = i / D, e = i % D;
      float kv = 0.f, vv = 0.f;
      if (k0 + j < Tk) {
        long long row = (long long)(ks + k0 + j);
        kv = p.k[row * p.ldk + h * D + e];
        vv = p.v[row * p.ldv + h * D + e];
      }
      Ks[j][e] = kv; Vs[j][e] = vv;
    }
    __syncthreads();
    const int kj = k0 + lane;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int e = 0; e < D; ++e) {
      float kv = Ks[lane][e];
#pragma unroll
      for (int r = 0; r < 4; ++r) s[r] = fmaf(Qs[qr0 + r][e], kv, s[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = q0 + qr0 + r;
      bool ok = (kj < Tk) && (qi < Tq);
      if (p.mask_mode == 1) ok = ok && ((qi < lx) ? (kj < lx) : (kj <= qi));
      if (rel) {
        int idx = kj - qi + W;
        if (idx >= 0 && idx <= 2 * W) s[r] += QRel[qr0 + r][idx];
      }
      float sv = ok ? s[r] : -CUDART_INF_F;
      float tmax = sv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      float m_new = fmaxf(m_run[r], tmax);
      float pj = 0.f, corr = 1.f;
      if (m_new != -CUDART_INF_F) {
        pj = ok ? expf(sv - m_new) : 0.f;
        corr = (m_run[r] == -CUDART_INF_F) ? 0.f : expf(m_run[r] - m_new);
      }
      float psum = pj;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      l_run[r] = l_run[r] * corr + psum;
      m_run[r] = m_new;
#pragma unroll
      for (int c = 0; c < DC; ++c) acc[r][c] *= corr;
      s[r] = pj;
    }
    // P.V (+ relative-position values inside the band)
    for (int j = 0; j < KT; ++j) {
      float vv[DC];
#pragma unroll
      for (int c = 0; c < DC; ++c) vv[c] = Vs[j][lane + 32 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float pj = __shfl_sync(0xffffffffu, s[r], j);
#pragma unroll
        for (int c = 0; c < DC; ++c) acc[r][c] = fmaf(pj, vv[c], acc[r][c]);
        if (rel) {
          int idx = (k0 + j) - (q0 + qr0 + r) + W;
          if (idx >= 0 && idx <= 2 * W) {
#pragma unroll
            for (int c = 0; c < DC; ++c) acc[r][c] = fmaf(pj, RelV[idx][lane + 32 * c], acc[r][c]);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int qi = q0 + qr0 + r;
    if (qi >= Tq) continue;
    float inv = 1.f / l_run[r];
#pragma unroll
    for (int c = 0; c < DC; ++c)
      p.o[(long long)(qs + qi) * p.ldo + h * D + lane + 32 * c] = acc[r][c] * inv;
  }
}

// ---------------------------------------------------------------------------
// T2S prefill attention (head dim 32, no relative positions): register-tiled flash attention in fp32.
// CTA = 64 queries of one (utterance, head); 256 threads as 16 x 16: thread (ty, tx) owns query rows
// 4ty..4ty+3 and, per 64-key tile, keys 4tx..4tx+3 for the scores / value columns 2tx, 2tx+1 for P.V.
// Both products are 4x4 / 4x2 register tiles fed by 16-byte shared loads (16 FMA per two LDS.128 instead
// of 4 FMA per five scalar loads in the generic kernel above: 823 -> ~250 us per layer at 100 x 242 rows).
// Row statistics are reduced over the 16 tx lanes with shuffles; key tiles that the mask excludes for the
// whole CTA (text rows never see audio keys, audio rows are causal) are skipped.
// ---------------------------------------------------------------------------
constexpr int FQ = 64, FK = 64, FLD = 68;      // tile sizes; padded leading dimension (floats)

__global__ void __launch_bounds__(256) prefill_attention32_kernel(Attn p) {
  __shared__ __align__(16) float Qs[32][FLD];   // [d][q]  (pre-scaled)
  __shared__ __align__(16) float Ks[32][FLD];   // [d][key]
  __shared__ __align__(16) float Vs[FK][32];    // [key][d]
  __shared__ __align__(16) float Ps[FQ][FLD];   // [q][key]

  const int b = blockIdx.z, h = blockIdx.y;
  const int qs = p.q_off ? p.q_off[b] : 0;
  const int Tq = p.q_off ? p.q_off[b + 1] - qs : p.max_q;
  const int ks = p.kv_off ? p.kv_off[b] : 0;
  const int Tk = p.kv_off ? p.kv_off[b + 1] - ks : p.max_q;
  const int q0 = blockIdx.x * FQ;
  if (q0 >= Tq) return;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int lx = (p.mask_mode == 1) ? p.lx[b] : 0;

  // Q tile, transposed to [d][q]
  for (int i = tid; i < FQ * 8; i += 256) {
    const int r = i >> 3, c4 = i & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Tq) v = *reinterpret_cast<const float4*>(p.q + (long long)(qs + q0 + r) * p.ldq + h * 32 + c4 * 4);
    Qs[c4 * 4 + 0][r] = v.x * p.scale; Qs[c4 * 4 + 1][r] = v.y * p.scale;
    Qs[c4 * 4 + 2][r] = v.z * p.scale; Qs[c4 * 4 + 3][r] = v.w * p.scale;
  }

  float m_run[4], l_run[4], acc[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) { m_run[r] = -CUDART_INF_F; l_run[r] = 0.f; acc[r][0] = 0.f; acc[r][1] = 0.f; }

  int k_end = Tk;
  if (p.mask_mode == 1) {
    const int last_q = min(q0 + FQ, Tq) - 1;
    k_end = (last_q < lx) ? lx : last_q + 1;
  }
  for (int k0 = 0; k0 < k_end; k0 += FK) {
    __syncthreads();                                         // previous tile fully consumed (and Qs visible)
    for (int i = tid; i < FK * 8; i += 256) {
      const int j = i >> 3, c4 = i & 7;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + j < Tk) {
        const long long row = (long long)(ks + k0 + j);
        kv = *reinterpret_cast<const float4*>(p.k + row * p.ldk + h * 32 + c4 * 4);
        vv = *reinterpret_cast<const float4*>(p.v + row * p.ldv + h * 32 + c4 * 4);
      }
      Ks[c4 * 4 + 0][j] = kv.x; Ks[c4 * 4 + 1][j] = kv.y; Ks[c4 * 4 + 2][j] = kv.z; Ks[c4 * 4 + 3][j] = kv.w;
      *reinterpret_cast<float4*>(&Vs[j][c4 * 4]) = vv;
    }
    __syncthreads();
    // ---- scores: 4 queries x 4 keys per thread
    float sc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) sc[r][c] = 0.f;
#pragma unroll 8
    for (int e = 0; e < 32; ++e) {
      const float4 qv = *reinterpret_cast<const float4*>(&Qs[e][ty * 4]);
      const float4 kv = *reinterpret_cast<const float4*>(&Ks[e][tx * 4]);
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) sc[r][c] = fmaf(qa[r], ka[c], sc[r][c]);
    }
    // ---- mask, online softmax (row statistics over the 16 tx lanes), P -> shared
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int qi = q0 + ty * 4 + r;
      float tmax = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int kj = k0 + tx * 4 + c;
        bool ok = (kj < Tk) && (qi < Tq);
        if (p.mask_mode == 1) ok = ok && ((qi < lx) ? (kj < lx) : (kj <= qi));
        sc[r][c] = ok ? sc[r][c] : -CUDART_INF_F;
        tmax = fmaxf(tmax, sc[r][c]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float m_new = fmaxf(m_run[r], tmax);
      float corr = 1.f, psum = 0.f;
      float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m_new != -CUDART_INF_F) {
        corr = (m_run[r] == -CUDART_INF_F) ? 0.f : expf(m_run[r] - m_new);
        pv.x = sc[r][0] == -CUDART_INF_F ? 0.f : expf(sc[r][0] - m_new);
        pv.y = sc[r][1] == -CUDART_INF_F ? 0.f : expf(sc[r][1] - m_new);
        pv.z = sc[r][2] == -CUDART_INF_F ? 0.f : expf(sc[r][2] - m_new);
        pv.w = sc[r][3] == -CUDART_INF_F ? 0.f : expf(sc[r][3] - m_new);
        psum = (pv.x + pv.y) + (pv.z + pv.w);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      l_run[r] = l_run[r] * corr + psum;
      m_run[r] = m_new;
      acc[r][0] *= corr; acc[r][1] *= corr;
      *reinterpret_cast<float4*>(&Ps[ty * 4 + r][tx * 4]) = pv;
    }
    __syncthreads();
    // ---- O += P . V : 4 queries x 2 value columns per thread
#pragma unroll 4
    for (int j = 0; j < FK; j += 4) {
      float4 pr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) pr[r] = *reinterpret_cast<const float4*>(&Ps[ty * 4 + r][j]);
      const float2 v0 = *reinterpret_cast<const float2*>(&Vs[j + 0][tx * 2]);
      const float2 v1 = *reinterpret_cast<const float2*>(&Vs[j + 1][tx * 2]);
      const float2 v2 = *reinterpret_cast<const float2*>(&Vs[j + 2][tx * 2]);
      const float2 v3 = *reinterpret_cast<const float2*>(&Vs[j + 3][tx * 2]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][0] = fmaf(pr[r].x, v0.x, acc[r][0]); acc[r][1] = fmaf(pr[r].x, v0.y, acc[r][1]);
        acc[r][0] = fmaf(pr[r].y, v1.x, acc[r][0]); acc[r][1] = fmaf(pr[r].y, v1.y, acc[r][1]);
        acc[r][0] = fmaf(pr[r].z, v2.x, acc[r][0]); acc[r][1] = fmaf(pr[r].z, v2.y, acc[r][1]);
        acc[r][0] = fmaf(pr[r].w, v3.x, acc[r][0]); acc[r][1] = fmaf(pr[r].w, v3.y, acc[r][1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int qi = q0 + ty * 4 + r;
    if (qi >= Tq) continue;
    const float inv = 1.f / l_run[r];
    *reinterpret_cast<float2*>(p.o + (long long)(qs + qi) * p.ldo + h * 32 + tx * 2) =
        make_float2(acc[r][0] * inv, acc[r][1] * inv);
  }
}

// ---------------------------------------------------------------------------
// T2S prefill attention on the tensor cores (round 2).  Head dim 32 makes this a register-fragment problem, so it
// uses warp-level mma.sync.m16n8k16 (fp16 in, fp32 accumulate) rather than tcgen05: per (64-query, 64-key) tile a
// warp holds 16 query rows; S = Q.K^T is 8 n-tiles x 2 k-steps, the score fragments ARE the A fragments of P.V
// (4 n-tiles x 4 k-steps, V fragments through ldmatrix.trans), online softmax in fp32 on the fragments
// (FlashAttention-2 register layout).  Precision: every operand enters as a hi + lo fp16 pair and each product is
// three MMAs (hi.hi + lo.hi + hi.lo), i.e. fp32-input accuracy up to the dropped lo.lo term (~2^-22 relative):
// a first version with K / V rounded to fp16 (two MMAs) cost 2 of the 100 bench sentences their greedy-token
// identity (near-ties of 3e-5 / 7e-5 on the flat fixture), this one keeps the prefill logits where the fp32 kernel
// has them.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_h2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// B fragment (k16 x n8) of a row-major [k][n] shared tile: two transposed 8x8 matrices
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& b0, uint32_t& b1, const __half* row_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
}

constexpr int MK = 64, KLD = 40;                      // key tile; padded smem leading dimension (halves)

// NW warps x 16 query rows per CTA (128 rows at NW = 8: each K / V tile is loaded and split once per 128 queries);
// the next key tile's fp32 rows are fetched into registers while the current one is multiplied.
template <int NW>
__global__ void __launch_bounds__(NW * 32) prefill_attention_mma_kernel(Attn p) {
  constexpr int MQ = NW * 16, NT = NW * 32, LPT = MK * 8 / NT;   // float4 per thread and tile (K and V each)
  __shared__ __align__(16) __half Kh[MK][KLD], Kl[MK][KLD];     // [key][d], hi / lo
  __shared__ __align__(16) __half Vh[MK][KLD], Vl[MK][KLD];     // [key][d], hi / lo
  const int b = blockIdx.z, h = blockIdx.y;
  const int qs = p.q_off ? p.q_off[b] : 0;
  const int Tq = p.q_off ? p.q_off[b + 1] - qs : p.max_q;
  const int ks = p.kv_off ? p.kv_off[b] : 0;
  const int Tk = p.kv_off ? p.kv_off[b + 1] - ks : p.max_q;
  const int q0 = blockIdx.x * MQ;
  if (q0 >= Tq) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int lx = (p.mask_mode == 1) ? p.lx[b] : 0;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;     // this thread's two query rows (within the utterance)

  // Q fragments (pre-scaled, hi / lo): k-step kk covers d = 16kk .. 16kk+15
  uint32_t qh[2][4], ql[2][4];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int half = 0; half < 2; ++half) {             // columns 2t (+8)
      const int d = kk * 16 + half * 8 + 2 * t;
      float2 a = make_float2(0.f, 0.f), c = a;
      if (r0 < Tq) a = *reinterpret_cast<const float2*>(p.q + (long long)(qs + r0) * p.ldq + h * 32 + d);
      if (r1 < Tq) c = *reinterpret_cast<const float2*>(p.q + (long long)(qs + r1) * p.ldq + h * 32 + d);
      split_h2(a.x * p.scale, a.y * p.scale, qh[kk][half * 2 + 0], ql[kk][half * 2 + 0]);
      split_h2(c.x * p.scale, c.y * p.scale, qh[kk][half * 2 + 1], ql[kk][half * 2 + 1]);
    }

  float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
  float o[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { o[j][0] = 0.f; o[j][1] = 0.f; o[j][2] = 0.f; o[j][3] = 0.f; }

  int k_end = Tk;
  if (p.mask_mode == 1) {
    const int last_q = min(q0 + MQ, Tq) - 1;
    k_end = (last_q < lx) ? lx : last_q + 1;           // text rows never see audio keys; audio rows are causal
  }
  float4 kreg[LPT], vreg[LPT];
  auto fetch = [&](int k0) {                           // raw fp32 rows of one key tile -> registers
#pragma unroll
    for (int u = 0; u < LPT; ++u) {
      const int i = tid + u * NT, r = i >> 3, c4 = i & 7;
      kreg[u] = make_float4(0.f, 0.f, 0.f, 0.f); vreg[u] = kreg[u];
      if (k0 + r < Tk) {
        kreg[u] = *reinterpret_cast<const float4*>(p.k + (long long)(ks + k0 + r) * p.ldk + h * 32 + c4 * 4);
        vreg[u] = *reinterpret_cast<const float4*>(p.v + (long long)(ks + k0 + r) * p.ldv + h * 32 + c4 * 4);
      }
    }
  };
  if (k_end > 0) fetch(0);
  for (int k0 = 0; k0 < k_end; k0 += MK) {
    __syncthreads();                                   // previous tile fully consumed
#pragma unroll
    for (int u = 0; u < LPT; ++u) {                    // K / V tile: fp32 -> fp16 hi / lo
      const int i = tid + u * NT, r = i >> 3, c4 = i & 7;
      uint2 hi, lo;
      split_h2(kreg[u].x, kreg[u].y, hi.x, lo.x); split_h2(kreg[u].z, kreg[u].w, hi.y, lo.y);
      *reinterpret_cast<uint2*>(&Kh[r][c4 * 4]) = hi; *reinterpret_cast<uint2*>(&Kl[r][c4 * 4]) = lo;
      split_h2(vreg[u].x, vreg[u].y, hi.x, lo.x); split_h2(vreg[u].z, vreg[u].w, hi.y, lo.y);
      *reinterpret_cast<uint2*>(&Vh[r][c4 * 4]) = hi; *reinterpret_cast<uint2*>(&Vl[r][c4 * 4]) = lo;
    }
    __syncthreads();
    if (k0 + MK < k_end) fetch(k0 + MK);               // in flight while this tile is multiplied

    // ---- S = Q K^T for this warp's 16 rows x 64 keys
    float sc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j][0] = 0.f; sc[j][1] = 0.f; sc[j][2] = 0.f; sc[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t h0 = *reinterpret_cast<const uint32_t*>(&Kh[j * 8 + g][kk * 16 + 2 * t]);
        const uint32_t h1 = *reinterpret_cast<const uint32_t*>(&Kh[j * 8 + g][kk * 16 + 8 + 2 * t]);
        const uint32_t e0 = *reinterpret_cast<const uint32_t*>(&Kl[j * 8 + g][kk * 16 + 2 * t]);
        const uint32_t e1 = *reinterpret_cast<const uint32_t*>(&Kl[j * 8 + g][kk * 16 + 8 + 2 * t]);
        mma16816(sc[j], qh[kk], h0, h1);
        mma16816(sc[j], ql[kk], h0, h1);
        mma16816(sc[j], qh[kk], e0, e1);
      }
    }
    // ---- mask (first_stage#[26-56]) + online softmax on the fragments: c0,c1 -> row r0, c2,c3 -> row r1
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + j * 8 + 2 * t + (e & 1);
        const int row = (e < 2) ? r0 : r1;
        bool ok = key < Tk && row < Tq;
        if (ok && p.mask_mode == 1) ok = (row < lx) ? (key < lx) : (key < lx || key <= row);
        if (!ok) sc[j][e] = -CUDART_INF_F;
        if (e < 2) mx0 = fmaxf(mx0, sc[j][e]); else mx1 = fmaxf(mx1, sc[j][e]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = (m0 == -CUDART_INF_F) ? 0.f : expf(m0 - mx0);
    const float c1 = (m1 == -CUDART_INF_F) ? 0.f : expf(m1 - mx1);
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float mref = (e < 2) ? mx0 : mx1;
        const float pv = (sc[j][e] == -CUDART_INF_F || mref == -CUDART_INF_F) ? 0.f : expf(sc[j][e] - mref);
        sc[j][e] = pv;
        if (e < 2) s0 += pv; else s1 += pv;
      }
    l0 += s0; l1 += s1;                                 // quad-partial sums; reduced once at the end
    m0 = mx0; m1 = mx1;
    // ---- O += P V : k-step kk = keys 16kk .. 16kk+15 = score n-tiles 2kk, 2kk+1
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph[4], pl[4];
      split_h2(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl[0]);
      split_h2(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl[1]);
      split_h2(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl[2]);
      split_h2(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {                    // value columns 8j .. 8j+7
        uint32_t h0, h1, e0, e1;
        ldsm_x2_trans(h0, h1, &Vh[kk * 16 + (lane & 15)][j * 8]);
        ldsm_x2_trans(e0, e1, &Vl[kk * 16 + (lane & 15)][j * 8]);
        mma16816(o[j], ph, h0, h1);
        mma16816(o[j], pl, h0, h1);
        mma16816(o[j], ph, e0, e1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (r0 < Tq)
      *reinterpret_cast<float2*>(p.o + (long long)(qs + r0) * p.ldo + h * 32 + j * 8 + 2 * t) = make_float2(o[j][0] * i0, o[j][1] * i0);
    if (r1 < Tq)
      *reinterpret_cast<float2*>(p.o + (long long)(qs + r1) * p.ldo + h * 32 + j * 8 + 2 * t) = make_float2(o[j][2] * i1, o[j][3] * i1);
  }
}

// ---------------------------------------------------------------------------
// Decode attention: grid (H=16, B); 128 threads.  8 lanes x float4 cover one
// 32-float key row, so each warp-load touches 4 keys = 512 contiguous bytes.
// ---------------------------------------------------------------------------
// FUSED: q / k_new / v_new are still split-K partials of the QKV GEMM ([nsplit][B][1536], bias deferred):
// the CTA finishes its own 3 x 32 columns, appends k_new / v_new to the cache at position kv_len[b] and
// attends over the kv_len[b] cached tokens plus the new one (no separate finish / cache-scatter kernel).
// Four tokens per thread group are loaded per round (8 independent 16-byte loads per thread in flight):
// the per-CTA chain is ceil(T / 64) HBM round trips instead of ceil(T / 16).
template <bool FUSED>
__global__ void __launch_bounds__(128) decode_attention_kernel(
    const float* __restrict__ q, int nsplit, long long split_stride, const float* __restrict__ bias,
    float* __restrict__ o, float* __restrict__ kv_base, long long utt_stride, long long layer_off, long long v_off,
    const int* __restrict__ kv_len, const int* __restrict__ active, int cap, float scale, int t_add, int ldq) {
  const int h = blockIdx.x, b = blockIdx.y;
  if (FUSED) pdl_trigger();
  // kv_len / active are written by the sampler, the cache rows < kv_len by earlier steps: all complete before
  // this step's first kernel started, so they may be read ahead of pdl_wait (only the QKV partials may not)
  const int kvl = kv_len[b];                                // both loads in flight before the branch: the start of
  const int act = active ? active[b] : 1;                   // this kernel is a chain of dependent round trips
  if (!act) return;
  const int T = kvl + (FUSED ? 0 : t_add);                  // cached tokens to stream
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, sub = lane & 7;
  float* K = kv_base + (long long)b * utt_stride + layer_off + (long long)h * cap * 32;
  float* V = K + v_off;
  const int jt = warp * 4 + grp;                            // this thread group's token within a 16-token slab

  auto load_batch = [&](int it, float4 (&k4)[4], float4 (&v4)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = it * 64 + u * 16 + jt;
      k4[u] = make_float4(0.f, 0.f, 0.f, 0.f); v4[u] = k4[u];
      if (j < T) {
        k4[u] = __ldg(reinterpret_cast<const float4*>(K + (long long)j * 32 + sub * 4));
        v4[u] = __ldg(reinterpret_cast<const float4*>(V + (long long)j * 32 + sub * 4));
      }
    }
  };
  const int iters = (T + 63) / 64;
  float4 k4[4], v4[4], kx[4], vx[4];
  load_batch(0, k4, v4);                                    // in flight across the wait below

  float4 q4, kn = make_float4(0.f, 0.f, 0.f, 0.f), vn = kn;
  if (FUSED) {
    const float* pq = q + (long long)b * ldq + h * 32 + sub * 4;
    q4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) {                                              // null: the producer already added it
      q4 = __ldg(reinterpret_cast<const float4*>(bias + h * 32 + sub * 4));
      kn = __ldg(reinterpret_cast<const float4*>(bias + 512 + h * 32 + sub * 4));
      vn = __ldg(reinterpret_cast<const float4*>(bias + 1024 + h * 32 + sub * 4));
    }
    pdl_wait();
    for (int sp = 0; sp < nsplit; ++sp) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride));          // producer data:
      const float4 c = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + 512));    // via L2, see
      const float4 d = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + 1024));   // common.cuh (PDL)
      q4.x += a.x; q4.y += a.y; q4.z += a.z; q4.w += a.w;
      kn.x += c.x; kn.y += c.y; kn.z += c.z; kn.w += c.w;
      vn.x += d.x; vn.y += d.y; vn.z += d.z; vn.w += d.w;
    }
    if (warp == 0 && grp == 0 && T < cap) {
      *reinterpret_cast<float4*>(K + (long long)T * 32 + sub * 4) = kn;
      *reinterpret_cast<float4*>(V + (long long)T * 32 + sub * 4) = vn;
    }
  } else {
    q4 = *reinterpret_cast<const float4*>(q + (long long)b * ldq + h * 32 + sub * 4);
  }
  q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;

  float m = -CUDART_INF_F, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
    if (it + 1 < iters) load_batch(it + 1, kx, vx);         // next round in flight while this one is reduced
    const int j0 = it * 64 + jt;
    float sc[4];
    float m_new = m;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float t = q4.x * k4[u].x + q4.y * k4[u].y + q4.z * k4[u].z + q4.w * k4[u].w;
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      sc[u] = (j0 + u * 16 < T) ? t : -CUDART_INF_F;
      m_new = fmaxf(m_new, sc[u]);
    }
    if (m_new != -CUDART_INF_F) {
      const float c = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
      l *= c; acc.x *= c; acc.y *= c; acc.z *= c; acc.w *= c;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float pj = (sc[u] == -CUDART_INF_F) ? 0.f : expf(sc[u] - m_new);
        l += pj;
        acc.x = fmaf(pj, v4[u].x, acc.x); acc.y = fmaf(pj, v4[u].y, acc.y);
        acc.z = fmaf(pj, v4[u].z, acc.z); acc.w = fmaf(pj, v4[u].w, acc.w);
      }
      m = m_new;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { k4[u] = kx[u]; v4[u] = vx[u]; }
  }
  if (FUSED && warp == 0 && grp == 0) {                      // the token of this step
    float t = q4.x * kn.x + q4.y * kn.y + q4.z * kn.z + q4.w * kn.w;
    t += __shfl_xor_sync(0x000000ffu, t, 1);
    t += __shfl_xor_sync(0x000000ffu, t, 2);
    t += __shfl_xor_sync(0x000000ffu, t, 4);
    const float m_new = fmaxf(m, t);
    const float c = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
    const float pj = expf(t - m_new);
    l = l * c + pj;
    acc.x = acc.x * c + pj * vn.x; acc.y = acc.y * c + pj * vn.y;
    acc.z = acc.z * c + pj * vn.z; acc.w = acc.w * c + pj * vn.w;
    m = m_new;
  }
  __shared__ float sm_m[16], sm_l[16], sm_acc[16][32];
  const int g = warp * 4 + grp;
  if (sub == 0) { sm_m[g] = m; sm_l[g] = l; }
  sm_acc[g][sub * 4 + 0] = acc.x; sm_acc[g][sub * 4 + 1] = acc.y;
  sm_acc[g][sub * 4 + 2] = acc.z; sm_acc[g][sub * 4 + 3] = acc.w;
  __syncthreads();
  if (warp == 0) {
    float M = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 16; ++i) M = fmaxf(M, sm_m[i]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float wgt = (sm_m[i] == -CUDART_INF_F) ? 0.f : expf(sm_m[i] - M);
      num = fmaf(sm_acc[i][lane], wgt, num);
      den = fmaf(sm_l[i], wgt, den);
    }
    o[(long long)b * 512 + h * 32 + lane] = num / den;
  }
}


// ---------------------------------------------------------------------------
// The same kernel over an fp16 KV cache (option kv_fp16, the default for batched decode: measured token parity in
// DESIGN.md §2).  A cached row is 32 halves = 64 bytes, so 4 lanes x 16 bytes cover one key and a warp-load touches
// 8 keys = 512 contiguous bytes; 128 threads cover 32 keys per pass, 4 passes (128 keys) in flight per round.
// q, the new token's k / v, the scores and the accumulators stay fp32; K / V are widened on load, the new token
// is rounded to fp16 only when it is appended (this step still attends to its fp32 value).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void h8_to_f8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
}
__device__ __forceinline__ uint4 f8_to_h8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  return u;
}

template <bool FUSED>
__global__ void __launch_bounds__(128) decode_attention16_kernel(
    const float* __restrict__ q, int nsplit, long long split_stride, const float* __restrict__ bias,
    float* __restrict__ o, __half* __restrict__ kv_base, long long utt_stride, long long layer_off, long long v_off,
    const int* __restrict__ kv_len, const int* __restrict__ active, int cap, float scale, int t_add, int ldq) {
  const int h = blockIdx.x, b = blockIdx.y;
  if (FUSED) pdl_trigger();
  const int kvl = kv_len[b];
  const int act = active ? active[b] : 1;
  if (!act) return;
  const int T = kvl + (FUSED ? 0 : t_add);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, sub = lane & 3;                // 8 keys per warp-load, 4 lanes x 8 halves per key
  __half* K = kv_base + (long long)b * utt_stride + layer_off + (long long)h * cap * 32;
  __half* V = K + v_off;
  const int jt = warp * 8 + grp;                            // this thread group's key within a 32-key slab

  auto load_batch = [&](int it, uint4 (&k8)[4], uint4 (&v8)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = it * 128 + u * 32 + jt;
      k8[u] = make_uint4(0u, 0u, 0u, 0u); v8[u] = k8[u];
      if (j < T) {
        k8[u] = __ldg(reinterpret_cast<const uint4*>(K + (long long)j * 32 + sub * 8));
        v8[u] = __ldg(reinterpret_cast<const uint4*>(V + (long long)j * 32 + sub * 8));
      }
    }
  };
  const int iters = (T + 127) / 128;
  uint4 k8[4], v8[4], kx[4], vx[4];
  load_batch(0, k8, v8);                                    // in flight across the wait below

  float q8[8], kn[8], vn[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { q8[e] = 0.f; kn[e] = 0.f; vn[e] = 0.f; }
  if (FUSED) {
    const float* pq = q + (long long)b * ldq + h * 32 + sub * 8;
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(bias + h * 32 + sub * 8 + e));
        const float4 c = __ldg(reinterpret_cast<const float4*>(bias + 512 + h * 32 + sub * 8 + e));
        const float4 d = __ldg(reinterpret_cast<const float4*>(bias + 1024 + h * 32 + sub * 8 + e));
        q8[e] = a.x; q8[e + 1] = a.y; q8[e + 2] = a.z; q8[e + 3] = a.w;
        kn[e] = c.x; kn[e + 1] = c.y; kn[e + 2] = c.z; kn[e + 3] = c.w;
        vn[e] = d.x; vn[e + 1] = d.y; vn[e + 2] = d.z; vn[e + 3] = d.w;
      }
    }
    pdl_wait();
    for (int sp = 0; sp < nsplit; ++sp) {
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + e));          // producer data:
        const float4 c = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + 512 + e));    // via L2 (PDL)
        const float4 d = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + 1024 + e));
        q8[e] += a.x; q8[e + 1] += a.y; q8[e + 2] += a.z; q8[e + 3] += a.w;
        kn[e] += c.x; kn[e + 1] += c.y; kn[e + 2] += c.z; kn[e + 3] += c.w;
        vn[e] += d.x; vn[e + 1] += d.y; vn[e + 2] += d.z; vn[e + 3] += d.w;
      }
    }
    if (warp == 0 && grp == 0 && T < cap) {
      *reinterpret_cast<uint4*>(K + (long long)T * 32 + sub * 8) = f8_to_h8(kn);
      *reinterpret_cast<uint4*>(V + (long long)T * 32 + sub * 8) = f8_to_h8(vn);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) q8[e] = q[(long long)b * ldq + h * 32 + sub * 8 + e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) q8[e] *= scale;

  float m = -CUDART_INF_F, l = 0.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (it + 1 < iters) load_batch(it + 1, kx, vx);         // next round in flight while this one is reduced
    const int j0 = it * 128 + jt;
    float sc[4];
    float m_new = m;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float kf[8];
      h8_to_f8(k8[u], kf);
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) t = fmaf(q8[e], kf[e], t);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      sc[u] = (j0 + u * 32 < T) ? t : -CUDART_INF_F;
      m_new = fmaxf(m_new, sc[u]);
    }
    if (m_new != -CUDART_INF_F) {
      const float c = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
      l *= c;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= c;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float pj = (sc[u] == -CUDART_INF_F) ? 0.f : expf(sc[u] - m_new);
        l += pj;
        float vf[8];
        h8_to_f8(v8[u], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e]);
      }
      m = m_new;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { k8[u] = kx[u]; v8[u] = vx[u]; }
  }
  if (FUSED && warp == 0 && grp == 0) {                      // the token of this step, in fp32
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) t = fmaf(q8[e], kn[e], t);
    t += __shfl_xor_sync(0x0000000fu, t, 1);
    t += __shfl_xor_sync(0x0000000fu, t, 2);
    const float m_new = fmaxf(m, t);
    const float c = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
    const float pj = expf(t - m_new);
    l = l * c + pj;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = acc[e] * c + pj * vn[e];
    m = m_new;
  }
  __shared__ float sm_m[32], sm_l[32], sm_acc[32][33];
  const int g = warp * 8 + grp;
  if (sub == 0) { sm_m[g] = m; sm_l[g] = l; }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm_acc[g][sub * 8 + e] = acc[e];
  __syncthreads();
  if (warp == 0) {
    float M = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 32; ++i) M = fmaxf(M, sm_m[i]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float wgt = (sm_m[i] == -CUDART_INF_F) ? 0.f : expf(sm_m[i] - M);
      num = fmaf(sm_acc[i][lane], wgt, num);
      den = fmaf(sm_l[i], wgt, den);
    }
    o[(long long)b * 512 + h * 32 + lane] = num / den;
  }
}


// ---------------------------------------------------------------------------
// Bulk-copy variant of the fp16-cache decode attention (the default): the K and V rows of one (utterance, head)
// are CONTIGUOUS in the head-major cache, so ONE thread streams them into a shared-memory ring with cp.async.bulk
// (128-key chunks of 8 KB + 8 KB, mbarrier expect_tx / complete_tx); all chunks of a ~20-character sentence
// (<= 384 cached tokens) are in flight before the CTA has even read q.  In-flight bytes are bounded by shared memory
// (48 KB per CTA, 4 CTAs per SM) instead of registers: the register-staged kernel above keeps ~8 KB per CTA in
// flight and a chain of three dependent load rounds per CTA, this one a single round trip for the whole slab.
// The arithmetic (and its order) is the register kernel's: 4 lanes x 8 halves per key, 8 keys per warp pass.
// ---------------------------------------------------------------------------
// Two shapes.  LONG caches (slab >= 640 rows): 128 threads, 128-key stages, 3 stages (48 KB, 4 CTAs per SM).
// SHORT caches (the ~20-character sentences of the headline workload, ~290 cached rows): 64 threads, 64-key stages,
// 2 stages (16 KB) so that 11-12 CTAs fit an SM and ALL 16 x B CTAs of a 100-utterance batch are resident in ONE
// wave — with 4-5 resident CTAs per SM the register-staged kernel needs three waves, the last one nearly empty.
template <int NW, int CH, int NST>
constexpr size_t att_bulk_smem() {
  return (size_t)NST * 2 * CH * 64 + 64 + (size_t)(NW * 8 * 2 + NW * 8 * 33) * sizeof(float) + 128;
}

__device__ __forceinline__ uint32_t att_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool FUSED, int NW, int CH, int NST>
__global__ void __launch_bounds__(NW * 32, NW == 2 ? 12 : 4) decode_attention16_bulk_kernel(
    const float* __restrict__ q, int nsplit, long long split_stride, const float* __restrict__ bias,
    float* __restrict__ o, __half* __restrict__ kv_base, long long utt_stride, long long layer_off, long long v_off,
    const int* __restrict__ kv_len, const int* __restrict__ active, int cap, float scale, int t_add, int ldq) {
  extern __shared__ uint8_t att_smem_raw[];
  const int h = blockIdx.x, b = blockIdx.y;
  if (FUSED) pdl_trigger();
  const int kvl = kv_len[b];
  const int act = active ? active[b] : 1;
  if (!act) return;
  const int T = kvl + (FUSED ? 0 : t_add);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, sub = lane & 3;
  __half* K = kv_base + (long long)b * utt_stride + layer_off + (long long)h * cap * 32;
  __half* V = K + v_off;
  constexpr int ACH = CH, AST = NST, NG = NW * 8;             // NG thread groups of 4 lanes: one key each per pass
  constexpr uint32_t ASTAGE = 2u * ACH * 64u;                 // K chunk + V chunk
  uint8_t* ring = att_smem_raw + ((128u - (att_smem_u32(att_smem_raw) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)AST * ASTAGE);
  float* sm_m = reinterpret_cast<float*>(ring + (size_t)AST * ASTAGE + 64);
  float* sm_l = sm_m + NG;
  float* sm_acc = sm_l + NG;                                  // [NG][33]
  const int nch = (T + ACH - 1) / ACH;

  auto issue = [&](int c) {                                   // thread 0: chunk c -> stage c % AST
    const int rows = min(ACH, T - c * ACH);
    const uint32_t bytes = (uint32_t)rows * 64u;
    uint64_t* bar = &full[c % AST];
    uint8_t* dst = ring + (size_t)(c % AST) * ASTAGE;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(att_smem_u32(bar)), "r"(2u * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(att_smem_u32(dst)), "l"(K + (long long)c * ACH * 32), "r"(bytes), "r"(att_smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(att_smem_u32(dst + ACH * 64)), "l"(V + (long long)c * ACH * 32), "r"(bytes), "r"(att_smem_u32(bar)) : "memory");
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < AST; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(att_smem_u32(&full[s])), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // cached rows were written by earlier steps / the prefill: complete before this step's first kernel started,
    // so the whole slab may be requested ahead of the PDL wait
    for (int c = 0; c < nch && c < AST; ++c) issue(c);
  }
  __syncthreads();

  float q8[8], kn[8], vn[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { q8[e] = 0.f; kn[e] = 0.f; vn[e] = 0.f; }
  if (FUSED) {
    const float* pq = q + (long long)b * ldq + h * 32 + sub * 8;
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(bias + h * 32 + sub * 8 + e));
        const float4 c = __ldg(reinterpret_cast<const float4*>(bias + 512 + h * 32 + sub * 8 + e));
        const float4 d = __ldg(reinterpret_cast<const float4*>(bias + 1024 + h * 32 + sub * 8 + e));
        q8[e] = a.x; q8[e + 1] = a.y; q8[e + 2] = a.z; q8[e + 3] = a.w;
        kn[e] = c.x; kn[e + 1] = c.y; kn[e + 2] = c.z; kn[e + 3] = c.w;
        vn[e] = d.x; vn[e + 1] = d.y; vn[e + 2] = d.z; vn[e + 3] = d.w;
      }
    }
    pdl_wait();
    for (int sp = 0; sp < nsplit; ++sp) {
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + e));
        const float4 c = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + 512 + e));
        const float4 d = __ldcg(reinterpret_cast<const float4*>(pq + sp * split_stride + 1024 + e));
        q8[e] += a.x; q8[e + 1] += a.y; q8[e + 2] += a.z; q8[e + 3] += a.w;
        kn[e] += c.x; kn[e + 1] += c.y; kn[e + 2] += c.z; kn[e + 3] += c.w;
        vn[e] += d.x; vn[e + 1] += d.y; vn[e + 2] += d.z; vn[e + 3] += d.w;
      }
    }
    if (warp == 0 && grp == 0 && T < cap) {
      *reinterpret_cast<uint4*>(K + (long long)T * 32 + sub * 8) = f8_to_h8(kn);
      *reinterpret_cast<uint4*>(V + (long long)T * 32 + sub * 8) = f8_to_h8(vn);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) q8[e] = q[(long long)b * ldq + h * 32 + sub * 8 + e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) q8[e] *= scale;

  float m = -CUDART_INF_F, l = 0.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int st = c % AST;
    const uint32_t parity = (uint32_t)((c / AST) & 1);
    {
      uint32_t done = 0;
      for (uint32_t i = 0; i < (1u << 24) && !done; ++i)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(att_smem_u32(&full[st])), "r"(parity) : "memory");
    }
    const uint8_t* kc = ring + (size_t)st * ASTAGE;
    const uint8_t* vc = kc + ACH * 64;
    const int rows = min(ACH, T - c * ACH);
    constexpr int NU = ACH / NG;                              // passes per stage
    float sc[NU];
    uint4 v8[NU];
    float m_new = m;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int j = u * NG + warp * 8 + grp;
      uint4 k8 = make_uint4(0u, 0u, 0u, 0u);
      v8[u] = k8;
      if (j < rows) {
        k8 = *reinterpret_cast<const uint4*>(kc + j * 64 + sub * 16);
        v8[u] = *reinterpret_cast<const uint4*>(vc + j * 64 + sub * 16);
      }
      float kf[8];
      h8_to_f8(k8, kf);
      float t = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) t = fmaf(q8[e], kf[e], t);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      sc[u] = (j < rows) ? t : -CUDART_INF_F;
      m_new = fmaxf(m_new, sc[u]);
    }
    if (m_new != -CUDART_INF_F) {
      const float cf = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
      l *= cf;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= cf;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const float pj = (sc[u] == -CUDART_INF_F) ? 0.f : expf(sc[u] - m_new);
        l += pj;
        float vf[8];
        h8_to_f8(v8[u], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vf[e], acc[e]);
      }
      m = m_new;
    }
    if (c + AST < nch) {                                      // long caches: recycle the stage
      __syncthreads();
      if (tid == 0) issue(c + AST);
    }
  }
  if (FUSED && warp == 0 && grp == 0) {                      // the token of this step, in fp32
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) t = fmaf(q8[e], kn[e], t);
    t += __shfl_xor_sync(0x0000000fu, t, 1);
    t += __shfl_xor_sync(0x0000000fu, t, 2);
    const float m_new = fmaxf(m, t);
    const float cf = (m == -CUDART_INF_F) ? 0.f : expf(m - m_new);
    const float pj = expf(t - m_new);
    l = l * cf + pj;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = acc[e] * cf + pj * vn[e];
    m = m_new;
  }
  const int g = warp * 8 + grp;
  if (sub == 0) { sm_m[g] = m; sm_l[g] = l; }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm_acc[g * 33 + sub * 8 + e] = acc[e];
  __syncthreads();
  if (warp == 0) {
    float M = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < NG; ++i) M = fmaxf(M, sm_m[i]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      float wgt = (sm_m[i] == -CUDART_INF_F) ? 0.f : expf(sm_m[i] - M);
      num = fmaf(sm_acc[i * 33 + lane], wgt, num);
      den = fmaf(sm_l[i], wgt, den);
    }
    o[(long long)b * 512 + h * 32 + lane] = num / den;
  }
}

}  // namespace

void launch_attention(const Attn& p, cudaStream_t s) {
  if (p.B <= 0 || p.max_q <= 0) return;
  GENIE_CHECK(p.window <= 4, "attention: window > 4 unsupported");
  if (p.d == 32 && p.rel_k == nullptr && p.rel_v == nullptr && ((p.ldq | p.ldk | p.ldv) & 3) == 0 && (p.ldo & 1) == 0 &&
      ((reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.k) | reinterpret_cast<uintptr_t>(p.v)) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(p.o) & 7) == 0) {
    if (p.use_mma) {
      prefill_attention_mma_kernel<8><<<dim3((p.max_q + 127) / 128, p.H, p.B), 256, 0, s>>>(p);
      GENIE_LAUNCHED("prefill_attention_mma");
      return;
    }
    prefill_attention32_kernel<<<dim3((p.max_q + FQ - 1) / FQ, p.H, p.B), 256, 0, s>>>(p);
    GENIE_LAUNCHED("prefill_attention32");
    return;
  }
  dim3 grid((p.max_q + QT - 1) / QT, p.H, p.B);
  switch (p.d) {
    case 32: attention_kernel<32><<<grid, 128, 0, s>>>(p); break;
    case 64: attention_kernel<64><<<grid, 128, 0, s>>>(p); break;
    case 96: attention_kernel<96><<<grid, 128, 0, s>>>(p); break;
    case 128: attention_kernel<128><<<grid, 128, 0, s>>>(p); break;
    default: GENIE_CHECK(false, "attention: unsupported head dim");
  }
  GENIE_LAUNCHED("attention");
}

void launch_decode_attention_fused(const float* part, int nsplit, long long split_stride, const float* bias, float* o,
                                   void* kv_base, int kv_f16, long long utt_stride, long long layer_off,
                                   long long v_off, const int* kv_len, const int* active, int B, int cap, float scale,
                                   cudaStream_t s) {
  if (B <= 0) return;
  // slab size (`cap`) is what the host knows when the step is captured: long slabs take the 128-thread / 48 KB ring,
  // short ones the 64-thread / 16 KB ring (one wave); GENIE_ATT_BULK=0 falls back to the register-staged kernel
  static const int bulk_mode = [] { const char* e = getenv("GENIE_ATT_BULK"); return e ? atoi(e) : 3; }();
  const bool long_cache = cap >= 640;
  if (kv_f16 && long_cache && (bulk_mode & 1)) {
    constexpr size_t smem = att_bulk_smem<4, 128, 3>();
    static DynSmemAttr attr;
    attr.ensure(decode_attention16_bulk_kernel<true, 4, 128, 3>, smem);
    launch_pdl(decode_attention16_bulk_kernel<true, 4, 128, 3>, dim3(16, B), dim3(128), smem, s, part, nsplit,
               split_stride, bias, o, reinterpret_cast<__half*>(kv_base), utt_stride, layer_off, v_off, kv_len, active,
               cap, scale, 0, 1536);
  } else if (kv_f16 && !long_cache && (bulk_mode & 2)) {
    constexpr size_t smem = att_bulk_smem<2, 64, 2>();
    static DynSmemAttr attr;
    attr.ensure(decode_attention16_bulk_kernel<true, 2, 64, 2>, smem);
    launch_pdl(decode_attention16_bulk_kernel<true, 2, 64, 2>, dim3(16, B), dim3(64), smem, s, part, nsplit,
               split_stride, bias, o, reinterpret_cast<__half*>(kv_base), utt_stride, layer_off, v_off, kv_len, active,
               cap, scale, 0, 1536);
  } else if (kv_f16)
    launch_pdl(decode_attention16_kernel<true>, dim3(16, B), dim3(128), 0, s, part, nsplit, split_stride, bias, o,
               reinterpret_cast<__half*>(kv_base), utt_stride, layer_off, v_off, kv_len, active, cap, scale, 0, 1536);
  else
    launch_pdl(decode_attention_kernel<true>, dim3(16, B), dim3(128), 0, s, part, nsplit, split_stride, bias, o,
               reinterpret_cast<float*>(kv_base), utt_stride, layer_off, v_off, kv_len, active, cap, scale, 0, 1536);
  GENIE_LAUNCHED("decode_attention");
}

}  // namespace genie
