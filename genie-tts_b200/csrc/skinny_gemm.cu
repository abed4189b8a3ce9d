// Decode-step linear layers for a small batch (M <= 128 rows): Y[M,N] = X[M,K] W[N,K]^T.
// Weight-streaming SIMT kernel: each warp owns CPW weight rows (fp16 as stored, exact) held in
// registers and sweeps all M activation rows; every (row, column) dot product uses the same
// lane partition and reduction order whatever M and the grid are, so a sequence's numerics do
// not depend on batch composition (and batch-1 is the same kernel).  fp32 FMA, fp32 accumulate.
// Split-K slices write raw partials that the following LayerNorm sums (with bias + residual).
#include "common.cuh"

namespace genie {
namespace {

constexpr int CHUNK = 256;     // K elements per pass: 32 lanes x 8 halves (one 16-byte load per lane)

template <int CPW, int NCHUNK>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(ConvGemm p, int M) {
  pdl_trigger();                               // the weight loads below overlap the predecessor's tail (PDL)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_base = (blockIdx.x * 8 + warp) * CPW;
  const bool live = n_base < p.Cout;           // dead warps still take part in the barriers below
  const int ks = blockIdx.y;
  const int k0 = ks * (NCHUNK * CHUNK);
  const __half* __restrict__ W = reinterpret_cast<const __half*>(p.w);

  // weights: CPW rows x NCHUNK chunks x 8 halves per lane
  float w[CPW][NCHUNK][8];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int n = n_base + c;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (live && n < p.Cout)
        u = __ldg(reinterpret_cast<const uint4*>(W + (long long)n * p.w_co_stride + k0 + ch * CHUNK + lane * 8));
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        w[c][ch][2 * e] = f.x; w[c][ch][2 * e + 1] = f.y;
      }
    }
  }
  pdl_wait();                                  // activations / residual: predecessor data, read through L2
  float* __restrict__ ybase = p.y + (long long)ks * p.split_stride;
  constexpr int RB = 4;                        // rows per round
  constexpr int KL = NCHUNK * CHUNK;           // K elements of this CTA's slice
  // the rows of a round are staged once per CTA in shared memory (every warp needs all of them; reading them
  // per warp through L2 cost 8x the weight traffic), 16-byte loads through L2 because they are producer data
  __shared__ __align__(16) float xs[RB][KL];
  for (int m0 = 0; m0 < M; m0 += RB) {
    __syncthreads();
    for (int i = threadIdx.x; i < RB * (KL / 4); i += 256) {
      const int r = i / (KL / 4), c4 = i - r * (KL / 4);
      const int m = (m0 + r < M) ? m0 + r : M - 1;       // clamp: tail rows are recomputed, not stored
      *reinterpret_cast<float4*>(&xs[r][c4 * 4]) =
          __ldcg(reinterpret_cast<const float4*>(p.x + (long long)m * p.ldx + k0 + c4 * 4));
    }
    __syncthreads();
    float4 xa[RB][NCHUNK], xc[RB][NCHUNK];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        xa[r][ch] = *reinterpret_cast<const float4*>(&xs[r][ch * CHUNK + lane * 8]);
        xc[r][ch] = *reinterpret_cast<const float4*>(&xs[r][ch * CHUNK + lane * 8 + 4]);
      }
    }
    float acc[RB][CPW];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
#pragma unroll
      for (int c = 0; c < CPW; ++c) acc[r][c] = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const float xv[8] = {xa[r][ch].x, xa[r][ch].y, xa[r][ch].z, xa[r][ch].w,
                             xc[r][ch].x, xc[r][ch].y, xc[r][ch].z, xc[r][ch].w};
#pragma unroll
        for (int c = 0; c < CPW; ++c)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[r][c] = fmaf(w[c][ch][e], xv[e], acc[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        float v = acc[r][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][c] = v;
      }
    // lane (r*CPW + c) stores element (row m0+r, column n_base+c)
    if (live && lane < RB * CPW) {
      const int r = lane / CPW, c = lane % CPW;
      float v = 0.f;
#pragma unroll
      for (int rr = 0; rr < RB; ++rr)
#pragma unroll
        for (int cc = 0; cc < CPW; ++cc)
          if (r == rr && c == cc) v = acc[rr][cc];
      const int m = m0 + r, n = n_base + c;
      if (m < M && n < p.Cout) {
        if (p.ksplit == 1) {
          if (p.bias) v += p.bias[n];
          if (p.act == ACT_RELU) v = v > 0.f ? v : 0.f;
          if (p.res) v += __ldcg(p.res + (long long)m * p.ldr + n);
        }
        ybase[(long long)m * p.ldy + n] = v;
      }
    }
  }
}

}  // namespace

// requirements: fp16 weights [N,K] row-major 16-byte aligned rows, K / ksplit in {256, 512}, M <= 128
bool skinny_gemm_supported(const ConvGemm& p) {
  if (!p.w_f16 || p.ntaps != 1 || p.in_off || p.M > 128 || p.pre_slope != 1.f || p.accumulate || p.bias2) return false;
  if (p.act != ACT_NONE && p.act != ACT_RELU) return false;
  if (p.Cin % p.ksplit) return false;
  const int kl = p.Cin / p.ksplit;
  return (kl == 256 || kl == 512) && (p.w_co_stride % 8 == 0) && (p.ldx % 4 == 0);
}

void launch_skinny_gemm(const ConvGemm& p, cudaStream_t s) {
  GENIE_CHECK(skinny_gemm_supported(p), "skinny_gemm: unsupported shape");
  if (p.M <= 0) return;
  const int kl = p.Cin / p.ksplit;
  // few rows: one column per warp so that enough CTAs stream the weights; more rows: 4 columns per warp
  // so that each activation row is read once per 4 columns
  const bool wide = p.M > 8;
  const int cpw = wide ? 4 : 1;
  dim3 grid((p.Cout + 8 * cpw - 1) / (8 * cpw), p.ksplit);
  if (wide) {
    if (kl == 512) launch_pdl(skinny_gemm_kernel<4, 2>, grid, dim3(256), 0, s, p, p.M);
    else launch_pdl(skinny_gemm_kernel<4, 1>, grid, dim3(256), 0, s, p, p.M);
  } else {
    if (kl == 512) launch_pdl(skinny_gemm_kernel<1, 2>, grid, dim3(256), 0, s, p, p.M);
    else launch_pdl(skinny_gemm_kernel<1, 1>, grid, dim3(256), 0, s, p, p.M);
  }
  GENIE_LAUNCHED("skinny_gemm");
}

}  // namespace genie
