// One HiFi-GAN resblock pair in ONE kernel (C = 16 / 32 / 64 channels, sm_100a):
//
//     y = conv2( lrelu( conv1( lrelu(x) ) + b1 ) ) + b2 + x  [+ running sum]
//
// conv1: k taps, dilation d;  conv2: k taps, dilation 1;  both C -> C.  These stages are HBM-bound: run as
// two kernels a pair moves 16 bytes per element (x in, fp16 hand-over out and back in, residual in, y out);
// fused it moves 8 (x in once, the residual rows re-read from L2, y out).
//
// A CTA produces T = 512 - (k - 1) output rows:
//   1. x rows [q0 - p2 - p1, q0 - p2 - p1 + 512 + 2 p1) are loaded once, activated, rounded to fp16 and stored as
//      K-major swizzled rows (tc_halo_conv.cu layout: every tap is a row-shifted UMMA descriptor);
//   2. conv1 = k x 4 accumulate-MMAs (128 rows each) into TMEM: the 512 intermediate rows [q0 - p2, q0 - p2 + 512);
//   3. epilogue 1: TMEM -> registers -> + b1, lrelu, zero outside the utterance (conv2's zero padding), fp16 ->
//      a second swizzled tile in shared memory (each thread owns one row: plain 16-byte stores);
//   4. conv2 = k x 4 MMAs on that tile with row shifts 0..k-1 into the same TMEM columns;
//   5. epilogue 2 (tc_epilogue.cuh): + b2 + residual (+ running sum), 16-byte coalesced stores.
// The weights of one conv (k x C x C fp16: <= 22.5 KB at C <= 32, up to 88 KB at C = 64) sit in shared memory
// whole; conv2's are fetched with cp.async while epilogue 1 runs.  C = 64 (round 2) uses 256-row tiles (two
// accumulators, SWIZZLE_128B rows) so that x tile + intermediate tile + 88 KB of weights fit one SM.
#include "common.cuh"
#include "tc_epilogue.cuh"
#include <cstdlib>

namespace genie {
namespace {

constexpr int NTHR = 256;
constexpr int IPAD = 16;                                 // rows behind the intermediate tile that conv2's last taps touch

struct PairGeom {
  int k, d1, p1, p2, T, R1;                              // taps, conv1 dilation, paddings, outputs per CTA, staged x rows
  uint32_t x_bytes, i_bytes, w_bytes;                    // shared-memory regions (1024-aligned)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {   // bounded: false on timeout
  for (uint32_t i = 0; i < (1u << 22); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}

template <int C, int MT>
__global__ void __launch_bounds__(NTHR, C >= 64 ? 1 : (C == 16 ? 4 : 2)) tc_pair_conv_kernel(ConvGemm p, const __half* __restrict__ w1,
                                                                             const float* __restrict__ bias1, int kpad1,
                                                                             PairGeom g, int* err_flag) {
  constexpr int ROWS = MT * 128;
  constexpr int ROWB = 2 * C;                          // bytes per fp16 row (32 / 64 / 128)
  constexpr int CH = ROWB / 16;
  constexpr uint32_t SWMASK = ROWB == 128 ? 7u : ROWB == 64 ? 3u : 1u;    // Swizzle<3|2|1, 4, 3>
  constexpr uint64_t LAYOUT = ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6;       // SWIZZLE_128B / 64B / 32B
  constexpr int CW = C < 32 ? C : 32;                  // accumulator columns per tcgen05.ld chunk
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr uint32_t UNIT = C * ROWB;                  // one tap of weights
  constexpr uint32_t TCOLS = MT * C < 32 ? 32 : MT * C;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_to[ROWS];
  __shared__ __align__(16) float s_b1[64], s_b2[64];

  const int seg = blockIdx.z;
  int r0 = 0, Tseg = p.M;
  if (p.in_off) { r0 = p.in_off[seg]; Tseg = p.in_off[seg + 1] - r0; }
  const int q0 = blockIdx.x * g.T;                     // first output row of this CTA
  if (q0 >= Tseg) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sX = base, sI = base + g.x_bytes, sW = sI + g.i_bytes;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int r = tid; r < ROWS; r += NTHR) s_to[r] = (r < g.T && q0 + r < Tseg) ? q0 + r : -1;
  if (tid < 64) {
    s_b1[tid] = (tid < C && bias1) ? bias1[tid] : 0.f;
    s_b2[tid] = (tid < C && p.bias) ? p.bias[tid] : 0.f;
  }

  // weights of one conv: k taps x [C rows x C halves], K-major swizzled like the activation tiles
  auto load_w = [&](const __half* __restrict__ w, int kpad) {
    const int total = g.k * C * CH;
    for (int idx = tid; idx < total; idx += NTHR) {
      const int c = idx % CH, n = (idx / CH) % C, tap = idx / (CH * C);
      const uint32_t off = (uint32_t)(n * ROWB + c * 16);
      cp_async16(sW + (uint32_t)tap * UNIT + (off ^ (((off >> 7) & SWMASK) << 4)), w + (long long)n * kpad + tap * C + c * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_w(w1, kpad1);

  // ---- x tile: staged row j = input row q0 - p2 - p1 + j
  {
    const float* __restrict__ xg = p.x + (long long)r0 * p.ldx;
    const float pre = p.pre_slope;
    constexpr int cq = C / 4;
    const int totalA = g.R1 * cq;
    const int tbase = q0 - g.p2 - g.p1;
    for (int i0 = 0; i0 < totalA; i0 += NTHR * 8) {
      float4 v[8];
      uint32_t so[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = i0 + k * NTHR + tid;
        const int rr = idx / cq, f = idx - rr * cq;
        const int t = tbase + rr;
        const uint32_t off = (uint32_t)(rr * ROWB + f * 8);
        so[k] = idx < totalA ? (off ^ (((off >> 7) & SWMASK) << 4)) : 0xffffffffu;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < totalA && (unsigned)t < (unsigned)Tseg)
          v[k] = __ldg(reinterpret_cast<const float4*>(xg + (long long)t * p.ldx + f * 4));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (so[k] == 0xffffffffu) continue;
        float4 a = v[k];
        a.x = fmaxf(a.x, a.x * pre); a.y = fmaxf(a.y, a.y * pre);
        a.z = fmaxf(a.z, a.z * pre); a.w = fmaxf(a.w, a.w * pre);
        const __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h01);
        pk.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(sbase + so[k]) = pk;
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t desc_hi = ((uint64_t)(SBO >> 4) << 32) | ((uint64_t)1 << 46) | (LAYOUT << 61) | ((uint64_t)1 << 16);

  auto issue = [&](uint32_t tile, int step, uint64_t* bar) {   // k taps x MT accumulators, tap m reads rows + m * step
    for (int m = 0; m < g.k; ++m) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int j = 0; j < C / 16; ++j) {
          const uint32_t aaddr = tile + (uint32_t)(mt * 128 + m * step) * ROWB + j * 32;
          const uint32_t waddr = sW + (uint32_t)m * UNIT + j * 32;
          umma_f16(tmem + (uint32_t)(mt * C), desc_hi | (uint64_t)((aaddr & 0x3FFFFu) >> 4),
                   desc_hi | (uint64_t)((waddr & 0x3FFFFu) >> 4), idesc, (uint32_t)((m | j) != 0));
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
  };
  if (tid == 0) issue(sX, g.d1, &bars[0]);
  bool ok = mbar_wait(&bars[0], 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // conv1 is done with the weight buffer: fetch conv2's weights under epilogue 1
  load_w(p.tc_w, p.tc_kpad);

  // ---- epilogue 1: intermediate row i = mt * 128 + (warp & 3) * 32 + lane <-> input row q0 - p2 + i
  {
    const int rq = (warp & 3) * 32;
    for (int mt = warp >> 2; mt < MT; mt += 2) {
      const int i = mt * 128 + rq + lane;
      const int t = q0 - g.p2 + i;
      const bool inside = (unsigned)t < (unsigned)Tseg;          // outside the utterance conv2 sees zero padding
#pragma unroll
      for (int cc = 0; cc < C; cc += 32) {                       // 32 accumulator columns per tcgen05.ld
        uint32_t v[32];
        tc_epi::tmem_load_chunk(tmem + ((uint32_t)rq << 16) + (uint32_t)(mt * C + cc), C >= 32, v);
#pragma unroll
        for (int c8 = 0; c8 < CW / 8; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a = __uint_as_float(v[c8 * 8 + 2 * e]) + s_b1[cc + c8 * 8 + 2 * e];
            float b = __uint_as_float(v[c8 * 8 + 2 * e + 1]) + s_b1[cc + c8 * 8 + 2 * e + 1];
            a = fmaxf(a, a * 0.1f); b = fmaxf(b, b * 0.1f);
            const __half2 h = __floats2half2_rn(inside ? a : 0.f, inside ? b : 0.f);
            pk[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          const uint32_t off = (uint32_t)(i * ROWB + cc * 2 + c8 * 16);
          *reinterpret_cast<uint4*>(sbase + g.x_bytes + (off ^ (((off >> 7) & SWMASK) << 4))) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    // the IPAD rows behind the tile are only read for accumulator rows that are never stored: keep them finite
    for (int idx = tid; idx < IPAD * CH; idx += NTHR)
      *reinterpret_cast<uint4*>(sbase + g.x_bytes + (uint32_t)ROWS * ROWB + idx * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) issue(sI, 1, &bars[1]);
  ok = mbar_wait(&bars[1], 0u) && ok;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && err_flag) atomicExch(err_flag, 1);

  // ---- epilogue 2: accumulator row o <-> output row q0 + o (o < T); per-warp tiles alias the dead x tile
  tc_epi::Args ea;
  ea.y = p.y; ea.res = p.res; ea.acc = p.accumulate ? p.y : nullptr;
  ea.ldy = p.ldy; ea.ldr = p.ldr;
  ea.act = ACT_NONE; ea.slope = 0.f; ea.oscale = p.out_scale;
  ea.Cout = C;
  ea.vec = tc_epi::vec_ok(p.y, p.ldy, p.res, p.ldr, C);
  float* tile = reinterpret_cast<float*>(sbase) + warp * (C >= 32 ? tc_epi::TILE_FLOATS : tc_epi::TILE_FLOATS16);
  if (ok) {
    const int rq = (warp & 3) * 32;
    for (int mt = warp >> 2; mt < MT; mt += 2) {
      if (mt * 128 >= g.T || q0 + mt * 128 >= Tseg) break;
#pragma unroll
      for (int cc = 0; cc < C; cc += 32) {
        uint32_t v[32];
        tc_epi::tmem_load_chunk(tmem + ((uint32_t)rq << 16) + (uint32_t)(mt * C + cc), C >= 32, v);
        if (C >= 32) tc_epi::store_chunk<32>(v, tile, s_b2 + cc, s_to + mt * 128 + rq, r0, cc, ea, lane);
        else tc_epi::store_chunk<16>(v, tile, s_b2, s_to + mt * 128 + rq, r0, 0, ea, lane);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
  }
}

template <int C, int MT>
void launch_pair(const ConvGemm& p, const __half* w1, const float* bias1, int kpad1, int d1, int* err_flag,
                 cudaStream_t s) {
  constexpr int ROWS = MT * 128;
  PairGeom g;
  g.k = p.ntaps; g.d1 = d1; g.p1 = d1 * (g.k - 1) / 2; g.p2 = (g.k - 1) / 2;
  g.T = ROWS - 2 * g.p2; g.R1 = ROWS + 2 * g.p1;
  auto up = [](size_t v) { return (uint32_t)((v + 1023) / 1024 * 1024); };
  // epilogue-2 transpose tiles alias the x tile; 16-column chunks (C = 16) need 2.5 KB per warp, which is what lets
  // four CTAs of the C = 16 stage share an SM (its tiles are a chain of latencies: 1.3x from the fourth CTA)
  const size_t epi = (size_t)(NTHR / 32) * (C >= 32 ? tc_epi::TILE_FLOATS : tc_epi::TILE_FLOATS16) * 4;
  g.x_bytes = up(std::max((size_t)g.R1 * 2 * C, epi));
  g.i_bytes = up((size_t)(ROWS + IPAD) * 2 * C);
  g.w_bytes = up((size_t)g.k * C * 2 * C);
  const size_t smem = (size_t)g.x_bytes + g.i_bytes + g.w_bytes + 1024;
  GENIE_CHECK(smem <= 227 * 1024, "tc_pair_conv: tile does not fit shared memory");
  static DynSmemAttr attr;
  attr.ensure(tc_pair_conv_kernel<C, MT>, smem);
  dim3 grid((p.M + g.T - 1) / g.T, 1, p.B);
  tc_pair_conv_kernel<C, MT><<<grid, NTHR, smem, s>>>(p, w1, bias1, kpad1, g, err_flag);
  GENIE_LAUNCHED("tc_pair_conv");
}

}  // namespace

bool tc_pair_conv_supported(int C, int k) {
  static const bool c64 = [] { const char* e = getenv("GENIE_PAIR64"); return !(e && e[0] == '0'); }();
  return (C == 16 || C == 32 || (C == 64 && c64)) && k >= 3 && k <= 11 && (k & 1);
}

// p describes conv2 (weights tc_w / tc_kpad, bias, res, y, accumulate, segments) with x = the PAIR's input and
// pre_slope = its pre-activation; w1 / bias1 / d1 describe conv1
void launch_tc_pair_conv(const ConvGemm& p, const __half* w1, const float* bias1, int kpad1, int d1, int* err_flag,
                         cudaStream_t s) {
  GENIE_CHECK(p.Cin == p.Cout && tc_pair_conv_supported(p.Cin, p.ntaps), "tc_pair_conv: unsupported shape");
  GENIE_CHECK(p.tc_w && w1 && p.in_off && p.out_off == p.in_off && p.ldx % 4 == 0 && p.in_shift_step == 1,
              "tc_pair_conv: needs packed weights and segment offsets shared by input and output");
  GENIE_CHECK(p.pre_slope >= 0.f && p.pre_slope <= 1.f && IPAD >= p.ntaps - 1, "tc_pair_conv: bad activation / taps");
  if (p.M <= 0 || p.B <= 0) return;
  if (p.Cin == 16) launch_pair<16, 4>(p, w1, bias1, kpad1, d1, err_flag, s);
  else if (p.Cin == 32) launch_pair<32, 4>(p, w1, bias1, kpad1, d1, err_flag, s);
  else launch_pair<64, 2>(p, w1, bias1, kpad1, d1, err_flag, s);
}

}  // namespace genie
