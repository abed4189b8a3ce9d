// tcgen05 k-tap convolution with the activation tile staged ONCE per CTA (sm_100a).
//
// tc_gemm.cu rebuilds the A operand for every tap (an im2col gather: load, convert, swizzled
// store), which is what bounds the HiFi-GAN resblock convs (k = 3/7/11, dilation 1/3/5): every
// stage of the generator costs the same ~300 us per tap whatever its width.  Here the CTA loads
// the rows [q0 + lo, q0 + MT*128 + hi) of the fp32 channels-last input once, applies the
// pre-activation, rounds to fp16 and stores them as K-major swizzled rows (one 64-channel slab
// per 128-byte row; 32 / 16 channel layers use the 64B / 32B swizzle modes).  Tap m is then just
// a shifted view: the A descriptor of MMA (tap, slab, k16 step) starts `shift(m)` rows into the
// slab.  The k-loop therefore only streams weights: cp.async (16 B, zero register traffic) from
// the packed fp16 [Cout][K] matrix into a 4-slot ring of swizzled [NT x 64] tiles, MMAs issued
// two slots behind the loads, slot reuse gated by tcgen05.commit -> mbarrier.
//
// MT accumulators of 128 rows each live side by side in TMEM, so narrow layers (NT = 16/32)
// amortise the halo (up to 50 rows) and the per-CTA setup over 512 output rows.
//
// Epilogue as in tc_gemm.cu: tcgen05.ld -> bias/activation in registers -> per-warp padded
// transpose -> coalesced row stores with the residual / accumulate adds.
#include "common.cuh"
#include "tc_epilogue.cuh"

#include <algorithm>
#include <cstdlib>

namespace genie {
namespace {

constexpr int NTHR = 256;
constexpr int NSLOT = 4;                            // weight ring slots (loads run 2 iterations ahead)
constexpr int AHEAD = 2;
constexpr size_t EPI_BYTES = (NTHR / 32) * tc_epi::TILE_FLOATS * 4;

struct HaloGeom {
  int lo = 0;                 // smallest tap shift: halo row rr holds input row q0 + lo + rr
  int R = 0;                  // rows staged per CTA
  int slabs = 1;              // 64-channel slabs (1 for Cin <= 64)
  int cpad = 0;               // channels per staged row incl. zero padding (24 -> 32, 48 -> 64, 96 -> 128)
  uint32_t slab_bytes = 0;
  int U = 1, NU = 1, NI = 1;  // units (tap, slab) per ring slot / total / ring iterations
  uint32_t slot_bytes = 0;
  int flags = 0;              // GENIE_TC_HALO bit 0: force wide layers onto the cp.async kernel (testing)
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}

template <int NT, int MT, int ROWB>
__global__ void __launch_bounds__(NTHR, (ROWB < 128 ? 4 : NT <= 64 ? 3 : 2)) tc_halo_conv_kernel(ConvGemm p, HaloGeom g, int* err_flag) {
  constexpr int KS = ROWB / 2;                      // channels per shared-memory row
  constexpr int CH = ROWB / 16;                     // 16-byte chunks per row
  constexpr uint32_t SWMASK = ROWB == 128 ? 7u : ROWB == 64 ? 3u : 1u;     // Swizzle<3|2|1, 4, 3>
  constexpr uint64_t LAYOUT = ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6;        // SWIZZLE_128B / 64B / 32B
  constexpr uint32_t SBO = 8 * ROWB;                // 8-row group pitch
  constexpr uint32_t UNIT_BYTES = NT * ROWB;        // one (tap, slab) weight tile
  constexpr uint32_t TCOLS = MT * NT < 32 ? 32 : MT * NT;
  constexpr int ROWS = MT * 128;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[NSLOT];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_to[ROWS];                        // output row of accumulator row (or -1)
  __shared__ __align__(16) float s_bias[NT < 32 ? 32 : NT];

  const int seg = blockIdx.z;
  int in0 = 0, Tin = p.M, out0 = 0, Tout = p.M_out;
  if (p.in_off) { in0 = p.in_off[seg]; Tin = p.in_off[seg + 1] - in0; }
  if (p.out_off) { out0 = p.out_off[seg]; Tout = p.out_off[seg + 1] - out0; }
  const int nq = Tin + p.q_extra;
  const int q0 = blockIdx.x * ROWS;
  if (q0 >= nq) return;
  const int n0 = (int)blockIdx.y * NT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base;
  const uint32_t sW = base + (uint32_t)g.slabs * g.slab_bytes;

  int n_mma = p.Cout - n0;
  if (n_mma > NT) n_mma = NT;
  n_mma = (n_mma + 15) & ~15;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int r = tid; r < ROWS; r += NTHR) {
    const int q = q0 + r;
    const int to = q * p.out_mul + p.out_add;
    s_to[r] = (q < nq && to >= 0 && to < Tout) ? to : -1;
  }
  for (int j = tid; j < NT; j += NTHR) {
    float bsum = 0.f;
    const int n = n0 + j;
    if (n < p.Cout) {
      if (p.bias) bsum += p.bias[n];
      if (p.bias2) bsum += p.bias2[(long long)seg * p.ldb2 + n];
    }
    s_bias[j] = bsum;
  }

  const __half* __restrict__ whi = p.tc_w;
  // ---- weight ring: iteration `it` = units [it*U, it*U + nu) into slot it % NSLOT
  auto load_w = [&](int it) {
    const int u0 = it * g.U;
    const int nu = min(g.U, g.NU - u0);
    const uint32_t dst = sW + (uint32_t)(it % NSLOT) * g.slot_bytes;
    const int total = nu * (NT * CH);
    for (int idx = tid; idx < total; idx += NTHR) {
      const int c = idx % CH, n = (idx / CH) % NT, ul = idx / (CH * NT);
      if (n >= n_mma) continue;
      const int u = u0 + ul;
      const int tap = u / g.slabs, sl = u - tap * g.slabs;
      // channel-padded layers: chunks beyond Cin read the next tap's weights (finite, multiplied by the zero
      // padding of the activation tile); only the end of the packed row needs the guard
      const int kcol = tap * p.Cin + sl * 64 + c * 8;
      const bool valid = n0 + n < p.Cout && kcol + 8 <= p.tc_kpad;
      const __half* src = whi + (long long)(n0 + n) * p.tc_kpad + kcol;
      const uint32_t off = (uint32_t)(n * ROWB + c * 16);
      cp_async16(dst + (uint32_t)ul * UNIT_BYTES + (off ^ (((off >> 7) & SWMASK) << 4)), valid ? src : whi,
                 valid ? 16 : 0);
    }
  };
#pragma unroll
  for (int it = 0; it < AHEAD; ++it) {
    if (it < g.NI) load_w(it);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- activation halo tile: each input element is loaded, activated, rounded and stored once
  if (p.x16) {
    // ready-made operand rows (fp16, already activated): 16-byte copies into the swizzled tile
    const __half* __restrict__ xh = p.x16 + (long long)in0 * p.Cin;
    const int c8n = g.cpad >> 3;                    // 16-byte chunks per staged row
    const int totalA = g.R * c8n;
    const int tbase = q0 + g.lo;
    for (int i0 = 0; i0 < totalA; i0 += NTHR * 8) {
      uint4 v[8];
      uint32_t so[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = i0 + k * NTHR + tid;
        const int rr = idx / c8n, f = idx - rr * c8n;
        const int t = tbase + rr;
        const int c = f * 8;
        const uint32_t off = (uint32_t)(rr * ROWB + (c & 63) * 2);
        so[k] = idx < totalA ? (uint32_t)(c >> 6) * g.slab_bytes + (off ^ (((off >> 7) & SWMASK) << 4)) : 0xffffffffu;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (idx < totalA && c < p.Cin && (unsigned)t < (unsigned)Tin)
          v[k] = __ldg(reinterpret_cast<const uint4*>(xh + (long long)t * p.Cin + c));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (so[k] != 0xffffffffu) *reinterpret_cast<uint4*>(sbase + so[k]) = v[k];
    }
  } else {
    const float* __restrict__ xg = p.x + (long long)in0 * p.ldx;
    const float pre = p.pre_slope;                  // 0 <= pre <= 1: lrelu(v) == max(v, v * pre)
    const int cq = g.cpad >> 2;                     // float4 per staged row (channels >= Cin are zero padding)
    const int totalA = g.R * cq;
    const int tbase = q0 + g.lo;
    for (int i0 = 0; i0 < totalA; i0 += NTHR * 8) {
      float4 v[8];
      uint32_t so[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int idx = i0 + k * NTHR + tid;
        const int rr = idx / cq, f = idx - rr * cq;
        const int t = tbase + rr;
        const int c = f * 4;
        const uint32_t off = (uint32_t)(rr * ROWB + (c & 63) * 2);
        so[k] = idx < totalA ? (uint32_t)(c >> 6) * g.slab_bytes + (off ^ (((off >> 7) & SWMASK) << 4)) : 0xffffffffu;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < totalA && c < p.Cin && (unsigned)t < (unsigned)Tin)
          v[k] = __ldg(reinterpret_cast<const float4*>(xg + (long long)t * p.ldx + c));
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t desc_hi = ((uint64_t)(SBO >> 4) << 32) | ((uint64_t)1 << 46) | (LAYOUT << 61) | ((uint64_t)1 << 16);
  bool ok = true;

  for (int it = 0; it < g.NI; ++it) {
    const int nx = it + AHEAD;                      // prefetch two iterations ahead
    if (nx < g.NI) {
      if (nx >= NSLOT) ok = mbar_wait(&bars[nx % NSLOT], (uint32_t)((nx / NSLOT - 1) & 1)) && ok;
      load_w(nx);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic / cp.async writes -> tensor-core reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int u0 = it * g.U;
      const int nu = min(g.U, g.NU - u0);
      const uint32_t wslot = sW + (uint32_t)(it % NSLOT) * g.slot_bytes;
      for (int ul = 0; ul < nu; ++ul) {
        const int u = u0 + ul;
        const int tap = u / g.slabs, sl = u - tap * g.slabs;
        const int shift = p.in_shift0 + tap * p.in_shift_step - g.lo;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int j = 0; j < KS / 16; ++j) {
            const uint32_t aaddr = sA + (uint32_t)sl * g.slab_bytes + (uint32_t)(mt * 128 + shift) * ROWB + j * 32;
            const uint32_t waddr = wslot + (uint32_t)ul * UNIT_BYTES + j * 32;
            const uint64_t da = desc_hi | (uint64_t)((aaddr & 0x3FFFFu) >> 4);
            const uint64_t db = desc_hi | (uint64_t)((waddr & 0x3FFFFu) >> 4);
            umma_f16(tmem + (uint32_t)(mt * NT), da, db, idesc, (uint32_t)((u | j) != 0));
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   ::"r"(smem_u32(&bars[it % NSLOT])) : "memory");
    }
  }
  {
    const int last = g.NI - 1;
    ok = mbar_wait(&bars[last % NSLOT], (uint32_t)((last / NSLOT) & 1)) && ok;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && err_flag) atomicExch(err_flag, 1);
  __syncthreads();                                  // every warp is past its waits: the tile region can be reused

  // ---- epilogue: chunk = (accumulator mt, 32 columns); warp w takes lane quarter (w & 3) of the chunks
  // with index parity (w >> 2)
  tc_epi::Args ea;
  ea.y = p.y; ea.res = p.res; ea.acc = p.accumulate ? p.y : nullptr;
  ea.ldy = p.ldy; ea.ldr = p.ldr;
  ea.act = p.act; ea.slope = p.act == ACT_RELU ? 0.f : p.act_slope; ea.oscale = p.out_scale;
  ea.Cout = p.Cout;
  ea.vec = tc_epi::vec_ok(p.y, p.ldy, p.res, p.ldr, p.Cout);
  ea.y16 = p.y16; ea.ldy16 = p.Cout;
  float* tile = reinterpret_cast<float*>(sbase) + warp * tc_epi::TILE_FLOATS;
  const int rq = (warp & 3) * 32;
  const int ncc = (n_mma + 31) >> 5;                // column chunks per accumulator
  if (ok) {
    for (int chunk = warp >> 2; chunk < MT * ncc; chunk += 2) {
      const int mt = chunk / ncc, c0 = (chunk - mt * ncc) * 32;
      if (q0 + mt * 128 >= nq) break;
      uint32_t v[32];
      const bool full = n_mma - c0 >= 32;
      tc_epi::tmem_load_chunk(tmem + ((uint32_t)rq << 16) + (uint32_t)(mt * NT + c0), full, v);
      if (full) tc_epi::store_chunk<32>(v, tile, s_bias + c0, s_to + mt * 128 + rq, out0, n0 + c0, ea, lane);
      else tc_epi::store_chunk<16>(v, tile, s_bias + c0, s_to + mt * 128 + rq, out0, n0 + c0, ea, lane);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
  }
}

// ---------------------------------------------------------------------------
// Wide layers (Cin >= 128): warp-specialised variant.  The weights are pre-tiled at load time
// ([n-tile][tap][slab] -> one contiguous, already swizzled 128 x 64 fp16 tile of 16 KB), so a single
// thread streams them with cp.async.bulk into a 4-slot ring (mbarrier expect_tx / complete_tx) and issues
// the MMAs; slot reuse is gated by tcgen05.commit.  The 8 loader warps stage the halo tile, signal it
// through an mbarrier and go straight to waiting for the accumulator: there is no CTA-wide barrier in
// the k-loop, and with two CTAs per SM one CTA's epilogue overlaps the other's MMAs.
// ---------------------------------------------------------------------------
constexpr int BNT = 128;                            // output columns per CTA
constexpr int BTHR = NTHR + 32;                     // 8 loader / epilogue warps + 1 producer / MMA warp
constexpr uint32_t BTILE = BNT * 128;               // 16 KB weight tile

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void pretile_w128_kernel(const __half* __restrict__ hi, int Cout, int kpad, int Cin, int ntaps, __half* tiles) {
  const int slabs = (Cin + 63) / 64;
  const long long total = (long long)((Cout + BNT - 1) / BNT) * ntaps * slabs * BNT * 8;     // 16-byte chunks
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i & 7), n = (int)((i >> 3) % BNT);
  const long long unit = i / (BNT * 8);
  const int sl = (int)(unit % slabs), tap = (int)((unit / slabs) % ntaps), nt = (int)(unit / ((long long)slabs * ntaps));
  uint4 v = make_uint4(0u, 0u, 0u, 0u);                                    // rows past Cout: zero (last N tile padded)
  if (nt * BNT + n < Cout && sl * 64 + c8 * 8 < Cin)                       // channels past Cin: zero (last slab padded)
    v = *reinterpret_cast<const uint4*>(hi + (long long)(nt * BNT + n) * kpad + tap * Cin + sl * 64 + c8 * 8);
  const uint32_t off = (uint32_t)(n * 128 + c8 * 16);
  *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(tiles) + unit * BTILE + (off ^ (((off >> 7) & 7u) << 4))) = v;
}

__global__ void __launch_bounds__(BTHR, 2) tc_halo_bulk_kernel(ConvGemm p, HaloGeom g, int* err_flag) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[NSLOT], empty[NSLOT], a_ready, acc_ready;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_to[128];
  __shared__ __align__(16) float s_bias[BNT];

  const int seg = blockIdx.z;
  int in0 = 0, Tin = p.M, out0 = 0, Tout = p.M_out;
  if (p.in_off) { in0 = p.in_off[seg]; Tin = p.in_off[seg + 1] - in0; }
  if (p.out_off) { out0 = p.out_off[seg]; Tout = p.out_off[seg + 1] - out0; }
  const int nq = Tin + p.q_extra;
  const int q0 = blockIdx.x * 128;
  if (q0 >= nq) return;
  const int n0 = (int)blockIdx.y * BNT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base;
  const uint32_t sW = base + (uint32_t)g.slabs * g.slab_bytes;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)BNT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&a_ready, NTHR);
    mbar_init(&acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 128) {
    const int q = q0 + tid;
    const int to = q * p.out_mul + p.out_add;
    s_to[tid] = (q < nq && to >= 0 && to < Tout) ? to : -1;
    float bsum = 0.f;
    const int n = n0 + tid;
    if (n < p.Cout) {
      if (p.bias) bsum += p.bias[n];
      if (p.bias2) bsum += p.bias2[(long long)seg * p.ldb2 + n];
    }
    s_bias[tid] = bsum;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool ok = true;

  if (warp == NTHR / 32) {
    // ===== producer + MMA issuer: one thread
    if (lane == 0) {
      const uint8_t* tiles = reinterpret_cast<const uint8_t*>(p.tc_tiles) + (size_t)blockIdx.y * g.NU * BTILE;
      const int npre = g.NU < NSLOT ? g.NU : NSLOT;
      for (int u = 0; u < npre; ++u) {
        mbar_expect_tx(&full[u], BTILE);
        bulk_g2s(sW + (uint32_t)u * BTILE, tiles + (size_t)u * BTILE, BTILE, &full[u]);
      }
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
      ok = mbar_wait(&a_ready, 0u) && ok;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int u = 0; u < g.NU; ++u) {
        const int slot = u % NSLOT;
        ok = mbar_wait(&full[slot], (uint32_t)((u / NSLOT) & 1)) && ok;
        const int tap = u / g.slabs, sl = u - tap * g.slabs;
        const int shift = p.in_shift0 + tap * p.in_shift_step - g.lo;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t aaddr = sA + (uint32_t)sl * g.slab_bytes + (uint32_t)shift * 128u + j * 32;
          const uint32_t waddr = sW + (uint32_t)slot * BTILE + j * 32;
          umma_f16(tmem, desc_hi | (uint64_t)((aaddr & 0x3FFFFu) >> 4), desc_hi | (uint64_t)((waddr & 0x3FFFFu) >> 4),
                   idesc, (uint32_t)((u | j) != 0));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     ::"r"(smem_u32(&empty[slot])) : "memory");
        // refill the slot of the PREVIOUS unit (its MMAs finish while this unit's run)
        if (u >= 1) {
          const int v = u - 1 + NSLOT, ps = (u - 1) % NSLOT;
          if (v < g.NU) {
            ok = mbar_wait(&empty[ps], (uint32_t)(((u - 1) / NSLOT) & 1)) && ok;
            mbar_expect_tx(&full[ps], BTILE);
            bulk_g2s(sW + (uint32_t)ps * BTILE, tiles + (size_t)v * BTILE, BTILE, &full[ps]);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   ::"r"(smem_u32(&acc_ready)) : "memory");
      if (!ok && err_flag) atomicExch(err_flag, 1);
    }
  } else {
    // ===== loader / epilogue warps: halo tile once, then wait for the accumulator
    if (p.x16) {
      const __half* __restrict__ xh = p.x16 + (long long)in0 * p.Cin;
      const int c8n = p.Cin >> 3;
      const int totalA = g.R * c8n;
      const int tbase = q0 + g.lo;
      for (int i0 = 0; i0 < totalA; i0 += NTHR * 8) {
        uint4 v[8];
        uint32_t so[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int idx = i0 + k * NTHR + tid;
          const int rr = idx / c8n, f = idx - rr * c8n;
          const int t = tbase + rr;
          const int c = f * 8;
          const uint32_t off = (uint32_t)(rr * 128 + (c & 63) * 2);
          so[k] = idx < totalA ? (uint32_t)(c >> 6) * g.slab_bytes + (off ^ (((off >> 7) & 7u) << 4)) : 0xffffffffu;
          v[k] = make_uint4(0u, 0u, 0u, 0u);
          if (idx < totalA && (unsigned)t < (unsigned)Tin)
            v[k] = __ldg(reinterpret_cast<const uint4*>(xh + (long long)t * p.Cin + c));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (so[k] != 0xffffffffu) *reinterpret_cast<uint4*>(sbase + so[k]) = v[k];
      }
    } else {
      const float* __restrict__ xg = p.x + (long long)in0 * p.ldx;
      const float pre = p.pre_slope;
      const int cq = p.Cin >> 2;
      const int totalA = g.R * cq;
      const int tbase = q0 + g.lo;
      for (int i0 = 0; i0 < totalA; i0 += NTHR * 8) {
        float4 v[8];
        uint32_t so[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int idx = i0 + k * NTHR + tid;
          const int rr = idx / cq, f = idx - rr * cq;
          const int t = tbase + rr;
          const int c = f * 4;
          const uint32_t off = (uint32_t)(rr * 128 + (c & 63) * 2);
          so[k] = idx < totalA ? (uint32_t)(c >> 6) * g.slab_bytes + (off ^ (((off >> 7) & 7u) << 4)) : 0xffffffffu;
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < totalA && (unsigned)t < (unsigned)Tin)
            v[k] = __ldg(reinterpret_cast<const float4*>(xg + (long long)t * p.ldx + c));
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
    mbar_arrive(&a_ready);
    ok = mbar_wait(&acc_ready, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && err_flag) atomicExch(err_flag, 1);
    // all 8 warps are past acc_ready only after every MMA has completed: the halo tile can be reused; the
    // per-warp epilogue tiles are disjoint, so no further CTA-wide barrier is needed here
    tc_epi::Args ea;
    ea.y = p.y; ea.res = p.res; ea.acc = p.accumulate ? p.y : nullptr;
    ea.ldy = p.ldy; ea.ldr = p.ldr;
    ea.act = p.act; ea.slope = p.act == ACT_RELU ? 0.f : p.act_slope; ea.oscale = p.out_scale;
    ea.Cout = p.Cout;
    ea.vec = tc_epi::vec_ok(p.y, p.ldy, p.res, p.ldr, p.Cout);
    ea.y16 = p.y16; ea.ldy16 = p.Cout;
    float* tile = reinterpret_cast<float*>(sbase) + warp * tc_epi::TILE_FLOATS;
    const int rq = (warp & 3) * 32;
    if (ok) {
      for (int c0 = (warp >> 2) * 32; c0 < BNT; c0 += 64) {
        uint32_t v[32];
        tc_epi::tmem_load_chunk(tmem + ((uint32_t)rq << 16) + (uint32_t)c0, true, v);
        tc_epi::store_chunk<32>(v, tile, s_bias + c0, s_to + rq, out0, n0 + c0, ea, lane);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)BNT));
  }
}

// ---------------------------------------------------------------------------
// Round 2: the same operands as tc_halo_bulk_kernel in a PERSISTENT, role-split CTA - one per SM.
//
// What bounds the one-tile kernel above (ncu, C = 128 k = 3: 2.4 TB/s read, tensor pipe 10 % active): every global
// load sits in a register until it is converted, so the bytes in flight per SM are threads x 8 x 16 B = 32-64 KB,
// and each CTA is a chain of latencies (stage the halo -> MMAs -> residual loads -> stores).  Here:
//   warps 0-3  loaders: halo tile of tile i+1 into the second A buffer while tile i is multiplied - fp32 rows
//              with 24 x 16 B per thread in flight (one CTA per SM leaves the registers for it), fp16 rows with
//              cp.async straight into the swizzled tile; they also PREFETCH the residual rows of the tile with
//              cp.async.bulk (one 256-byte row segment per thread and column half, mbarrier complete_tx) so that
//              the epilogue never waits on a global load
//   warp  8    one thread streams the pre-tiled weights (cp.async.bulk ring, running ahead across tile
//              boundaries) and issues the MMAs of tile i into TMEM accumulator i & 1; its loop carries slots,
//              phases and descriptor words incrementally (no division between two tcgen05.mma)
//   warps 4-7  epilogue: drain accumulator (i-1) & 1 while tile i is multiplied; residual from shared memory
// All waits are bounded mbarrier polls (err_flag on timeout).
// ---------------------------------------------------------------------------
constexpr int PLD = 128;                            // loader threads (warps 0-3); epilogue threads = warps 4-7
constexpr uint32_t RES_HALF = 128 * 64 * 4;         // residual staging: 128 rows x 64 columns fp32 per half

struct PipeTile { int seg, q0, ny, in0, Tin, out0, Tout, nq; };

__device__ __forceinline__ bool pipe_tile(const ConvGemm& p, int t, int TX, int NY, PipeTile& o) {
  o.ny = t % NY;
  const int r = t / NY;
  o.seg = r / TX;
  o.q0 = (r - o.seg * TX) * 128;
  o.in0 = 0; o.Tin = p.M; o.out0 = 0; o.Tout = p.M_out;
  if (p.in_off) { o.in0 = __ldg(p.in_off + o.seg); o.Tin = __ldg(p.in_off + o.seg + 1) - o.in0; }
  if (p.out_off) { o.out0 = __ldg(p.out_off + o.seg); o.Tout = __ldg(p.out_off + o.seg + 1) - o.out0; }
  o.nq = o.Tin + p.q_extra;
  return o.q0 < o.nq;
}

// stages: a tile's K range is staged in `stages` passes of <= 2 slabs (128 input channels) through the same A buffer,
// all accumulating into one TMEM tile: Cin = 192 / 256 / 384 ... need no more shared memory than Cin = 128
struct PipeCfg { int TX, total, nslot, nA, use_res, stages, slabs_total, NY; };

// PUNR: 16-byte loads in flight per loader thread (fp32 rows); MINB: CTAs per SM the registers are budgeted for
template <int PUNR, int MINB>
__global__ void __launch_bounds__(BTHR, MINB) tc_halo_pipe_kernel(ConvGemm p, HaloGeom g, PipeCfg cfg, int* err_flag) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[NSLOT], empty[NSLOT], a_ready[2], a_free[2], acc_ready[2], acc_free[2],
      res_ready[2], res_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_to[2][128];
  __shared__ __align__(16) float s_bias[2][BNT];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NY = cfg.NY;
  const int TX = cfg.TX, total = cfg.total, nslot = cfg.nslot, nA = cfg.nA;
  const int S = cfg.stages, SLT = cfg.slabs_total;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_bytes = (uint32_t)g.slabs * g.slab_bytes;
  const uint32_t sA = base;
  const uint32_t oW = a_bytes * (uint32_t)nA;
  const uint32_t sW = base + oW;
  const uint32_t oR = oW + (uint32_t)nslot * BTILE;                 // residual staging (2 halves) when cfg.use_res
  const uint32_t oE = oR + (cfg.use_res ? 2 * RES_HALF : 0u);       // epilogue transpose tiles: 4 warps

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(2 * BNT)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_ready[i], PLD); mbar_init(&a_free[i], 1);
      mbar_init(&acc_ready[i], 1); mbar_init(&acc_free[i], 128);
      mbar_init(&res_ready[i], 1); mbar_init(&res_free[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  bool ok = true;

  if (warp == 8) {
    // ===== weight producer + MMA issuer: one thread
    if (lane == 0) {
      const uint8_t* tiles = reinterpret_cast<const uint8_t*>(p.tc_tiles);
      // prefetch cursor: runs up to nslot units ahead of the MMAs, across tile boundaries
      PipeTile pt;
      int ptile = (int)blockIdx.x - (int)gridDim.x;
      int ps = 0, ptap = 0, psl = 0;                 // prefetch cursor inside the tile: stage, tap, slab of the stage
      bool pvalid = false;
      auto padvance = [&]() {
        for (ptile += (int)gridDim.x; ptile < total; ptile += (int)gridDim.x)
          if (pipe_tile(p, ptile, TX, NY, pt)) return true;
        return false;
      };
      pvalid = padvance();
      const size_t unit_tile = (size_t)p.ntaps * SLT * BTILE;       // bytes of one N tile's weights
      auto fetch_into = [&](int slot) {
        const uint8_t* src = tiles + (size_t)pt.ny * unit_tile + ((size_t)ptap * SLT + 2 * ps + psl) * BTILE;
        mbar_expect_tx(&full[slot], BTILE);
        bulk_g2s(sW + (uint32_t)slot * BTILE, src, BTILE, &full[slot]);
        const int nsl = SLT - 2 * ps < 2 ? SLT - 2 * ps : 2;
        if (++psl == nsl) {
          psl = 0;
          if (++ptap == p.ntaps) {
            ptap = 0;
            if (++ps == S) { ps = 0; pvalid = padvance(); }
          }
        }
      };
      for (int k = 0; k < nslot && pvalid; ++k) fetch_into(k);

      const uint32_t idesc = (1u << 4) | ((uint32_t)(BNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t desc_hi32 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);      // SBO, version, SWIZZLE_128B
      const uint32_t lbo16 = 1u << 16;
      const uint32_t a_lo0 = ((sA & 0x3FFFFu) >> 4) + (uint32_t)((p.in_shift0 - g.lo) * 8) + lbo16;
      const uint32_t a_tap_step = (uint32_t)(p.in_shift_step * 8), a_slab_step = g.slab_bytes >> 4;
      const uint32_t w_lo0 = ((sW & 0x3FFFFu) >> 4) + lbo16;
      int slot = 0, prev_slot = 0;
      uint32_t ring_phase = 0, prev_phase = 0;
      bool first_unit = true;
      int i = 0, ab = 0;
      uint32_t a_phase = 0;                          // parity of the current use of A buffer ab
      PipeTile tl;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        if (!pipe_tile(p, t, TX, NY, tl)) continue;
        const int b = i & 1;
        if (i >= 2) ok = mbar_wait(&acc_free[b], (uint32_t)(((i >> 1) - 1) & 1)) && ok;   // tile i-2 has left TMEM
        const uint32_t acc = tmem + (uint32_t)(b * BNT);
        for (int st = 0; st < S; ++st) {
          const int nsl = SLT - 2 * st < 2 ? SLT - 2 * st : 2;
          ok = mbar_wait(&a_ready[ab], a_phase) && ok;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t a_tap = a_lo0 + (uint32_t)ab * (a_bytes >> 4);
          for (int tap = 0; tap < p.ntaps; ++tap) {
            for (int sl = 0; sl < nsl; ++sl) {
              ok = mbar_wait(&full[slot], ring_phase) && ok;
              const uint32_t a_lo = a_tap + (uint32_t)sl * a_slab_step;
              const uint32_t w_lo = w_lo0 + (uint32_t)slot * (BTILE >> 4);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(acc, ((uint64_t)desc_hi32 << 32) | (uint64_t)(a_lo + j * 2),
                         ((uint64_t)desc_hi32 << 32) | (uint64_t)(w_lo + j * 2), idesc, (uint32_t)((st | tap | sl | j) != 0));
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                           ::"r"(smem_u32(&empty[slot])) : "memory");
              // refill the slot of the PREVIOUS unit (its MMAs finish while this unit's run)
              if (!first_unit && pvalid) {
                ok = mbar_wait(&empty[prev_slot], prev_phase) && ok;
                fetch_into(prev_slot);
              }
              first_unit = false;
              prev_slot = slot; prev_phase = ring_phase;
              if (++slot == nslot) { slot = 0; ring_phase ^= 1u; }
            }
            a_tap += a_tap_step;
          }
          if (st == S - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(&acc_ready[b])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                       ::"r"(smem_u32(&a_free[ab])) : "memory");
          if (++ab == nA) { ab = 0; a_phase ^= 1u; }
        }
        ++i;
      }
      if (!ok && err_flag) atomicExch(err_flag, 1);
    }
  } else if (warp < 4) {
    // ===== loaders
    if (p.Cin & 63) {
      // channel count padded to whole 64-channel slabs (V2ProPlus: 96): the pad columns are never written by the
      // staging loops and must read as zero (their weights are zero, but 0 x stale NaN is not)
      for (uint32_t o = (uint32_t)tid * 16u; o < a_bytes * (uint32_t)nA; o += PLD * 16u)
        *reinterpret_cast<uint4*>(sbase + o) = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("bar.sync 2, 128;" ::: "memory");        // the four loader warps only
    }
    PipeTile tl;
    int i = 0, ab = 0, sc = 0;                       // sc: A-buffer uses so far (tile stages)
    uint32_t a_phase = 0;
    const bool h16 = p.x16 != nullptr;
    const float pre = p.pre_slope;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      if (!pipe_tile(p, t, TX, NY, tl)) continue;
      const int tbase = tl.q0 + g.lo;
      for (int st = 0; st < S; ++st) {
        const int c0 = st * 128;
        const int nch = p.Cin - c0 < 128 ? p.Cin - c0 : 128;          // input channels of this stage
        const int cpr = h16 ? (nch >> 3) : (nch >> 2);                // 16-byte chunks per row
        const int rstep = PLD / cpr, fstep = PLD - rstep * cpr;
        const int rr0 = tid / cpr, f0 = tid - rr0 * cpr;
        uint8_t* abuf = sbase + (uint32_t)ab * a_bytes;
        if (sc >= nA) {                                              // the MMAs of use sc - nA have released the buffer
          ok = mbar_wait(&a_free[ab], a_phase ^ 1u) && ok;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (h16) {
          const __half* __restrict__ xh = p.x16 + (long long)tl.in0 * p.Cin + c0;
          const uint32_t abuf_u32 = sA + (uint32_t)ab * a_bytes;
          int rr = rr0, f = f0;
          while (rr < g.R) {
            const int tt = tbase + rr, c = f * 8;
            const uint32_t off = (uint32_t)(rr * 128 + (c & 63) * 2);
            const bool in = (unsigned)tt < (unsigned)tl.Tin;
            cp_async16(abuf_u32 + (uint32_t)(c >> 6) * g.slab_bytes + (off ^ (((off >> 7) & 7u) << 4)),
                       in ? (const void*)(xh + (long long)tt * p.Cin + c) : (const void*)xh, in ? 16 : 0);
            f += fstep; rr += rstep;
            if (f >= cpr) { f -= cpr; ++rr; }
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
          const float* __restrict__ xg = p.x + (long long)tl.in0 * p.ldx + c0;
          int rr = rr0, f = f0;
          while (rr < g.R) {
            float4 v[PUNR];
            const int rs = rr, fs = f;
#pragma unroll
            for (int k = 0; k < PUNR; ++k) {
              v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              const int tt = tbase + rr;
              if (rr < g.R && (unsigned)tt < (unsigned)tl.Tin)
                v[k] = __ldg(reinterpret_cast<const float4*>(xg + (long long)tt * p.ldx + f * 4));
              f += fstep; rr += rstep;
              if (f >= cpr) { f -= cpr; ++rr; }
            }
            int r2 = rs, f2 = fs;
#pragma unroll
            for (int k = 0; k < PUNR; ++k) {
              if (r2 < g.R) {
                const int c = f2 * 4;
                const uint32_t off = (uint32_t)(r2 * 128 + (c & 63) * 2);
                float4 a = v[k];
                a.x = fmaxf(a.x, a.x * pre); a.y = fmaxf(a.y, a.y * pre);
                a.z = fmaxf(a.z, a.z * pre); a.w = fmaxf(a.w, a.w * pre);
                const __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(abuf + (uint32_t)(c >> 6) * g.slab_bytes + (off ^ (((off >> 7) & 7u) << 4))) = pk;
              }
              f2 += fstep; r2 += rstep;
              if (f2 >= cpr) { f2 -= cpr; ++r2; }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
        mbar_arrive(&a_ready[ab]);
        ++sc;
        if (++ab == nA) { ab = 0; a_phase ^= 1u; }
      }
      if (cfg.use_res) {
        // residual rows of this tile -> staging halves (thread = row); the epilogue of tile i-1 must have left them
        const int q = tl.q0 + tid;
        const int lim = tl.nq < tl.Tout ? tl.nq : tl.Tout;
        const bool rv = q < lim;
        int nv = lim - tl.q0;
        nv = nv < 0 ? 0 : (nv > 128 ? 128 : nv);
        const float* rsrc = p.res + (long long)(tl.out0 + q) * p.ldr + tl.ny * BNT;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (i >= 1) ok = mbar_wait(&res_free[h], (uint32_t)((i - 1) & 1)) && ok;
          if (tid == 0) mbar_expect_tx(&res_ready[h], (uint32_t)nv * 256u);
          if (rv) bulk_g2s(base + oR + (uint32_t)h * RES_HALF + (uint32_t)tid * 256u, rsrc + h * 64, 256u, &res_ready[h]);
        }
      }
      ++i;
    }
    if (!ok && err_flag) atomicExch(err_flag, 1);
  } else {
    // ===== epilogue warps 4-7: TMEM lanes (warp - 4) * 32 ..., accumulator i & 1
    const int ew = warp - 4, etid = tid - PLD;
    const int rq = ew * 32;
    tc_epi::Args ea;
    ea.y = p.y; ea.res = p.res; ea.acc = p.accumulate ? p.y : nullptr;
    ea.ldy = p.ldy; ea.ldr = p.ldr;
    ea.act = p.act; ea.slope = p.act == ACT_RELU ? 0.f : p.act_slope; ea.oscale = p.out_scale;
    ea.Cout = p.Cout;
    ea.vec = tc_epi::vec_ok(p.y, p.ldy, p.res, p.ldr, p.Cout);
    ea.y16 = p.y16; ea.ldy16 = p.Cout;
    float* tile = reinterpret_cast<float*>(sbase + oE) + ew * tc_epi::TILE_FLOATS;
    const float* rstage = reinterpret_cast<const float*>(sbase + oR);
    PipeTile tl;
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      if (!pipe_tile(p, t, TX, NY, tl)) continue;
      const int b = i & 1;
      const int n0 = tl.ny * BNT;
      {
        const int q = tl.q0 + etid;
        const int to = q * p.out_mul + p.out_add;
        s_to[b][etid] = (q < tl.nq && to >= 0 && to < tl.Tout) ? to : -1;
        float bsum = 0.f;
        const int n = n0 + etid;
        if (n < p.Cout) {
          if (p.bias) bsum += p.bias[n];
          if (p.bias2) bsum += p.bias2[(long long)tl.seg * p.ldb2 + n];
        }
        s_bias[b][etid] = bsum;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps only
      const bool got = mbar_wait(&acc_ready[b], (uint32_t)((i >> 1) & 1));
      ok = got && ok;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        bool rgot = true;
        if (cfg.use_res) { rgot = mbar_wait(&res_ready[h], (uint32_t)(i & 1)); ok = rgot && ok; }
#pragma unroll 1
        for (int c0 = h * 64; c0 < h * 64 + 64; c0 += 32) {
          uint32_t v[32];
          tc_epi::tmem_load_chunk(tmem + ((uint32_t)rq << 16) + (uint32_t)(b * BNT + c0), true, v);
          if (c0 + 32 >= BNT) {                             // accumulator b is in registers: hand it back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_free[b]);
          }
          if (cfg.use_res) { ea.res_s = rstage + h * (RES_HALF / 4) + rq * 64 + (c0 - h * 64); ea.res_s_ld = 64; }
          if (got && rgot) tc_epi::store_chunk<32>(v, tile, s_bias[b] + c0, s_to[b] + rq, tl.out0, n0 + c0, ea, lane);
        }
        if (cfg.use_res) mbar_arrive(&res_free[h]);
      }
      ++i;
    }
    if (!ok && err_flag) atomicExch(err_flag, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * BNT)));
  }
}

int g_halo_bulk_max_cin = []() { const char* e = getenv("GENIE_HALO_BULK_MAX_CIN"); return e ? atoi(e) : 128; }();
// widest layer the persistent K-staged kernel takes (V2: 256, V2ProPlus: 192 / 384 / 768-input conv_pre stays generic)
int g_halo_pipe_max_cin = []() { const char* e = getenv("GENIE_HALO_PIPE_MAX_CIN"); return e ? atoi(e) : 384; }();

int halo_pipe_mode() {
  static int mode = [] { const char* e = getenv("GENIE_HALO_PIPE"); return e ? atoi(e) : 1; }();
  return mode;
}

int device_sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// persistent pipelined form; false = the operands do not fit (caller falls back to the one-tile kernel).
// GENIE_HALO_PIPE: 0 off, 1 (default) two CTAs per SM (single A buffer, residual from global memory: the SIMT loader /
// epilogue work needs the warps of two CTAs), 4 one fat CTA per SM (double A buffer, staged residual; +2: no staging)
bool launch_halo_pipe(const ConvGemm& p, const HaloGeom& g_in, int* err_flag, cudaStream_t s) {
  constexpr size_t EPI4 = 4 * tc_epi::TILE_FLOATS * 4;
  const int mode = halo_pipe_mode();
  const bool fat = (mode & 4) != 0;
  const size_t budget = fat ? 220 * 1024 : 109 * 1024;
  HaloGeom g = g_in;
  PipeCfg cfg{};
  cfg.slabs_total = g.slabs;
  cfg.stages = (g.slabs + 1) / 2;
  cfg.NY = (p.Cout + BNT - 1) / BNT;
  if (g.slabs > 2) g.slabs = 2;                              // the A buffer holds one stage: <= 2 slabs
  const size_t a_bytes = (size_t)g.slabs * g.slab_bytes;
  cfg.use_res = (fat && p.res && p.Cout % BNT == 0 && !p.y16 && p.out_mul == 1 && p.out_add == 0 && p.q_extra == 0 && (p.ldr & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(p.res) & 15) == 0 && tc_epi::vec_ok(p.y, p.ldy, p.res, p.ldr, p.Cout) &&
                 !(mode & 2)) ? 1 : 0;
  const size_t fixed = EPI4 + 1024 + (cfg.use_res ? 2 * RES_HALF : 0);
  static const int tries_fat[][2] = {{2, 4}, {2, 3}, {2, 2}, {1, 4}, {1, 3}, {1, 2}};
  static const int tries_two[][2] = {{1, 4}, {1, 3}, {1, 2}};
  for (int k = 0; k < (fat ? 6 : 3); ++k) {
    const int* tr = fat ? tries_fat[k] : tries_two[k];
    if ((size_t)tr[0] * a_bytes + (size_t)tr[1] * BTILE + fixed <= budget) { cfg.nA = tr[0]; cfg.nslot = tr[1]; break; }
  }
  if (cfg.nA == 0) return false;
  const size_t smem = (size_t)cfg.nA * a_bytes + (size_t)cfg.nslot * BTILE + fixed;
  const int nq = p.M + p.q_extra;
  cfg.TX = (nq + 127) / 128;
  const long long total = (long long)cfg.TX * cfg.NY * p.B;
  if (total <= 0 || total > 0x3fffffff) return false;
  cfg.total = (int)total;
  const int grid = (int)std::min<long long>(total, (long long)device_sm_count() * (fat ? 1 : 2));
  if (fat) {
    static DynSmemAttr attr;
    attr.ensure(tc_halo_pipe_kernel<24, 1>, smem);
    tc_halo_pipe_kernel<24, 1><<<grid, BTHR, smem, s>>>(p, g, cfg, err_flag);
  } else {
    static DynSmemAttr attr;
    attr.ensure(tc_halo_pipe_kernel<8, 2>, smem);
    tc_halo_pipe_kernel<8, 2><<<grid, BTHR, smem, s>>>(p, g, cfg, err_flag);
  }
  GENIE_LAUNCHED("tc_halo_pipe");
  return true;
}

bool launch_halo_bulk(const ConvGemm& p, int flags, int* err_flag, cudaStream_t s) {
  HaloGeom g;
  const int s_first = p.in_shift0, s_last = p.in_shift0 + (p.ntaps - 1) * p.in_shift_step;
  g.lo = s_first < s_last ? s_first : s_last;
  const int hi = s_first < s_last ? s_last : s_first;
  g.R = 128 + (hi - g.lo);
  g.slabs = (p.Cin + 63) / 64; g.cpad = g.slabs * 64;
  g.slab_bytes = (uint32_t)(((size_t)g.R * 128 + 1023) / 1024 * 1024);
  g.NU = p.ntaps * g.slabs;
  g.U = 1; g.NI = g.NU; g.slot_bytes = BTILE; g.flags = flags;
  if (halo_pipe_mode() && p.Cin <= g_halo_pipe_max_cin && launch_halo_pipe(p, g, err_flag, s)) return true;
  if (p.Cin > g_halo_bulk_max_cin || p.Cout % BNT != 0) return false;     // the one-tile kernel: Cin = Cout = 128 only
  size_t smem = (size_t)g.slabs * g.slab_bytes + (size_t)NSLOT * BTILE;
  if (smem < EPI_BYTES) smem = EPI_BYTES;
  smem += 1024;
  if (smem > 200 * 1024) return false;
  static DynSmemAttr attr;
  attr.ensure(tc_halo_bulk_kernel, smem);
  const int nq = p.M + p.q_extra;
  dim3 grid((nq + 127) / 128, p.Cout / BNT, p.B);
  tc_halo_bulk_kernel<<<grid, BTHR, smem, s>>>(p, g, err_flag);
  GENIE_LAUNCHED("tc_halo_bulk");
  return true;
}

template <int NT, int MT, int ROWB>
bool launch_halo(const ConvGemm& p, int flags, int* err_flag, cudaStream_t s) {
  HaloGeom g;
  const int s_first = p.in_shift0, s_last = p.in_shift0 + (p.ntaps - 1) * p.in_shift_step;
  g.lo = s_first < s_last ? s_first : s_last;
  const int hi = s_first < s_last ? s_last : s_first;
  g.R = MT * 128 + (hi - g.lo);
  g.cpad = p.Cin > 64 ? (p.Cin + 63) / 64 * 64 : ROWB / 2;
  g.slabs = g.cpad > 64 ? g.cpad / 64 : 1;
  g.slab_bytes = (uint32_t)(((size_t)g.R * ROWB + 1023) / 1024 * 1024);
  g.NU = p.ntaps * g.slabs;
  g.U = (NT <= 64 ? 8192 : 16384) / (NT * ROWB);   // ring slot: 8 KB where that buys a third / fourth CTA per SM
  if (g.U < 1) g.U = 1;
  if (g.U > g.NU) g.U = g.NU;
  g.NI = (g.NU + g.U - 1) / g.U;
  g.slot_bytes = (uint32_t)g.U * NT * ROWB;
  g.flags = flags;
  const int nslots = g.NI < NSLOT ? g.NI : NSLOT;
  size_t smem = (size_t)g.slabs * g.slab_bytes + (size_t)nslots * g.slot_bytes;
  if (smem < EPI_BYTES) smem = EPI_BYTES;
  smem += 1024;
  if (smem > 200 * 1024) return false;
  static DynSmemAttr attr;
  attr.ensure(tc_halo_conv_kernel<NT, MT, ROWB>, smem);
  const int nq = p.M + p.q_extra;
  dim3 grid((nq + MT * 128 - 1) / (MT * 128), (p.Cout + NT - 1) / NT, p.B);
  tc_halo_conv_kernel<NT, MT, ROWB><<<grid, NTHR, smem, s>>>(p, g, err_flag);
  GENIE_LAUNCHED("tc_halo_conv");
  return true;
}


// GENIE_TC_HALO: -1 disables the path, 1 forces wide layers onto the cp.async kernel; default 0 = on
int halo_mode() {
  static int mode = [] {
    const char* e = getenv("GENIE_TC_HALO");
    return e ? atoi(e) : 0;
  }();
  return mode;
}

}  // namespace

// a C -> C conv pair may hand its intermediate over in fp16 when both convs take tc_halo_conv_kernel
bool tc_halo_fp16_pair_ok(int C, int ntaps) {
  if (halo_mode() != 0 || ntaps < 2 || C % 8 != 0) return false;
  if (C == 16 || (C > 16 && C <= 64) || (C > 64 && C < 128)) return true;
  if (C % 64 != 0 || C < 128) return false;
  if (halo_pipe_mode() && C <= g_halo_pipe_max_cin) return true;                // tc_halo_pipe_kernel (any such C)
  return C <= g_halo_bulk_max_cin && C % BNT == 0;                              // tc_halo_bulk_kernel
}

// fp16 [Cout][kpad] -> pre-swizzled 128 x 64 tiles [Cout/128][tap][Cin/64] for tc_halo_bulk_kernel
bool pretile_w128_supported(int Cin, int Cout, int ntaps) { return Cin % 8 == 0 && Cin > 64 && Cout % 16 == 0 && ntaps >= 2; }
long long pretile_w128_halves(int Cin, int Cout, int ntaps) {
  return (long long)((Cout + BNT - 1) / BNT) * BNT * ntaps * ((Cin + 63) / 64) * 64;
}
void launch_pretile_w128(const __half* hi, int Cout, int kpad, int Cin, int ntaps, __half* tiles, cudaStream_t s) {
  const long long total = (long long)((Cout + BNT - 1) / BNT) * ntaps * ((Cin + 63) / 64) * BNT * 8;
  pretile_w128_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(hi, Cout, kpad, Cin, ntaps, tiles);
  GENIE_LAUNCHED("pretile_w128");
}

// k-tap convs whose activation halo fits in shared memory; false => caller uses tc_conv_gemm
bool try_launch_tc_halo_conv(const ConvGemm& p, int* err_flag, cudaStream_t s) {
  const int mode = halo_mode();
  if (mode < 0) { GENIE_CHECK(!p.x16 && !p.y16, "fp16 hand-over needs the halo conv kernel"); return false; }
  if (p.ntaps < 2 || p.tc_wlo != nullptr || p.tc_split_a || p.ksplit != 1 || p.tc_nt != 0) return false;
  if (p.pre_slope < 0.f || p.pre_slope > 1.f) return false;
  if (p.act != ACT_NONE && p.act != ACT_RELU && p.act != ACT_LRELU) return false;
  if (p.ldx % 4 != 0) return false;
  if (p.M + p.q_extra <= 0 || p.B <= 0) return true;
  if (p.Cin % 8 != 0) return false;
  if (p.Cin == 16 && p.Cout <= 16) return launch_halo<16, 4, 32>(p, mode, err_flag, s);
  if (p.Cin > 16 && p.Cin <= 32 && p.Cout <= 32) return launch_halo<32, 4, 64>(p, mode, err_flag, s);   // 24 (V2ProPlus), 32
  if (p.Cin > 32 && p.Cin < 64 && p.Cout <= 64) return launch_halo<64, 2, 128>(p, mode, err_flag, s);    // 48
  // 96 (V2ProPlus): as a two-slab layer with zero pad columns on the persistent kernel when its weights are pre-tiled
  static const bool padded_env = [] { const char* e = getenv("GENIE_HALO_PIPE_PADDED"); return !(e && e[0] == '0'); }();
  if (padded_env && p.Cin > 64 && p.Cin < 128 && p.tc_tiles != nullptr && halo_pipe_mode() && !(mode & 1) &&
      launch_halo_bulk(p, mode, err_flag, s))
    return true;
  if (p.Cin > 64 && p.Cin < 128 && p.Cout <= 128) return launch_halo<128, 1, 128>(p, mode, err_flag, s); // 96
  // wider layers: the per-tap gather of tc_gemm.cu at two CTAs per SM is faster than one halo CTA per SM
  // (measured: C=256 k=11 586 vs 1269 us, C=128 k=7 1067 vs 1194 us) unless forced for testing
  if (p.Cin % 64 == 0 && p.Cin >= 128 && p.tc_tiles != nullptr && !(mode & 1) &&
      ((halo_pipe_mode() && p.Cin <= g_halo_pipe_max_cin) || (p.Cin <= g_halo_bulk_max_cin && p.Cout % BNT == 0))) {
    if (launch_halo_bulk(p, mode, err_flag, s)) return true;
    GENIE_CHECK(!p.x16 && !p.y16, "fp16 hand-over needs the halo conv kernel");
    return false;
  }
  if (p.Cin % 64 != 0 || (p.Cin > 64 && !(mode & 1))) return false;
  if (p.Cout <= 64) return launch_halo<64, 2, 128>(p, mode, err_flag, s);
  return launch_halo<128, 1, 128>(p, mode, err_flag, s);
}

}  // namespace genie
