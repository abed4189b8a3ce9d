// tcgen05 implicit-GEMM for 1-D convolutions / linear layers (sm_100a).
//
//   D[128 rows (time / tokens), N<=256 (Cout)] += A[128, 64] * W[N, 64]^T   per k-block
//
// * accumulators live in TMEM (tcgen05.alloc, 128 lanes x N fp32 columns), read back with
//   tcgen05.ld for the fused epilogue (bias, per-utterance conditioning bias, activation,
//   residual, accumulate, transposed-conv phase scatter);
// * both operands are K-major 128B-swizzled fp16 tiles in shared memory, described by UMMA
//   shared-memory descriptors; the 4 loader warps gather the A tile straight from the fp32
//   channels-last activations (tap shift, zero padding at utterance edges and the
//   pre-activation leaky-relu are applied on the way) and split it into fp16 hi (+ lo) parts;
// * fp32 fidelity where token parity needs it: x = x_hi + x_lo with fp16-exact weights gives
//   (x_hi + x_lo) . w accumulated in fp32 (two MMAs per k-step), ~2^-22 relative;
// * 2-stage pipeline: MMAs of stage s (issued by one thread, completion signalled with
//   tcgen05.commit -> mbarrier) overlap the loads of stage s^1.
//
// Validated against the exact SIMT path (conv_gemm.cu) by genie_debug_tc_selftest and by the
// end-to-end parity tests.
#include "common.cuh"
#include "tc_epilogue.cuh"

#include <cstdlib>

namespace genie {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
// bounded wait (a wedged pipeline must not hang the GPU): false on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // layout: SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}

// byte offset of element (row r, 8-half chunk c8) inside a K-major SW128 tile
__device__ __forceinline__ uint32_t swz(int r, int c8) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4));
}


template <int NT, int SPLIT_A, int W_LO>
constexpr size_t tc_smem_bytes() { return 2 * ((size_t)BM * 128 * SPLIT_A + (size_t)NT * 128 * (1 + W_LO)) + 1024; }
// two CTAs per SM whenever the stage ring allows it: the k-loop of one overlaps the epilogue of the other
template <int NT, int SPLIT_A, int W_LO>
constexpr int tc_min_blocks() { return tc_smem_bytes<NT, SPLIT_A, W_LO>() <= 113 * 1024 ? 2 : 1; }

// 8 warps per CTA where two CTAs share an SM, 16 where the stage ring only leaves room for one: the loaders
// (fp32 -> fp16 hi/lo conversion) and the epilogue need the warps, the MMAs are issued by one thread either way
template <int NT, int SPLIT_A, int W_LO>
constexpr int tc_threads() { return tc_min_blocks<NT, SPLIT_A, W_LO>() == 2 ? 256 : 512; }

template <int NT, int SPLIT_A, int W_LO>
__global__ void __launch_bounds__(tc_threads<NT, SPLIT_A, W_LO>(), tc_min_blocks<NT, SPLIT_A, W_LO>())
tc_conv_gemm_kernel(ConvGemm p, int* err_flag) {
  constexpr int NTHR = tc_threads<NT, SPLIT_A, W_LO>();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_to[BM];                          // epilogue: output row of accumulator row (-1 = skip)
  __shared__ __align__(16) float s_bias[NT];

  const int seg = blockIdx.z;
  int in0 = 0, Tin = p.M, out0 = 0, Tout = p.M_out;
  if (p.in_off) { in0 = p.in_off[seg]; Tin = p.in_off[seg + 1] - in0; }
  if (p.out_off) { out0 = p.out_off[seg]; Tout = p.out_off[seg + 1] - out0; }
  const int nq = Tin + p.q_extra;
  const int q0 = blockIdx.x * BM;
  if (q0 >= nq) return;                      // uniform per CTA, before any allocation
  const int ntn = (p.Cout + NT - 1) / NT;
  const int n0 = ((int)blockIdx.y % ntn) * NT;
  const int ks = (int)blockIdx.y / ntn;             // split-K slice
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  constexpr uint32_t A_BYTES = BM * 128;            // 16 KB
  constexpr uint32_t W_BYTES = NT * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES * SPLIT_A + W_BYTES * (1 + W_LO);
  static_assert(2 * STAGE_BYTES >= (NTHR / 32) * tc_epi::TILE_FLOATS * 4, "epilogue tiles alias the stage ring");
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));

  int n_mma = p.Cout - n0;                          // columns this CTA really needs
  if (n_mma > NT) n_mma = NT;
  n_mma = (n_mma + 15) & ~15;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(NT < 32 ? 32 : NT)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // epilogue tables (row offsets, combined bias) are independent of the k-loop: fill them up front
  const bool plain = p.ksplit == 1;
  if (tid < BM) {
    const int q = q0 + tid;
    const int to = q * p.out_mul + p.out_add;
    const bool rok = q < nq && to >= 0 && to < Tout;
    s_to[tid] = rok ? to : -1;
  }
  for (int j = tid; j < NT; j += NTHR) {
    float bsum = 0.f;
    const int n = n0 + j;
    if (plain && n < p.Cout) {
      if (p.bias) bsum += p.bias[n];
      if (p.bias2) bsum += p.bias2[(long long)seg * p.ldb2 + n];
    }
    s_bias[j] = bsum;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // one MMA covers at most 256 columns: the 512-wide tile (split-fp16 linears: halves the number of CTAs that
  // re-convert the same A rows) issues two per k-step into adjacent TMEM column ranges
  const int n_a = n_mma > 256 ? 256 : n_mma, n_b = n_mma - n_a;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(n_a >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t idesc_b = (1u << 4) | ((uint32_t)(n_b >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const int Ktot = p.ntaps * p.Cin;
  const int KB = p.tc_kpad / BK;
  const float* __restrict__ xg = p.x + (long long)in0 * p.ldx;
  const float pre = p.pre_slope;                    // 0 <= pre <= 1: lrelu(v) == max(v, v * pre)
  const __half* __restrict__ whi = p.tc_w;
  const __half* __restrict__ wlo = p.tc_wlo;
  bool ok = true;

  // loader geometry: A tile = 128 rows x 16 float4; thread -> column group c4, rows ar0 + ARS * it
  constexpr int ARS = NTHR / 16, WRS = NTHR / 8;    // row steps (multiples of 8: the swizzle phase is kept)
  const int c4 = tid & 15, ar0 = tid >> 4;
  const uint32_t a_off0 = swz(ar0, c4 >> 1) + (uint32_t)(c4 & 1) * 8u;   // rows step by ARS -> +ARS*128 bytes
  constexpr int AIT = BM / ARS;                     // 8 (4)
  constexpr int WIT = (NT + WRS - 1) / WRS;         // uint4 per thread per W tile
  const int wn0 = tid >> 3, wc8 = tid & 7;          // W tile: rows wn0 + WRS * it, 16-byte chunk wc8
  const uint32_t w_off0 = swz(wn0, wc8);            // rows step by WRS -> +WRS*128 bytes

  const int kb_lo = (int)((long long)KB * ks / p.ksplit), kb_hi = (int)((long long)KB * (ks + 1) / p.ksplit);
  for (int kb = kb_lo; kb < kb_hi; ++kb) {
    const int it_k = kb - kb_lo;
    const int s = it_k & 1;
    uint8_t* st = sbase + (size_t)s * STAGE_BYTES;
    uint8_t* sA = st;
    uint8_t* sW = st + A_BYTES * SPLIT_A;
    // ---- all global loads of the stage first (independent 16-byte loads in flight), then the
    // buffer-free wait, then convert + swizzled stores; the other resident CTA covers the latency
    float4 av[AIT];
    constexpr int HIT = SPLIT_A == 2 ? BM * 8 / NTHR : 1;     // 16-byte chunks per thread per pre-split part
    uint4 hv[HIT], lv[HIT];
    const bool presplit = SPLIT_A == 2 && p.x16 != nullptr;   // A arrives as fp16 hi / lo rows (linear, 1 tap)
    if (presplit) {
#pragma unroll
      for (int it = 0; it < HIT; ++it) {
        const int ci = tid + it * NTHR;
        const int r = ci >> 3, c8 = ci & 7;
        hv[it] = make_uint4(0u, 0u, 0u, 0u); lv[it] = hv[it];
        const int t = q0 + r;
        if (t < nq && t < Tin && kb * BK + c8 * 8 < Ktot) {
          const long long o = (long long)(in0 + t) * p.Cin + kb * BK + c8 * 8;
          hv[it] = __ldg(reinterpret_cast<const uint4*>(p.x16 + o));
          lv[it] = __ldg(reinterpret_cast<const uint4*>(p.x16_lo + o));
        }
      }
    } else {
      const int kk = kb * BK + c4 * 4;
      const int tap = kk / p.Cin;
      const int ci = kk - tap * p.Cin;
      const int t0 = q0 + ar0 + p.in_shift0 + tap * p.in_shift_step;
      const bool kok = kk < Ktot;
      const float* __restrict__ xp = xg + (long long)t0 * p.ldx + ci;
      const long long rstep = (long long)ARS * p.ldx;
#pragma unroll
      for (int it = 0; it < AIT; ++it) {
        const int t = t0 + it * ARS;
        av[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kok && (q0 + ar0 + it * ARS) < nq && (unsigned)t < (unsigned)Tin)
          av[it] = __ldg(reinterpret_cast<const float4*>(xp + it * rstep));
      }
    }
    uint4 wv[WIT], wl[W_LO ? WIT : 1];
    {
      const __half* __restrict__ wp = whi + (long long)(n0 + wn0) * p.tc_kpad + kb * BK + wc8 * 8;
      const long long wdelta = wlo - whi;
#pragma unroll
      for (int it = 0; it < WIT; ++it) {
        const int n = wn0 + it * WRS;
        wv[it] = make_uint4(0u, 0u, 0u, 0u);
        if (W_LO) wl[it] = wv[it];
        if (n < n_mma && n0 + n < p.Cout) {
          wv[it] = __ldg(reinterpret_cast<const uint4*>(wp + (long long)it * WRS * p.tc_kpad));
          if (W_LO) wl[it] = __ldg(reinterpret_cast<const uint4*>(wp + (long long)it * WRS * p.tc_kpad + wdelta));
        }
      }
    }
    if (it_k >= 2) ok = mbar_wait(&bars[s], (uint32_t)(((it_k >> 1) - 1) & 1)) && ok;
    // ---- A tile: 128 rows x 64 k (fp32 -> fp16 hi/lo, swizzled; or a plain copy of the pre-split parts)
    if (presplit) {
#pragma unroll
      for (int it = 0; it < HIT; ++it) {
        const int ci = tid + it * NTHR;
        const uint32_t off = swz(ci >> 3, ci & 7);
        *reinterpret_cast<uint4*>(sA + off) = hv[it];
        *reinterpret_cast<uint4*>(sA + A_BYTES * (SPLIT_A - 1) + off) = lv[it];
      }
    }
#pragma unroll
    for (int it = 0; it < (presplit ? 0 : AIT); ++it) {
      float4 v = av[it];
      v.x = fmaxf(v.x, v.x * pre); v.y = fmaxf(v.y, v.y * pre);
      v.z = fmaxf(v.z, v.z * pre); v.w = fmaxf(v.w, v.w * pre);
      const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
      const uint32_t off = a_off0 + (uint32_t)it * (ARS * 128u);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h01);
      pk.y = *reinterpret_cast<const uint32_t*>(&h23);
      *reinterpret_cast<uint2*>(sA + off) = pk;
      if (SPLIT_A == 2) {
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y);
        const __half2 l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        pk.x = *reinterpret_cast<const uint32_t*>(&l01);
        pk.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(sA + A_BYTES + off) = pk;
      }
    }
    // ---- W tile: n_mma rows x 64 k (pre-packed fp16, zero padded in K)
#pragma unroll
    for (int it = 0; it < WIT; ++it) {
      if (wn0 + it * WRS < n_mma) {
        *reinterpret_cast<uint4*>(sW + w_off0 + (uint32_t)it * (WRS * 128u)) = wv[it];
        if (W_LO) *reinterpret_cast<uint4*>(sW + W_BYTES + w_off0 + (uint32_t)it * (WRS * 128u)) = wl[it];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aA = smem_u32(sA), aW = smem_u32(sW);
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {
        const uint64_t da = umma_desc(aA + j * 32), db = umma_desc(aW + j * 32);
        umma_f16(tmem, da, db, idesc, (uint32_t)((it_k | j) != 0));
        if (SPLIT_A == 2) umma_f16(tmem, umma_desc(aA + A_BYTES + j * 32), db, idesc, 1u);
        if (W_LO) umma_f16(tmem, da, umma_desc(aW + W_BYTES + j * 32), idesc, 1u);
        if (NT > 256 && n_b > 0) {
          const uint64_t db2 = umma_desc(aW + 256 * 128 + j * 32);
          umma_f16(tmem + 256u, da, db2, idesc_b, (uint32_t)((it_k | j) != 0));
          if (SPLIT_A == 2) umma_f16(tmem + 256u, umma_desc(aA + A_BYTES + j * 32), db2, idesc_b, 1u);
        }
      }
      // completion of everything issued so far -> frees this stage's buffers
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   ::"r"(smem_u32(&bars[s])) : "memory");
    }
  }
  {
    const int last = kb_hi - kb_lo - 1;
    if (last >= 0) ok = mbar_wait(&bars[last & 1], (uint32_t)((last >> 1) & 1)) && ok;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && err_flag) atomicExch(err_flag, 1);

  // ---- epilogue: TMEM lane = output row.  Warp w reads lane quarter (w & 3) and the 32-column chunks
  // c0 = 32 * (w >> 2) + 64 * i.  Kept lean on purpose (it touches NT/2 outputs per thread): row offsets and
  // the combined bias come from the shared tables; each 32x32 chunk is transposed through a padded
  // per-warp tile so that global stores / residual loads are one 128-byte line per warp instruction.
  tc_epi::Args ea;
  ea.y = p.y + (long long)ks * p.split_stride;
  ea.res = plain ? p.res : nullptr; ea.acc = plain && p.accumulate ? p.y : nullptr;
  ea.ldy = p.ldy; ea.ldr = p.ldr;
  ea.act = plain ? p.act : (int)ACT_NONE; ea.slope = p.act == ACT_RELU ? 0.f : p.act_slope;
  ea.oscale = plain ? p.out_scale : 1.f;
  ea.Cout = p.Cout;
  ea.vec = tc_epi::vec_ok(ea.y, p.ldy, ea.res, p.ldr, p.Cout);
  ea.y16 = plain ? p.y16 : nullptr; ea.y16_lo = p.y16_lo; ea.ldy16 = p.Cout;
  float* tile = reinterpret_cast<float*>(sbase) + warp * tc_epi::TILE_FLOATS;
  const int rq = (warp & 3) * 32;                               // first row of this warp's lane quarter
  if (ok) {
    for (int c0 = (warp >> 2) * 32; c0 < n_mma; c0 += NTHR / 4) {
      uint32_t v[32];
      const bool full = n_mma - c0 >= 32;
      tc_epi::tmem_load_chunk(tmem + ((uint32_t)rq << 16) + (uint32_t)c0, full, v);
      if (full) tc_epi::store_chunk<32>(v, tile, s_bias + c0, s_to + rq, out0, n0 + c0, ea, lane);
      else tc_epi::store_chunk<16>(v, tile, s_bias + c0, s_to + rq, out0, n0 + c0, ea, lane);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem), "r"((uint32_t)(NT < 32 ? 32 : NT)));
  }
}

template <int NT, int SPLIT_A, int W_LO>
void launch_tc(const ConvGemm& p, int* err_flag, cudaStream_t s) {
  constexpr size_t smem = tc_smem_bytes<NT, SPLIT_A, W_LO>();
  static DynSmemAttr attr;
  attr.ensure(tc_conv_gemm_kernel<NT, SPLIT_A, W_LO>, smem);
  const int nq = p.M + p.q_extra;
  dim3 grid((nq + BM - 1) / BM, ((p.Cout + NT - 1) / NT) * p.ksplit, p.B);
  tc_conv_gemm_kernel<NT, SPLIT_A, W_LO><<<grid, tc_threads<NT, SPLIT_A, W_LO>(), smem, s>>>(p, err_flag);
  GENIE_LAUNCHED("tc_conv_gemm");
}

template <int SPLIT_A, int W_LO>
void dispatch_nt(const ConvGemm& p, int* err_flag, cudaStream_t s) {
  // smallest tile that covers Cout in one CTA column; wide layers take the widest tile whose stage ring
  // still leaves room for two CTAs per SM (256 for single-pass fp16, 128 for the hi/lo split forms)
  // hi/lo-split activations: every N tile re-converts the same A rows, so the wide tile wins although only one
  // CTA fits per SM (measured at 24 200 rows: K=2048 N=512 293 -> 226 us, K=512 N=1536 221 -> 214 us)
  static const int nt_env = [] { const char* e = getenv("GENIE_TC_NT_SPLIT"); return e ? atoi(e) : 256; }();
  if constexpr (SPLIT_A == 2 && W_LO == 0) {
    // 512-wide tile (one CTA per SM, all 512 TMEM columns): N = 1536 / 2048 of the T2S linears
    static const int wide_env = [] { const char* e = getenv("GENIE_TC_NT512"); return e ? atoi(e) : 1; }();
    // (N = 512 itself stays on the 256-wide tile: one CTA per 128 rows leaves 190 CTAs for 148 SMs at the bench
    // size, 206 vs 190 us)
    if (wide_env && p.tc_nt == 0 && p.Cout % 512 == 0 && p.Cout >= 1024 && p.ksplit == 1) { launch_tc<512, 2, 0>(p, err_flag, s); return; }
  }
  if (SPLIT_A == 2 && !W_LO && nt_env == 256 && p.tc_nt == 0 && p.Cout > 128) { launch_tc<256, SPLIT_A, W_LO>(p, err_flag, s); return; }
  constexpr bool wide_ok = tc_min_blocks<256, SPLIT_A, W_LO>() == 2;
  constexpr bool mid_ok = tc_min_blocks<128, SPLIT_A, W_LO>() == 2;
  if (p.tc_nt == 32 || (p.tc_nt == 0 && p.Cout <= 32)) launch_tc<32, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.tc_nt == 64 || (p.tc_nt == 0 && p.Cout <= 64)) launch_tc<64, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.tc_nt == 128) launch_tc<128, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.tc_nt == 256) launch_tc<256, SPLIT_A, W_LO>(p, err_flag, s);
  else if (!mid_ok) launch_tc<64, SPLIT_A, W_LO>(p, err_flag, s);
  else if (!wide_ok || p.Cout <= 128 || (p.Cout % 256 != 0 && p.Cout % 128 == 0 && p.Cout > 256))
    launch_tc<128, SPLIT_A, W_LO>(p, err_flag, s);
  else launch_tc<256, SPLIT_A, W_LO>(p, err_flag, s);
}

}  // namespace

void launch_tc_conv_gemm(const ConvGemm& p, int* err_flag, cudaStream_t s) {
  GENIE_CHECK(p.tc_w != nullptr && p.tc_kpad % BK == 0, "tc_conv_gemm: weights not packed");
  GENIE_CHECK(p.Cin % 4 == 0 && p.ldx % 4 == 0, "tc_conv_gemm: Cin/ldx must be multiples of 4");
  GENIE_CHECK(p.ksplit >= 1 && p.ksplit <= p.tc_kpad / BK, "tc_conv_gemm: bad ksplit");
  const int nq = p.M + p.q_extra;
  if (nq <= 0 || p.B <= 0) return;
  if (try_launch_tc_halo_conv(p, err_flag, s)) return;
  GENIE_CHECK(!p.x16 || (p.x16_lo && p.tc_split_a && p.ntaps == 1 && p.Cin % 8 == 0 && !p.in_off),
              "tc_conv_gemm: a pre-split fp16 operand needs hi and lo parts, the split-A form and a plain linear layer");
  GENIE_CHECK(!p.y16 || (p.ksplit == 1 && !p.res && !p.accumulate && p.Cout % 4 == 0), "tc_conv_gemm: fp16 output is exclusive");
  const bool wlo = p.tc_wlo != nullptr;
  if (p.tc_split_a) {
    if (wlo) dispatch_nt<2, 1>(p, err_flag, s); else dispatch_nt<2, 0>(p, err_flag, s);
  } else {
    if (wlo) dispatch_nt<1, 1>(p, err_flag, s); else dispatch_nt<1, 0>(p, err_flag, s);
  }
}

}  // namespace genie
