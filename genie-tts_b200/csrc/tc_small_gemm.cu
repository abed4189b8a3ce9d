// Single-shot tcgen05 GEMM for the decode step of a batch (8 < rows <= 128):
//   P[ks][rows, N-tile] = act_in(sum_s X_s + a_bias)[rows, K-slice ks] . W[N-tile, K-slice ks]^T
// One CTA owns a 256-wide K slice and an NT-wide N tile.  The whole slice of both operands is
// brought into shared memory at once (every global load of the CTA is in flight before the first
// store), one barrier, 32 back-to-back MMAs (x_hi and x_lo against the fp16-exact weights), one
// commit: the latency chain is launch -> loads -> MMA -> read-out instead of 8 dependent k-blocks.
// Outputs are raw split-K partials; the consumers sum them and apply bias / residual / activation:
// LayerNorm (elementwise.cu), the fused decode attention (attention.cu) or this kernel's own A-operand loader (FFN2 reads
// relu(sum FFN1 partials + bias)).
#include "kernels.cuh"
#include <cstdlib>

namespace genie {
namespace {

constexpr int KS = 256;                       // K slice per CTA
constexpr int NKB = KS / 64;                  // 4 swizzle-128B k-blocks

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint32_t swz(int r, int c8) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4));
}

// THR = 512: rows 0..127 (batch <= 128).  THR = 256: rows 0..63 only (batch <= 64; operand rows 64..127 stay
// uninitialised, their accumulator rows are never read) with half the register file, so that a bandwidth-bound
// kernel of the other half-batch (decode attention) can share the SM.
template <int NT, int THR>
__global__ void __launch_bounds__(THR) tc_small_gemm_kernel(SmallGemm p, int* err_flag) {
  constexpr int NTHR = THR, RSTEP = THR / 16;
  // THR = 256 stores 64 operand rows per k-block: the 128-row MMA then reads rows 64..127 from the next
  // k-block / the weight tiles (in bounds, finite or not: those accumulator rows are never read), and two
  // such CTAs (81 KB, half the registers each) fit on an SM
  constexpr uint32_t ABYTES = THR == 256 ? 8192u : 16384u;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * NT, ks = blockIdx.y, k0 = ks * KS;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sA = smem_raw + (base - smem_u32(smem_raw));     // [hi|lo][NKB][128 x 128 B]
  uint8_t* sAlo = sA + NKB * ABYTES;
  uint8_t* sW = sAlo + NKB * ABYTES;                          // [NKB][NT x 128 B]

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(NT < 32 ? 32 : NT)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- weights: NT rows x 256 k (fp16 as stored), all loads first
  constexpr int WIT = (NT * 32 + NTHR - 1) / NTHR;          // uint4 per thread
  uint4 wv[WIT];
#pragma unroll
  for (int it = 0; it < WIT; ++it) {
    const int idx = tid + it * NTHR;
    const int n = idx >> 5, c = idx & 31;       // c: 8-half chunk within the 256-wide slice
    wv[it] = make_uint4(0u, 0u, 0u, 0u);
    if (n < NT && n0 + n < p.N) wv[it] = __ldg(reinterpret_cast<const uint4*>(p.w + (long long)(n0 + n) * p.ldw + k0 + c * 8));
  }
  // ---- activations: 128 rows x 256 k fp32 (sum of a_nsplit partials + bias, optional relu) -> fp16 hi/lo.
  // The kernel is one dependent chain (launch -> loads -> MMA -> read-out), so the whole slice of a
  // partial is in flight at once: 16 float4 per thread (row r0 + 32*i, k-block kb, i = j >> 2, kb = j & 3)
  // = one L2 round trip per partial instead of one per k-block.
  const int c4 = tid & 15;                      // float4 column within a 64-wide k-block
  const int r0 = tid >> 4;                      // 0..31
  // weights above are constant; activations come from the predecessor, which may still have been running when
  // this CTA started: they are read after the wait and through L2 (never the non-coherent path)
  pdl_wait();
  float4 av[16];
  {
    const float* colp = p.x + k0 + c4 * 4;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int r = r0 + (j >> 2) * RSTEP;
      av[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.M) av[j] = __ldcg(reinterpret_cast<const float4*>(colp + (long long)r * p.ldx + (j & 3) * 64));
    }
#pragma unroll 1
    for (int sp = 1; sp < p.a_nsplit; ++sp) {
      const float* cs = colp + sp * p.a_stride;
      float4 tv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = r0 + (j >> 2) * RSTEP;
        tv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.M) tv[j] = __ldcg(reinterpret_cast<const float4*>(cs + (long long)r * p.ldx + (j & 3) * 64));
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { av[j].x += tv[j].x; av[j].y += tv[j].y; av[j].z += tv[j].z; av[j].w += tv[j].w; }
    }
  }
  // this kernel runs once per CTA, so straight-line code size (cold instruction fetch) matters: the conversion
  // is unrolled over the 16 held values only (an earlier, fully unrolled loader cost 21 us per launch)
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    const float4 bb = p.a_bias ? __ldg(reinterpret_cast<const float4*>(p.a_bias + k0 + kb * 64 + c4 * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + i * RSTEP;
      float4 v = av[i * 4 + kb];
      if (r < p.M) {
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        if (p.a_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y);
      const __half2 l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
      const uint32_t off = (uint32_t)kb * ABYTES + swz(r, c4 >> 1) + (uint32_t)(c4 & 1) * 8u;
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h01); pk.y = *reinterpret_cast<const uint32_t*>(&h23);
      *reinterpret_cast<uint2*>(sA + off) = pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&l01); pk.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(sAlo + off) = pk;
    }
  }
#pragma unroll
  for (int it = 0; it < WIT; ++it) {
    const int idx = tid + it * NTHR;
    const int n = idx >> 5, c = idx & 31;
    if (n < NT) *reinterpret_cast<uint4*>(sW + (c >> 3) * (NT * 128) + swz(n, c & 7)) = wv[it];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t aA = smem_u32(sA), aL = smem_u32(sAlo), aW = smem_u32(sW);
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t db = umma_desc(aW + kb * (NT * 128) + j * 32);
        umma_f16(tmem, umma_desc(aA + kb * ABYTES + j * 32), db, idesc, (uint32_t)((kb | j) != 0));
        umma_f16(tmem, umma_desc(aL + kb * ABYTES + j * 32), db, idesc, 1u);
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(&bar)) : "memory");
  }
  // ---- wait for the accumulator (bounded: a wedged pipeline must not hang the GPU)
  bool ok = false;
  for (uint32_t i = 0; i < (1u << 22) && !ok; ++i) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    ok = done != 0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && err_flag) atomicExch(err_flag, 1);

  // ---- read-out: warps 0..3 own TMEM lanes 32w..32w+31 (= output rows); transposed through smem so
  // that each warp instruction stores one row's 32 consecutive floats
  if (warp < 4) {
    float* tile = reinterpret_cast<float*>(sA) + warp * (32 * 33);    // operand buffers are free now
    float* yb = p.y + (long long)ks * p.split_stride;
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int n = n0 + c0 + lane;
      if (ok && n < p.N) {
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int row = warp * 32 + r;
          if (row < p.M) yb[(long long)row * p.ldy + n] = tile[r * 33 + lane];
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem), "r"((uint32_t)(NT < 32 ? 32 : NT)));
  }
}

template <int NT, int THR>
void launch_nt(const SmallGemm& p, int* err_flag, cudaStream_t s) {
  constexpr size_t smem = 2 * NKB * (THR == 256 ? 8192 : 16384) + (size_t)NKB * NT * 128 + 1024;
  static DynSmemAttr attr;
  attr.ensure(tc_small_gemm_kernel<NT, THR>, smem);
  launch_pdl(tc_small_gemm_kernel<NT, THR>, dim3((p.N + NT - 1) / NT, p.K / KS), dim3(THR), smem, s, p, err_flag);
  GENIE_LAUNCHED("tc_small_gemm");
}

}  // namespace

void launch_tc_small_gemm(const SmallGemm& p, int nt, int* err_flag, cudaStream_t s) {
  GENIE_CHECK(p.M >= 1 && p.M <= 128 && p.K % KS == 0 && p.ldx % 4 == 0 && p.ldw % 8 == 0, "tc_small_gemm: bad shape");
  // experiment knob: GENIE_SMALL_NT=64 doubles the N tile of the <= 64-row variant (half the CTAs per GEMM: less
  // pressure on the 2-CTA-per-SM shared-memory slots when several decode chains run side by side)
  static const int env_nt = [] { const char* e = getenv("GENIE_SMALL_NT"); return e ? atoi(e) : 0; }();
  if (nt == 64) launch_nt<64, 512>(p, err_flag, s);
  else if (p.M <= 64 && env_nt == 64 && p.N % 64 == 0) launch_nt<64, 256>(p, err_flag, s);
  else if (p.M <= 64) launch_nt<32, 256>(p, err_flag, s);
  else launch_nt<32, 512>(p, err_flag, s);
}


}  // namespace genie
