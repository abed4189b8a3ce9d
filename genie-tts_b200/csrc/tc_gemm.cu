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

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_MISH: { float sp = v > 20.f ? v : log1pf(expf(v)); return v * tanhf(sp); }
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// byte offset of element (row r, 8-half chunk c8) inside a K-major SW128 tile
__device__ __forceinline__ uint32_t swz(int r, int c8) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4));
}

template <int NT, int SPLIT_A, int W_LO>
__global__ void __launch_bounds__(128) tc_conv_gemm_kernel(ConvGemm p, int* err_flag) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;

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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const uint32_t idesc = (1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const int Ktot = p.ntaps * p.Cin;
  const int KB = p.tc_kpad / BK;
  const float* __restrict__ xg = p.x + (long long)in0 * p.ldx;
  const float pre = p.pre_slope;
  const __half* __restrict__ whi = p.tc_w;
  const __half* __restrict__ wlo = p.tc_wlo;
  bool ok = true;

  const int kb_lo = (int)((long long)KB * ks / p.ksplit), kb_hi = (int)((long long)KB * (ks + 1) / p.ksplit);
  for (int kb = kb_lo; kb < kb_hi; ++kb) {
    const int it_k = kb - kb_lo;
    const int s = it_k & 1;
    uint8_t* st = sbase + (size_t)s * STAGE_BYTES;
    uint8_t* sA = st;
    uint8_t* sW = st + A_BYTES * SPLIT_A;
    // ---- global loads of the whole stage first (16 + NT/16 independent 16-byte loads per thread in
    // flight), then the buffer-free wait, then convert + swizzled stores: the load latency overlaps the
    // MMAs of the previous stages
    float4 av[16];
    // column group of this thread is the same for all 16 A loads: c4 = tid & 15, rows r = (tid>>4) + 8*it
    {
      const int c4 = tid & 15;
      const int kk = kb * BK + c4 * 4;
      const int tap = kk / p.Cin;
      const int ci = kk - tap * p.Cin;
      const int tsh = p.in_shift0 + tap * p.in_shift_step;
      const bool kok = kk < Ktot;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int r = (tid >> 4) + it * 8;
        const int t = q0 + r + tsh;
        av[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kok && (q0 + r) < nq && t >= 0 && t < Tin)
          av[it] = __ldg(reinterpret_cast<const float4*>(xg + (long long)t * p.ldx + ci));
      }
    }
    constexpr int WIT = NT / 16;                   // NT*8 uint4 per tile / 128 threads
    uint4 wv[WIT], wl[W_LO ? WIT : 1];
#pragma unroll
    for (int it = 0; it < WIT; ++it) {
      const int idx = tid + it * 128;
      const int n = idx >> 3, c8 = idx & 7;
      wv[it] = make_uint4(0u, 0u, 0u, 0u);
      if (W_LO) wl[it] = wv[it];
      if (n < n_mma && n0 + n < p.Cout) {
        const long long o = (long long)(n0 + n) * p.tc_kpad + kb * BK + c8 * 8;
        wv[it] = __ldg(reinterpret_cast<const uint4*>(whi + o));
        if (W_LO) wl[it] = __ldg(reinterpret_cast<const uint4*>(wlo + o));
      }
    }
    if (it_k >= 2) ok = mbar_wait(&bars[s], (uint32_t)(((it_k >> 1) - 1) & 1)) && ok;
    // ---- A tile: 128 rows x 64 k (fp32 -> fp16 hi/lo, swizzled)
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int r = (tid >> 4) + it * 8, c4 = tid & 15;
      float4 v = av[it];
      if (pre != 1.f) {
        v.x = v.x > 0.f ? v.x : v.x * pre; v.y = v.y > 0.f ? v.y : v.y * pre;
        v.z = v.z > 0.f ? v.z : v.z * pre; v.w = v.w > 0.f ? v.w : v.w * pre;
      }
      const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
      const uint32_t off = swz(r, c4 >> 1) + (uint32_t)(c4 & 1) * 8u;
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
      const int idx = tid + it * 128;
      const int n = idx >> 3, c8 = idx & 7;
      if (n < n_mma) {
        *reinterpret_cast<uint4*>(sW + swz(n, c8)) = wv[it];
        if (W_LO) *reinterpret_cast<uint4*>(sW + W_BYTES + swz(n, c8)) = wl[it];
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

  // ---- epilogue: TMEM lane = output row (warp w owns lanes 32w..32w+31).  Each 32x32 chunk goes
  // through a padded per-warp smem tile so that global stores / residual loads are row-contiguous
  // (one 128-byte line per warp instruction) instead of 32 strided rows.
  float* tile = reinterpret_cast<float*>(sbase) + warp * (32 * 33);
  const int row_base = q0 + warp * 32;
  const float* bias2 = p.bias2 ? p.bias2 + (long long)seg * p.ldb2 : nullptr;
  for (int c0 = 0; c0 < n_mma; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    if (n_mma - c0 >= 32) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
    } else {   // 16-column tail (n_mma is a multiple of 16)
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
#pragma unroll
      for (int j = 16; j < 32; ++j) v[j] = 0u;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]);
    __syncwarp();
    const int n = n0 + c0 + lane;                    // this lane's output column
    const bool col_ok = n < p.Cout && (c0 + lane) < n_mma;
    float badd = 0.f;
    if (col_ok && p.ksplit == 1) {
      if (p.bias) badd += p.bias[n];
      if (bias2) badd += bias2[n];
    }
    const bool plain = p.ksplit == 1;
    const bool has_res = plain && p.res != nullptr, has_acc = plain && p.accumulate;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 8) {
      float xv[8], rv[8], cv[8];
      long long orow[8];
      bool rok[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {                 // all loads of 8 rows are issued before any store
        const int q = row_base + r0 + i;
        const int to = q * p.out_mul + p.out_add;
        rok[i] = ok && col_ok && q < nq && to >= 0 && to < Tout;
        orow[i] = (long long)out0 + to;
        xv[i] = tile[(r0 + i) * 33 + lane];
        rv[i] = (rok[i] && has_res) ? p.res[orow[i] * p.ldr + n] : 0.f;
        cv[i] = (rok[i] && has_acc) ? p.y[orow[i] * p.ldy + n] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (!rok[i]) continue;
        float x = xv[i];
        if (plain) x = apply_act(x + badd, p.act, p.act_slope) * p.out_scale + rv[i] + cv[i];
        p.y[(long long)ks * p.split_stride + orow[i] * p.ldy + n] = x;
      }
    }
    __syncwarp();
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
  constexpr size_t smem = 2 * ((size_t)BM * 128 * SPLIT_A + (size_t)NT * 128 * (1 + W_LO)) + 1024;
  static bool configured = false;
  if (!configured) {
    GENIE_CUDA(cudaFuncSetAttribute(tc_conv_gemm_kernel<NT, SPLIT_A, W_LO>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int nq = p.M + p.q_extra;
  dim3 grid((nq + BM - 1) / BM, ((p.Cout + NT - 1) / NT) * p.ksplit, p.B);
  tc_conv_gemm_kernel<NT, SPLIT_A, W_LO><<<grid, 128, smem, s>>>(p, err_flag);
  GENIE_LAUNCHED("tc_conv_gemm");
}

template <int SPLIT_A, int W_LO>
void dispatch_nt(const ConvGemm& p, int* err_flag, cudaStream_t s) {
  // smallest tile that covers Cout in one CTA column; wide layers use 256 (or 128 when that tiles exactly)
  if (p.tc_nt == 32 || (p.tc_nt == 0 && p.Cout <= 32)) launch_tc<32, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.tc_nt == 64 || (p.tc_nt == 0 && p.Cout <= 64)) launch_tc<64, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.tc_nt == 128) launch_tc<128, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.tc_nt == 256) launch_tc<256, SPLIT_A, W_LO>(p, err_flag, s);
  else if (p.Cout <= 128 || (p.Cout % 256 != 0 && p.Cout % 128 == 0 && p.Cout > 256))
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
  const bool wlo = p.tc_wlo != nullptr;
  if (p.tc_split_a) {
    if (wlo) dispatch_nt<2, 1>(p, err_flag, s); else dispatch_nt<2, 0>(p, err_flag, s);
  } else {
    if (wlo) dispatch_nt<1, 1>(p, err_flag, s); else dispatch_nt<1, 0>(p, err_flag, s);
  }
}

}  // namespace genie
