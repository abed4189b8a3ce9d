// Shared epilogue of the tcgen05 conv / linear kernels (tc_gemm.cu, tc_halo_conv.cu).
//
// A warp owns a 32-row x 32-column chunk of the accumulator: after tcgen05.ld thread `lane` holds row
// `lane`, 32 consecutive columns.  Bias and activation are applied in registers, the chunk is transposed
// through a padded per-warp tile (stride 36 floats: both the row-wise 16-byte stores and the 8-lanes-per-row
// 16-byte loads are bank-conflict free), and the rows go to global memory as 16-byte accesses, 8 lanes per
// 128-byte row segment.  All residual / accumulate loads of the chunk (8 x 16 B per thread) are issued before
// the first store: the epilogue is pure HBM traffic, so the bytes in flight per warp decide its speed (the
// first version, one 4-byte load per lane and row, ran at ~1 TB/s and was 60% of the kernel).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace genie {
namespace tc_epi {

constexpr int TILE_LD = 36;
constexpr int TILE_FLOATS = 32 * TILE_LD;           // per warp
constexpr int TILE_LD16 = 20;                       // 16-column chunks: 80-byte rows (also conflict free), 2.5 KB per warp
constexpr int TILE_FLOATS16 = 32 * TILE_LD16;

struct Args {
  float* y; const float* res; const float* acc;     // acc = y when accumulating, else null
  int ldy, ldr;
  int act; float slope, oscale;                     // ACT_NONE or (leaky-)ReLU as max(x, x * slope)
  int Cout;
  bool vec;                                         // 16-byte path legal (alignment, Cout % 4 == 0)
  __half* y16 = nullptr; int ldy16 = 0;             // fp16 output instead of y (no residual / accumulate): the
                                                    // consumer conv reads it as its ready-made A operand
  __half* y16_lo = nullptr;                         // + fp16(v - fp16(v)) for the hi/lo-split consumers
  // residual rows already staged in shared memory (tc_halo_pipe_kernel): the chunk's 32 rows x 32 columns start at
  // res_s, rows res_s_ld floats apart; replaces the global loads of `res`
  const float* res_s = nullptr; int res_s_ld = 0;
};

__host__ __device__ __forceinline__ bool vec_ok(const float* y, int ldy, const float* res, int ldr, int Cout) {
  return ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(res)) & 15) == 0 && (ldy & 3) == 0 &&
         (ldr & 3) == 0 && (Cout & 3) == 0;
}

// rows[r]: output row (relative to out0) of chunk row r, or -1; ncol0: global column of the chunk's first
// column; sbias: bias of that column onwards (shared memory)
template <int CW>
__device__ __forceinline__ void store_chunk(const uint32_t (&v)[32], float* tile, const float* sbias, const int* rows,
                                            long long out0, int ncol0, const Args& a, int lane) {
  static_assert(CW == 16 || CW == 32, "chunk width");
  constexpr int LD = CW == 16 ? TILE_LD16 : TILE_LD;
#pragma unroll
  for (int k = 0; k < CW / 4; ++k) {
    float4 o;
    o.x = __uint_as_float(v[4 * k + 0]) + sbias[4 * k + 0];
    o.y = __uint_as_float(v[4 * k + 1]) + sbias[4 * k + 1];
    o.z = __uint_as_float(v[4 * k + 2]) + sbias[4 * k + 2];
    o.w = __uint_as_float(v[4 * k + 3]) + sbias[4 * k + 3];
    if (a.act != ACT_NONE) {
      o.x = fmaxf(o.x, o.x * a.slope); o.y = fmaxf(o.y, o.y * a.slope);
      o.z = fmaxf(o.z, o.z * a.slope); o.w = fmaxf(o.w, o.w * a.slope);
    }
    o.x *= a.oscale; o.y *= a.oscale; o.z *= a.oscale; o.w *= a.oscale;
    *reinterpret_cast<float4*>(tile + lane * LD + 4 * k) = o;
  }
  __syncwarp();
  if (a.y16) {
    constexpr int CPR = CW / 4, RPP = 32 / CPR, NP = 32 / RPP;
    const int rsub = lane / CPR, c4 = lane % CPR;
    const int n = ncol0 + c4 * 4;
    if (n < a.Cout) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int to = rows[i * RPP + rsub];
        if (to < 0) continue;
        const float4 t = *reinterpret_cast<const float4*>(tile + (i * RPP + rsub) * LD + c4 * 4);
        const __half2 h01 = __floats2half2_rn(t.x, t.y), h23 = __floats2half2_rn(t.z, t.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h01);
        pk.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(a.y16 + (out0 + to) * a.ldy16 + n) = pk;
        if (a.y16_lo) {
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(t.x - f01.x, t.y - f01.y);
          const __half2 l23 = __floats2half2_rn(t.z - f23.x, t.w - f23.y);
          pk.x = *reinterpret_cast<const uint32_t*>(&l01);
          pk.y = *reinterpret_cast<const uint32_t*>(&l23);
          *reinterpret_cast<uint2*>(a.y16_lo + (out0 + to) * a.ldy16 + n) = pk;
        }
      }
    }
  } else if (a.vec) {
    constexpr int CPR = CW / 4;                     // lanes per row
    constexpr int RPP = 32 / CPR;                   // rows per pass
    constexpr int NP = 32 / RPP;                    // passes
    const int rsub = lane / CPR, c4 = lane % CPR;
    const int n = ncol0 + c4 * 4;
    const bool cok = n < a.Cout;
    float4 add[NP];
    int to[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      to[i] = cok ? rows[i * RPP + rsub] : -1;
      add[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (to[i] >= 0 && a.res_s)
        add[i] = *reinterpret_cast<const float4*>(a.res_s + (i * RPP + rsub) * a.res_s_ld + c4 * 4);
      else if (to[i] >= 0 && a.res)
        add[i] = __ldg(reinterpret_cast<const float4*>(a.res + (out0 + to[i]) * a.ldr + n));
    }
    if (a.acc) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        if (to[i] >= 0) {
          const float4 t = *reinterpret_cast<const float4*>(a.acc + (out0 + to[i]) * a.ldy + n);
          add[i].x += t.x; add[i].y += t.y; add[i].z += t.z; add[i].w += t.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (to[i] >= 0) {
        float4 t = *reinterpret_cast<const float4*>(tile + (i * RPP + rsub) * LD + c4 * 4);
        t.x += add[i].x; t.y += add[i].y; t.z += add[i].z; t.w += add[i].w;
        *reinterpret_cast<float4*>(a.y + (out0 + to[i]) * a.ldy + n) = t;
      }
    }
  } else {                                          // scalar fallback: lane = column
    const int n = ncol0 + lane;
    if (lane < CW && n < a.Cout) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int to = rows[r];
        if (to < 0) continue;                       // warp-uniform
        float x = tile[r * LD + lane];
        if (a.res) x += a.res[(out0 + to) * a.ldr + n];
        if (a.acc) x += a.acc[(out0 + to) * a.ldy + n];
        a.y[(out0 + to) * a.ldy + n] = x;
      }
    }
  }
  __syncwarp();
}

// accumulator chunk (this warp's 32 TMEM lanes x 32 or 16 columns) -> registers
__device__ __forceinline__ void tmem_load_chunk(uint32_t taddr, bool full, uint32_t (&v)[32]) {
  if (full) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  } else {   // 16-column tail (the MMA N is a multiple of 16)
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
}

}  // namespace tc_epi
}  // namespace genie
