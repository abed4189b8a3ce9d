// Counter-based Philox4x32-10 -> N(0,1) (Box-Muller), shared by the T2S sampler
// and the vocoder's z_p noise.  Keyed by (seed), counter (idx, step, utt, stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace genie {

__device__ __forceinline__ uint2 philox_mulhilo(uint32_t a, uint32_t b) {
  unsigned long long p = (unsigned long long)a * b;
  return make_uint2((uint32_t)(p >> 32), (uint32_t)p);
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint2 a = philox_mulhilo(0xD2511F53u, ctr.x), b = philox_mulhilo(0xCD9E8D57u, ctr.z);
    ctr = make_uint4(b.x ^ ctr.y ^ key.x, b.y, a.x ^ ctr.w ^ key.y, a.y);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t utt, uint32_t step, uint32_t idx,
                                               uint32_t stream = 0u) {
  uint4 r = philox4x32_10(make_uint4(idx, step, utt, stream), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float u1 = ((float)(r.x >> 8) + 1.0f) * 5.9604644775390625e-08f;   // (0, 1]
  float u2 = (float)(r.y >> 8) * 5.9604644775390625e-08f;            // [0, 1)
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

}  // namespace genie
