// Unit self-test of the tcgen05 implicit-GEMM against the exact SIMT path on
// random data (no model needed).  Exposed through genie_debug_tc_selftest.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

namespace genie {

__global__ void selftest_pack_kernel(const float* __restrict__ w, int K, int Cout, int kpad, __half* hi, __half* lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * kpad) return;
  const int co = (int)(i / kpad), kk = (int)(i % kpad);
  const float v = kk < K ? w[(long long)co * K + kk] : 0.f;
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

// mode: 1 = x_hi.w_hi, 2 = (x_hi+x_lo).w_hi, 3 = + x_hi.w_lo.  Two ragged segments, dilated taps,
// pre-activation, bias, residual: the same parameter space the vocoder uses.
void tc_selftest(int M, int Cin, int Cout, int ntaps, int dil, int mode, int exact_w, float* max_err, float* ref_max) {
  cudaStream_t s = nullptr;
  std::mt19937 rng(1234u + M + 7 * Cin + 13 * Cout + 31 * ntaps);
  std::normal_distribution<float> nd(0.f, 1.f);
  const int K = ntaps * Cin, kpad = ((K + 63) / 64) * 64;
  const int M1 = M / 3, rows = M;
  std::vector<float> hx((size_t)rows * Cin), hw((size_t)Cout * K), hb(Cout), hr((size_t)rows * Cout);
  for (auto& v : hx) v = nd(rng);
  for (auto& v : hw) {
    v = nd(rng) / std::sqrt((float)K);
    if (exact_w) v = __half2float(__float2half_rn(v));
  }
  for (auto& v : hb) v = 0.1f * nd(rng);
  for (auto& v : hr) v = nd(rng);
  int hoff[3] = {0, M1, rows};
  float *x, *w, *b, *r, *y0, *y1; int* off; __half *hi, *lo; int* err;
  GENIE_CUDA(cudaMalloc(&x, hx.size() * 4)); GENIE_CUDA(cudaMalloc(&w, hw.size() * 4));
  GENIE_CUDA(cudaMalloc(&b, hb.size() * 4)); GENIE_CUDA(cudaMalloc(&r, hr.size() * 4));
  GENIE_CUDA(cudaMalloc(&y0, hr.size() * 4)); GENIE_CUDA(cudaMalloc(&y1, hr.size() * 4));
  GENIE_CUDA(cudaMalloc(&off, sizeof(hoff))); GENIE_CUDA(cudaMalloc(&err, 4));
  GENIE_CUDA(cudaMalloc(&hi, (size_t)Cout * kpad * 2)); GENIE_CUDA(cudaMalloc(&lo, (size_t)Cout * kpad * 2));
  GENIE_CUDA(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
  GENIE_CUDA(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  GENIE_CUDA(cudaMemcpy(b, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  GENIE_CUDA(cudaMemcpy(r, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
  GENIE_CUDA(cudaMemcpy(off, hoff, sizeof(hoff), cudaMemcpyHostToDevice));
  GENIE_CUDA(cudaMemset(err, 0, 4));
  GENIE_CUDA(cudaMemset(y0, 0, hr.size() * 4)); GENIE_CUDA(cudaMemset(y1, 0, hr.size() * 4));
  selftest_pack_kernel<<<(unsigned)(((long long)Cout * kpad + 255) / 256), 256>>>(w, K, Cout, kpad, hi, lo);

  ConvGemm p;
  p.x = x; p.ldx = Cin; p.w = w; p.w_f16 = 0; p.w_co_stride = K; p.w_tap_stride = Cin; p.bias = b;
  p.res = r; p.ldr = Cout; p.ldy = Cout; p.Cin = Cin; p.Cout = Cout; p.ntaps = ntaps; p.in_shift_step = dil;
  p.in_shift0 = -dil * (ntaps - 1) / 2; p.pre_slope = 0.1f; p.in_off = off; p.out_off = off; p.B = 2;
  p.M = std::max(M1, rows - M1); p.M_out = p.M;
  p.y = y0;
  launch_conv_gemm(p, s);
  __half* tiles = nullptr;
  if (pretile_w128_supported(Cin, Cout, ntaps)) {
    GENIE_CUDA(cudaMalloc(&tiles, (size_t)pretile_w128_halves(Cin, Cout, ntaps) * 2));
    launch_pretile_w128(hi, Cout, kpad, Cin, ntaps, tiles, s);
    p.tc_tiles = tiles;
  }
  p.y = y1; p.tc_w = hi; p.tc_wlo = mode >= 3 ? lo : nullptr; p.tc_kpad = kpad; p.tc_split_a = mode >= 2;
  launch_tc_conv_gemm(p, err, s);
  GENIE_CUDA(cudaDeviceSynchronize());
  if (getenv("GENIE_SELFTEST_TIME")) {
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    ConvGemm ps = p; ps.tc_w = nullptr; ps.y = y0;
    cudaEventRecord(e0, s);
    for (int i = 0; i < 10; ++i) launch_conv_gemm(ps, s);
    cudaEventRecord(e1, s);
    for (int i = 0; i < 10; ++i) launch_tc_conv_gemm(p, err, s);
    cudaEventRecord(e2, s);
    cudaDeviceSynchronize();
    float t0 = 0, t1 = 0;
    cudaEventElapsedTime(&t0, e0, e1); cudaEventElapsedTime(&t1, e1, e2);
    const double fl = 2.0 * rows * (double)K * Cout;
    fprintf(stderr, "  [time] simt %.1f us (%.1f TF/s)  tc %.1f us (%.1f TF/s)\n", t0 * 100, fl / (t0 * 1e-4) / 1e12,
            t1 * 100, fl / (t1 * 1e-4) / 1e12);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  }
  std::vector<float> a(hr.size()), c(hr.size());
  int herr = 0;
  GENIE_CUDA(cudaMemcpy(a.data(), y0, a.size() * 4, cudaMemcpyDeviceToHost));
  GENIE_CUDA(cudaMemcpy(c.data(), y1, c.size() * 4, cudaMemcpyDeviceToHost));
  GENIE_CUDA(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  float me = 0.f, rm = 0.f;
  for (size_t i = 0; i < a.size(); ++i) {
    float d = std::fabs(a[i] - c[i]);
    if (!(d <= me)) me = d;            // NaN-propagating max
    rm = std::max(rm, std::fabs(a[i]));
  }
  cudaFree(x); cudaFree(w); cudaFree(b); cudaFree(r); cudaFree(y0); cudaFree(y1); cudaFree(off); cudaFree(err);
  cudaFree(hi); cudaFree(lo); if (tiles) cudaFree(tiles);
  GENIE_CHECK(herr == 0, "tcgen05 pipeline timed out in selftest");
  *max_err = me; *ref_max = rm;
}

}  // namespace genie
