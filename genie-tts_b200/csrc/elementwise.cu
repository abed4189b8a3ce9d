// Small fused / elementwise kernels around the GEMM, conv and attention kernels.
#include "kernels.cuh"
#include <math_constants.h>

namespace genie {
namespace {

inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

// ---- LayerNorm (+residual), one warp per row --------------------------------
// x may be `nsplit` split-K partial sums `split_stride` floats apart; `lin_bias` is the producing
// linear layer's bias (deferred from the split-K GEMM epilogue)
template <int N>   // N = C / 32 values per lane, all in registers
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                 const float* __restrict__ g, const float* __restrict__ b,
                                 float* __restrict__ y, int rows, int nsplit, long long split_stride,
                                 const float* __restrict__ lin_bias) {
  constexpr int C = N * 32;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * C + lane;
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = xr[32 * i];
  // split-K partials: two slices per pass so that 2N loads are in flight before the adds
  int sp = 1;
  for (; sp + 1 < nsplit; sp += 2) {
    const float* xs = xr + sp * split_stride;
    const float* xt = xs + split_stride;
    float a[N], c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = xs[32 * i]; c[i] = xt[32 * i]; }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += a[i] + c[i];
  }
  for (; sp < nsplit; ++sp) {
    const float* xs = xr + sp * split_stride;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += xs[32 * i];
  }
  if (lin_bias) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += lin_bias[lane + 32 * i];
  }
  if (res) {
    const float* rr = res + (long long)row * C + lane;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += rr[32 * i];
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) sum += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) { float d = v[i] - mean; var = fmaf(d, d, var); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = 1.f / sqrtf(var / (float)C + 1e-5f);
  float* yr = y + (long long)row * C + lane;
#pragma unroll
  for (int i = 0; i < N; ++i) yr[32 * i] = (v[i] - mean) * rstd * g[lane + 32 * i] + b[lane + 32 * i];
}

// Decode-step LayerNorm (few rows, C = 512): one 128-thread CTA per row, one float4 per thread.  Every
// load of the row (up to 8 split-K partials, residual, deferred bias, gamma, beta) is issued before the
// first add, so the kernel is one L2 round trip + two block reductions instead of a chain of
// partial-pair round trips per warp (8.8 -> ~3 us for the 8-partial FFN2 output at 100 rows).
__global__ void __launch_bounds__(128) layernorm_row512_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                               const float* __restrict__ g, const float* __restrict__ b,
                                                               float* __restrict__ y, int nsplit, long long split_stride,
                                                               const float* __restrict__ lin_bias) {
  constexpr int C = 512;
  const int row = blockIdx.x, tid = threadIdx.x;
  pdl_trigger();
  const float* xr = x + (long long)row * C + tid * 4;
  float4 lb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lin_bias) lb = __ldg(reinterpret_cast<const float4*>(lin_bias + tid * 4));
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g + tid * 4));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(b + tid * 4));
  pdl_wait();
  float4 part[8];
#pragma unroll
  for (int sp = 0; sp < 8; ++sp) {
    part[sp] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sp < nsplit) part[sp] = __ldcg(reinterpret_cast<const float4*>(xr + sp * split_stride));
  }
  float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
  if (res) rr = __ldcg(reinterpret_cast<const float4*>(res + (long long)row * C + tid * 4));
  float4 v = part[0];
#pragma unroll
  for (int sp = 1; sp < 8; ++sp) { v.x += part[sp].x; v.y += part[sp].y; v.z += part[sp].z; v.w += part[sp].w; }
  v.x += lb.x + rr.x; v.y += lb.y + rr.y; v.z += lb.z + rr.z; v.w += lb.w + rr.w;

  __shared__ float red[2][4];
  float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((tid & 31) == 0) red[0][tid >> 5] = sum;
  __syncthreads();
  const float mean = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) / (float)C;
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  float var = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  if ((tid & 31) == 0) red[1][tid >> 5] = var;
  __syncthreads();
  const float rstd = 1.f / sqrtf(((red[1][0] + red[1][1]) + (red[1][2] + red[1][3])) / (float)C + 1e-5f);
  float4 o4;
  o4.x = dx * rstd * gg.x + bb.x; o4.y = dy * rstd * gg.y + bb.y;
  o4.z = dz * rstd * gg.z + bb.z; o4.w = dw * rstd * gg.w + bb.w;
  *reinterpret_cast<float4*>(y + (long long)row * C + tid * 4) = o4;
}

__device__ __forceinline__ float pe_value(int pos, int c, const float* div_term) {
  // interleaved sin/cos (t2s_encoder#[71-79]): even c -> sin(pos*div[c/2]), odd -> cos
  float ang = (float)pos * div_term[c >> 1];
  return (c & 1) ? cosf(ang) : sinf(ang);
}

__global__ void text_embed_pe_kernel(float* x, const long long* seq, const int* pos, const float* emb,
                                     const float* alpha, const float* div_term, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * 512) return;
  int r = (int)(i >> 9), c = (int)(i & 511);
  float v = emb[seq[r] * 512 + c] + x[i];
  x[i] = v + alpha[0] * pe_value(pos[r], c, div_term);
}

__global__ void audio_embed_pe_kernel(float* out, const int* tok, const int* pos, const float* emb,
                                      const float* alpha, const float* div_term, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * 512) return;
  int r = (int)(i >> 9), c = (int)(i & 511);
  out[i] = emb[(long long)tok[r] * 512 + c] + alpha[0] * pe_value(pos[r], c, div_term);
}

__global__ void decode_embed_kernel(float* out, const int* hist, int hist_ld, const int* hist_len,
                                    const int* active, const float* emb, const float* alpha,
                                    const float* div_term, int prompt_len_is_in_hist) {
  pdl_trigger();   // launched normally (full barrier after the sampler); lets the QKV GEMM start its weight loads
  const int b = blockIdx.x, c = threadIdx.x;   // 512 threads
  if (active && !active[b]) return;
  const int n = hist_len[b];
  const int tok = hist[(long long)b * hist_ld + n - 1];
  // the newest token sits at 1-based audio position n (stage#[15-33]: positions 1..len(y_emb))
  out[(long long)b * 512 + c] = emb[(long long)tok * 512 + c] + alpha[0] * pe_value(n, c, div_term);
}

template <typename KVT>
__global__ void kv_scatter_kernel(const float* __restrict__ qkv, int ld, KVT* __restrict__ kv_base,
                                  long long utt_stride, long long layer_off, long long v_off, int cap,
                                  const int* __restrict__ row_off, const int* __restrict__ dst_pos0,
                                  const int* __restrict__ row2utt, int rows, const int* __restrict__ active,
                                  const int* __restrict__ slot_of) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over rows*1024/4 float4s
  if (i >= (long long)rows * 256) return;
  int r = (int)(i >> 8), c4 = (int)(i & 255);       // c4: float4 index within [K(512) | V(512)]
  int b = row2utt ? row2utt[r] : r;
  if (active && !active[b]) return;
  int pos = dst_pos0[b] + (row_off ? r - row_off[b] : 0);
  if (slot_of) b = slot_of[b];
  int isv = c4 >> 7, col = (c4 & 127) * 4;          // col within 512
  int h = col >> 5, e = col & 31;
  float4 v = *reinterpret_cast<const float4*>(qkv + (long long)r * ld + 512 + isv * 512 + col);
  KVT* dst = kv_base + (long long)b * utt_stride + layer_off + (isv ? v_off : 0) +
             ((long long)h * cap + pos) * 32 + e;
  if constexpr (sizeof(KVT) == 2) {
    uint2 u;
    __half2* hp = reinterpret_cast<__half2*>(&u);
    hp[0] = __floats2half2_rn(v.x, v.y); hp[1] = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(dst) = u;
  } else {
    *reinterpret_cast<float4*>(dst) = v;
  }
}

__global__ void __launch_bounds__(256) prefill_bert_gather_kernel(float* __restrict__ bert, PrefillMeta pm,
                                                                 const float* __restrict__ text_bert) {
  const int r = blockIdx.x, b = pm.row2utt[r], i = r - pm.row_off[b];
  if (i >= pm.lx[b]) return;                                   // audio row
  const int lr = pm.lr[b];
  const float* src = nullptr;
  if (i < lr) { if (pm.ref_bert[b]) src = pm.ref_bert[b] + (long long)i * 1024; }
  else if (text_bert) src = text_bert + (long long)(pm.txt_in_off[b] + i - lr) * 1024;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (src) v = *reinterpret_cast<const float4*>(src + threadIdx.x * 4);
  *reinterpret_cast<float4*>(bert + (long long)(pm.txt_off[b] + i) * 1024 + threadIdx.x * 4) = v;
}

__global__ void __launch_bounds__(128) prefill_embed_kernel(float* __restrict__ x, PrefillMeta pm,
                                                            const long long* __restrict__ text_seq,
                                                            const float* __restrict__ proj,
                                                            const float* __restrict__ proj_bias,
                                                            const float* __restrict__ text_emb,
                                                            const float* __restrict__ text_alpha, int text_vocab,
                                                            const float* __restrict__ audio_emb,
                                                            const float* __restrict__ audio_alpha,
                                                            const float* __restrict__ div_term, int* err) {
  const int r = blockIdx.x, b = pm.row2utt[r], i = r - pm.row_off[b];
  const int lx = pm.lx[b];
  float* out = x + (long long)r * 512;
  for (int c = threadIdx.x; c < 512; c += 128) {
    float v;
    if (i < lx) {
      const int lr = pm.lr[b];
      long long id = i < lr ? pm.ref_seq[b][i] : text_seq[pm.txt_in_off[b] + i - lr];
      if (id < 0 || id >= text_vocab) { if (err) *err = 2; id = 0; }
      const float base = proj ? proj[(long long)(pm.txt_off[b] + i) * 512 + c] : proj_bias[c];
      const float e = text_emb[id * 512 + c] + base;            // same association as the graph: (emb + bert) + PE
      v = e + text_alpha[0] * pe_value(i + 1, c, div_term);
    } else {
      const int j = i - lx;
      int tok = pm.prompt_tok[b][j];
      if (tok < 0 || tok > 1024) { if (err) *err = 2; tok = 0; }
      v = audio_emb[(long long)tok * 512 + c] + audio_alpha[0] * pe_value(j + 1, c, div_term);
    }
    out[c] = v;
  }
}

__global__ void __launch_bounds__(128) slot_init_kernel(const SlotInit* __restrict__ init,
                                                        const int* const* __restrict__ prompt_tok, int* hist,
                                                        int hist_ld, int* hist_len, int* kv_len, int* active,
                                                        int* stop_step, SlotParams* params) {
  const SlotInit in = init[blockIdx.x];
  int* row = hist + (long long)in.slot * hist_ld;
  const int* src = prompt_tok[blockIdx.x];
  for (int i = threadIdx.x; i < hist_ld; i += 128) row[i] = i < in.hist_len ? src[i] : 0;
  if (threadIdx.x == 0) {
    hist_len[in.slot] = in.hist_len; kv_len[in.slot] = in.kv_len; active[in.slot] = 1; stop_step[in.slot] = -1;
    params[in.slot] = in.p;
  }
}

// ---- VITS helpers -------------------------------------------------------------
__global__ void gather_rows_kernel(float* out, int ldo, const float* table, int C, const long long* idx,
                                   int rows, int repeat, int table_rows, int* err) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * repeat * C) return;
  int c = (int)(i % C);
  long long ro = i / C;
  int r = (int)(ro / repeat);
  long long id = idx[r];
  if (table_rows > 0 && (id < 0 || id >= table_rows)) {   // never read outside the table: flag (-> error) and clamp
    if (err) *err = 2;
    id = 0;
  }
  out[ro * ldo + c] = table[id * C + c];
}

__global__ void gated_act_kernel(const float* x, int ldx, float* y, int ldy, int H, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * H) return;
  int c = (int)(i % H); long long r = i / H;
  float a = x[r * ldx + c], b = x[r * ldx + H + c];
  y[r * ldy + c] = tanhf(a) * (1.f / (1.f + expf(-b)));
}

__global__ void glu_residual_kernel(const float* y2, int ld2, float* x, int ldx, int H, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * H) return;
  int c = (int)(i % H); long long r = i / H;
  float a = y2[r * ld2 + c], b = y2[r * ld2 + H + c];
  x[r * ldx + c] += a * (1.f / (1.f + expf(-b)));
}

__global__ void flip_channels_kernel(const float* x, float* y, int C, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * C) return;
  int c = (int)(i % C); long long r = i / C;
  y[r * C + c] = x[r * C + (C - 1 - c)];
}

__global__ void sub_cols_kernel(float* z, int ldz, int col0, const float* m, int ldm, int C, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * C) return;
  int c = (int)(i % C); long long r = i / C;
  z[r * ldz + col0 + c] -= m[r * ldm + c];
}

__global__ void zp_kernel(const float* stats, const float* noise, float* zp, float scale, int rows) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * 192) return;
  int c = (int)(i % 192); long long r = i / 192;
  float m = stats[r * 384 + c], logs = stats[r * 384 + 192 + c];
  float nz = noise ? noise[i] : 0.f;
  zp[i] = m + nz * expf(logs) * scale;
}


// conv_post (Cout = 1, k = 7, no bias) fused with the preceding leaky-relu(0.01)
// and the final tanh (vits#[8450-8452]).  One thread per output sample.
// Each input row is read ONCE: the thread that owns row t computes its 7 tap contributions
// d[t][j] = sum_c lrelu(x[t][c]) * w[j][c]; the outputs are then sums over neighbouring rows in shared
// memory.  (The first version re-read every row for each of its 7 taps: 742 us for 737 MB.)
constexpr int CP_ROWS = 256, CP_OUT = CP_ROWS - 6;
__global__ void __launch_bounds__(CP_ROWS) conv_post_tanh_kernel(const float* __restrict__ x, int C,
                                                                 const float* __restrict__ w,
                                                                 float* __restrict__ audio, const int* __restrict__ off) {
  extern __shared__ float ws[];   // [7*C] weights | [CP_ROWS][7] tap contributions
  float* d = ws + 7 * C;
  for (int i = threadIdx.x; i < 7 * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int r0 = off[b], T = off[b + 1] - r0;
  const int t0 = blockIdx.x * CP_OUT;
  if (t0 >= T) return;
  const int ti = t0 - 3 + (int)threadIdx.x;                 // row owned by this thread
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ti >= 0 && ti < T) {
    const float* xr = x + (long long)(r0 + ti) * C;
    for (int c = 0; c < C; c += 4) {
      float4 v = *reinterpret_cast<const float4*>(xr + c);
      v.x = v.x > 0.f ? v.x : v.x * 0.01f; v.y = v.y > 0.f ? v.y : v.y * 0.01f;
      v.z = v.z > 0.f ? v.z : v.z * 0.01f; v.w = v.w > 0.f ? v.w : v.w * 0.01f;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        acc[j] = fmaf(v.x, ws[j * C + c], acc[j]); acc[j] = fmaf(v.y, ws[j * C + c + 1], acc[j]);
        acc[j] = fmaf(v.z, ws[j * C + c + 2], acc[j]); acc[j] = fmaf(v.w, ws[j * C + c + 3], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) d[threadIdx.x * 7 + j] = acc[j];
  __syncthreads();
  const int t = t0 + (int)threadIdx.x;
  if ((int)threadIdx.x < CP_OUT && t < T) {
    // out[t] = sum_j x[t + j - 3] . w[j]: row t + j - 3 sits at slot threadIdx.x + j.  Same order of the
    // j-sum as before; the c-sum is now per tap (rounding differs in the last bits only)
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) o += d[(threadIdx.x + j) * 7 + j];
    audio[r0 + t] = tanhf(o);
  }
}

// spectrogram framing (vits#[3-36]): reflect pad 704 each side, frames of 2048 hop 640,
// periodic Hann window 0.5 - 0.5 cos(2 pi n / 2048)
__global__ void stft_frames_kernel(const float* audio, int n, float* frames, int F) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)F * 2048) return;
  int f = (int)(i >> 11), k = (int)(i & 2047);
  int t = f * 640 + k - 704;
  if (t < 0) t = -t;
  if (t >= n) t = 2 * (n - 1) - t;
  float win = 0.5f - 0.5f * cospif((float)k / 1024.f);
  frames[i] = audio[t] * win;
}

__global__ void dft_matrix_kernel(float* w) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1408LL * 2048) return;
  int row = (int)(i >> 11), n = (int)(i & 2047);
  int bin = row >> 1;
  int ph = (bin * n) & 2047;                 // exact phase reduction
  float a = (float)ph / 1024.f;              // angle / pi
  w[i] = (row & 1) ? -sinpif(a) : cospif(a);
}

__global__ void magnitude_kernel(const float* reim, float* mag, int F) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)F * 704) return;
  long long f = i / 704; int b = (int)(i % 704);
  float re = reim[f * 1408 + 2 * b], im = reim[f * 1408 + 2 * b + 1];
  mag[i] = sqrtf(re * re + im * im + 1e-6f);
}

__global__ void mean_rows_kernel(const float* x, int ld, int C, int rows, float* out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(long long)r * ld + c];
  out[c] = s / (float)rows;
}

__global__ void prelu_add_kernel(float* ge, const float* add, const float* slope, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float v = ge[c] + add[c];
  ge[c] = v < 0.f ? v * slope[c] : v;
}

__global__ void row_sqnorm_kernel(const float* x, int ld, int C, int rows, float* out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { float v = x[(long long)row * ld + c]; s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

__global__ void vq_argmax_kernel(const float* x2, const float* xe, const float* e2, int rows, long long* codes) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float best = -CUDART_INF_F; int bi = 0;
  for (int c = lane; c < 1024; c += 32) {
    float d = -((x2[row] - 2.f * xe[(long long)row * 1024 + c]) + e2[c]);
    if (d > best) { best = d; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) codes[row] = bi;
}


__global__ void transpose_kernel(const float* src, int rows, int cols, float* dst) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
  for (int i = 0; i < 32; i += 8)
    if (r + i < rows && c < cols) tile[threadIdx.y + i][threadIdx.x] = src[(long long)(r + i) * cols + c];
  __syncthreads();
  int oc = blockIdx.y * 32 + threadIdx.x, orow = blockIdx.x * 32 + threadIdx.y;
  for (int i = 0; i < 32; i += 8)
    if (orow + i < cols && oc < rows) dst[(long long)(orow + i) * rows + oc] = tile[threadIdx.x][threadIdx.y + i];
}

}  // namespace

void launch_layernorm(const float* x, const float* res, const float* g, const float* b, float* y, int rows, int C,
                      cudaStream_t s, int nsplit, long long split_stride, const float* lin_bias) {
  if (rows <= 0) return;
  if (C == 512 && rows <= 512 && nsplit <= 8 && (split_stride & 3) == 0) {
    launch_pdl(layernorm_row512_kernel, dim3(rows), dim3(128), 0, s, x, res, g, b, y, nsplit, split_stride, lin_bias);
    GENIE_LAUNCHED("layernorm");
    return;
  }
  switch (C) {
    case 512: layernorm_kernel<16><<<nblk(rows, 8), 256, 0, s>>>(x, res, g, b, y, rows, nsplit, split_stride, lin_bias); break;
    case 192: layernorm_kernel<6><<<nblk(rows, 8), 256, 0, s>>>(x, res, g, b, y, rows, nsplit, split_stride, lin_bias); break;
    case 128: layernorm_kernel<4><<<nblk(rows, 8), 256, 0, s>>>(x, res, g, b, y, rows, nsplit, split_stride, lin_bias); break;
    default: GENIE_CHECK(false, "layernorm: unsupported width");
  }
  GENIE_LAUNCHED("layernorm");
}
void launch_text_embed_pe(float* x, const long long* seq, const int* pos, const float* emb, const float* alpha,
                          const float* div_term, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  text_embed_pe_kernel<<<nblk((long long)rows * 512, 256), 256, 0, s>>>(x, seq, pos, emb, alpha, div_term, rows);
  GENIE_LAUNCHED("text_embed_pe");
}
void launch_audio_embed_pe(float* out, const int* tok, const int* pos, const float* emb, const float* alpha,
                           const float* div_term, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  audio_embed_pe_kernel<<<nblk((long long)rows * 512, 256), 256, 0, s>>>(out, tok, pos, emb, alpha, div_term, rows);
  GENIE_LAUNCHED("audio_embed_pe");
}
void launch_decode_embed(float* out, const int* hist, int hist_ld, const int* hist_len, const int* active,
                         const float* emb, const float* alpha, const float* div_term, int B, cudaStream_t s) {
  if (B <= 0) return;
  decode_embed_kernel<<<B, 512, 0, s>>>(out, hist, hist_ld, hist_len, active, emb, alpha, div_term, 0);
  GENIE_LAUNCHED("decode_embed");
}
void launch_kv_scatter(const float* qkv, int ld, void* kv_base, int kv_f16, long long utt_stride, long long layer_off,
                       long long v_off, int cap, const int* row_off, const int* dst_pos0, const int* row2utt,
                       int rows, const int* active, cudaStream_t s, const int* slot_of) {
  if (rows <= 0) return;
  if (kv_f16)
    kv_scatter_kernel<__half><<<nblk((long long)rows * 256, 256), 256, 0, s>>>(
        qkv, ld, reinterpret_cast<__half*>(kv_base), utt_stride, layer_off, v_off, cap, row_off, dst_pos0, row2utt, rows,
        active, slot_of);
  else
    kv_scatter_kernel<float><<<nblk((long long)rows * 256, 256), 256, 0, s>>>(
        qkv, ld, reinterpret_cast<float*>(kv_base), utt_stride, layer_off, v_off, cap, row_off, dst_pos0, row2utt, rows,
        active, slot_of);
  GENIE_LAUNCHED("kv_scatter");
}
void launch_prefill_bert_gather(float* bert, const PrefillMeta& pm, const float* text_bert, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  prefill_bert_gather_kernel<<<rows, 256, 0, s>>>(bert, pm, text_bert);
  GENIE_LAUNCHED("prefill_bert_gather");
}
void launch_prefill_embed(float* x, const PrefillMeta& pm, const long long* text_seq, const float* proj,
                          const float* proj_bias, const float* text_emb, const float* text_alpha, int text_vocab,
                          const float* audio_emb, const float* audio_alpha, const float* div_term, int rows,
                          int* err, cudaStream_t s) {
  if (rows <= 0) return;
  prefill_embed_kernel<<<rows, 128, 0, s>>>(x, pm, text_seq, proj, proj_bias, text_emb, text_alpha, text_vocab,
                                            audio_emb, audio_alpha, div_term, err);
  GENIE_LAUNCHED("prefill_embed");
}
void launch_slot_init(const SlotInit* init, int n, const int* const* prompt_tok, int* hist, int hist_ld, int* hist_len,
                      int* kv_len, int* active, int* stop_step, SlotParams* params, cudaStream_t s) {
  if (n <= 0) return;
  slot_init_kernel<<<n, 128, 0, s>>>(init, prompt_tok, hist, hist_ld, hist_len, kv_len, active, stop_step, params);
  GENIE_LAUNCHED("slot_init");
}
void launch_gather_rows(float* out, int ldo, const float* table, int C, const long long* idx, int rows, int repeat,
                        cudaStream_t s, int table_rows, int* err) {
  if (rows <= 0) return;
  gather_rows_kernel<<<nblk((long long)rows * repeat * C, 256), 256, 0, s>>>(out, ldo, table, C, idx, rows, repeat,
                                                                            table_rows, err);
  GENIE_LAUNCHED("gather_rows");
}
void launch_gated_act(const float* x, int ldx, float* y, int ldy, int H, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  gated_act_kernel<<<nblk((long long)rows * H, 256), 256, 0, s>>>(x, ldx, y, ldy, H, rows);
  GENIE_LAUNCHED("gated_act");
}
void launch_glu_residual(const float* y2, int ld2, float* x, int ldx, int H, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  glu_residual_kernel<<<nblk((long long)rows * H, 256), 256, 0, s>>>(y2, ld2, x, ldx, H, rows);
  GENIE_LAUNCHED("glu_residual");
}
void launch_flip_channels(const float* x, float* y, int C, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  flip_channels_kernel<<<nblk((long long)rows * C, 256), 256, 0, s>>>(x, y, C, rows);
  GENIE_LAUNCHED("flip_channels");
}
void launch_sub_cols(float* z, int ldz, int col0, const float* m, int ldm, int C, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  sub_cols_kernel<<<nblk((long long)rows * C, 256), 256, 0, s>>>(z, ldz, col0, m, ldm, C, rows);
  GENIE_LAUNCHED("sub_cols");
}
void launch_zp(const float* stats, const float* noise, float* zp, float scale, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  zp_kernel<<<nblk((long long)rows * 192, 256), 256, 0, s>>>(stats, noise, zp, scale, rows);
  GENIE_LAUNCHED("zp");
}
void launch_conv_post_tanh(const float* x, int C, const float* w, float* audio, const int* off, int B, int maxT,
                           cudaStream_t s) {
  if (B <= 0 || maxT <= 0) return;
  conv_post_tanh_kernel<<<dim3(nblk(maxT, CP_OUT), B), CP_ROWS, (7 * C + CP_ROWS * 7) * sizeof(float), s>>>(x, C, w, audio, off);
  GENIE_LAUNCHED("conv_post_tanh");
}
void launch_stft_frames(const float* audio, int n, float* frames, int F, cudaStream_t s) {
  stft_frames_kernel<<<nblk((long long)F * 2048, 256), 256, 0, s>>>(audio, n, frames, F);
  GENIE_LAUNCHED("stft_frames");
}
void launch_dft_matrix(float* w, cudaStream_t s) {
  dft_matrix_kernel<<<nblk(1408LL * 2048, 256), 256, 0, s>>>(w);
  GENIE_LAUNCHED("dft_matrix");
}
void launch_magnitude(const float* reim, float* mag, int F, cudaStream_t s) {
  magnitude_kernel<<<nblk((long long)F * 704, 256), 256, 0, s>>>(reim, mag, F);
  GENIE_LAUNCHED("magnitude");
}
void launch_mean_rows(const float* x, int ld, int C, int rows, float* out, cudaStream_t s) {
  mean_rows_kernel<<<nblk(C, 128), 128, 0, s>>>(x, ld, C, rows, out);
  GENIE_LAUNCHED("mean_rows");
}
void launch_prelu_add(float* ge, const float* add, const float* slope, int C, cudaStream_t s) {
  prelu_add_kernel<<<nblk(C, 128), 128, 0, s>>>(ge, add, slope, C);
  GENIE_LAUNCHED("prelu_add");
}
void launch_vq_argmax(const float* x2, int, const float* xe, const float* e2, int rows, long long* codes,
                      cudaStream_t s) {
  if (rows <= 0) return;
  vq_argmax_kernel<<<nblk(rows, 8), 256, 0, s>>>(x2, xe, e2, rows, codes);
  GENIE_LAUNCHED("vq_argmax");
}
void launch_row_sqnorm(const float* x, int ld, int C, int rows, float* out, cudaStream_t s) {
  if (rows <= 0) return;
  row_sqnorm_kernel<<<nblk(rows, 8), 256, 0, s>>>(x, ld, C, rows, out);
  GENIE_LAUNCHED("row_sqnorm");
}
void launch_transpose(const float* src, int rows, int cols, float* dst, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return;
  transpose_kernel<<<dim3(nblk(cols, 32), nblk(rows, 32)), dim3(32, 8), 0, s>>>(src, rows, cols, dst);
  GENIE_LAUNCHED("transpose");
}

}  // namespace genie
