// Launchers for the small fused / elementwise kernels (elementwise.cu, sampler.cu).
#pragma once
#include "common.cuh"

namespace genie {

// Per-slot decode parameters, resident in device memory: the captured decode-step graph reads them from here, so
// it depends on neither a request's sampling configuration nor its seed (defaults = the constants baked into the
// reference graphs, stage#[1780-1790]; top_p is an extension the graphs do not have, 1.0 = off).
struct SlotParams {
  int top_k; int greedy; int honour_stop;
  int hist_max;                 // decoding stops once the history (prompt + generated) holds this many tokens
  float temperature, penalty, top_p;
  int utt;                      // Philox key part: index of the utterance within its admission call
  unsigned long long seed;
};


// QKV split-K partials [nsplit][B][1536] (+ bias) -> q, append k/v at kv_len[b], attend over kv_len[b] + 1 tokens
// kv_f16: the cache holds halves (strides stay in elements)
void launch_decode_attention_fused(const float* part, int nsplit, long long split_stride, const float* bias, float* o,
                                   void* kv_base, int kv_f16, long long utt_stride, long long layer_off,
                                   long long v_off, const int* kv_len, const int* active, int B, int cap, float scale,
                                   cudaStream_t s);

// y[r,:] = LN(x[r,:] (+ res[r,:])) * g + b   (eps 1e-5), C multiple of 32, C <= 1024
// x may be nsplit split-K partials (split_stride apart) with the producer's bias deferred to here
void launch_layernorm(const float* x, const float* res, const float* g, const float* b, float* y,
                      int rows, int C, cudaStream_t s, int nsplit = 1, long long split_stride = 0,
                      const float* lin_bias = nullptr);

// x[r,:] += emb[seq[r],:] + alpha * PE(pos[r])     (text rows; pos is 1-based per utterance)
void launch_text_embed_pe(float* x, const long long* seq, const int* pos, const float* emb, const float* alpha,
                          const float* div_term, int rows, cudaStream_t s);
// out[r,:] = emb[tok[r],:] + alpha * PE(pos[r]);  raw (no PE) copy optional
void launch_audio_embed_pe(float* out, const int* tok, const int* pos, const float* emb, const float* alpha,
                           const float* div_term, int rows, cudaStream_t s);
// decode-step variant: token = hist[b, hist_len[b]-1], position = hist_len[b] - 0 (1-based audio position)
void launch_decode_embed(float* out, const int* hist, int hist_ld, const int* hist_len, const int* active,
                         const float* emb, const float* alpha, const float* div_term, int B, cudaStream_t s);

// scatter K,V columns of qkv rows into the head-major cache:
//   cache[b][layer][kv][h][pos][32], pos = dst_pos0[b] + (row - row_off[b])
//   slot_of (optional): cache slab of utterance b is slot_of[b] instead of b
void launch_kv_scatter(const float* qkv, int ld, void* kv_base, int kv_f16, long long utt_stride, long long layer_off,
                       long long v_off, int cap, const int* row_off, const int* dst_pos0, const int* row2utt,
                       int rows, const int* active, cudaStream_t s, const int* slot_of = nullptr);

// ---- prefill input assembly for a ragged batch of newly admitted utterances (one launch each instead of
// per-utterance copies): rows of utterance b are [row_off[b], row_off[b+1]) = Lr ref phones, Lt text phones
// (together lx[b] text rows), Ly prompt tokens
struct PrefillMeta {
  const int* row2utt;            // [R]
  const int* row_off;            // [n+1]
  const int* txt_off;            // [n+1] offsets in text-row space (sum of lx)
  const int* txt_in_off;         // [n]   offset of the utterance's text_seq / text_bert rows in the call's input
  const int* lr; const int* lx;  // [n]
  const long long* const* ref_seq;   // [n] device pointers (prompt)
  const float* const* ref_bert;      // [n] device pointers or null
  const int* const* prompt_tok;      // [n] device pointers
};
// bert[txt_off[b] + i, :] = ref_bert / text_bert row or zeros   (t2s_encoder: Concat(ref_bert, text_bert))
void launch_prefill_bert_gather(float* bert, const PrefillMeta& pm, const float* text_bert, int rows, cudaStream_t s);
// x[r, :] = Emb_text[id] + proj[txt row] (or bias) + alpha_t PE(i+1)  |  Emb_audio[tok] + alpha_a PE(j+1)
void launch_prefill_embed(float* x, const PrefillMeta& pm, const long long* text_seq, const float* proj,
                          const float* proj_bias, const float* text_emb, const float* text_alpha, int text_vocab,
                          const float* audio_emb, const float* audio_alpha, const float* div_term, int rows,
                          int* err, cudaStream_t s);
// per-slot state of newly admitted utterances: history row (prompt tokens, rest zero), lengths, flags, parameters
struct SlotInit { int slot, kv_len, hist_len, pad; SlotParams p; };
void launch_slot_init(const SlotInit* init, int n, const int* const* prompt_tok, int* hist, int hist_ld, int* hist_len,
                      int* kv_len, int* active, int* stop_step, SlotParams* params, cudaStream_t s);

// decode-step single-shot tcgen05 GEMM (tc_small_gemm.cu): raw split-K partials of
// act_in(sum_s x_s + a_bias) . W^T for rows <= 128, K multiple of 256
struct SmallGemm {
  const float* x = nullptr; int ldx = 0; int a_nsplit = 1; long long a_stride = 0;
  const float* a_bias = nullptr; int a_relu = 0;
  const __half* w = nullptr; int ldw = 0;
  int N = 0, K = 0, M = 0;
  float* y = nullptr; int ldy = 0; long long split_stride = 0;
};
void launch_tc_small_gemm(const SmallGemm& p, int nt, int* err_flag, cudaStream_t s);

// ---- persistent decode step for batch <= 8 (t2s_persistent.cu)
struct StepLayerPtrs {
  const __half *wqkv, *wout, *wff1, *wff2;
  const float *bqkv, *bout, *bff1, *bff2, *ln1g, *ln1b, *ln2g, *ln2b;
};
struct PersistentStep {
  const StepLayerPtrs* layers = nullptr; int n_layers = 24;      // device array
  const __half* wpredict = nullptr; const float* bpredict = nullptr; int vocab = 1025;
  float* h = nullptr;          // [B,512] in: embedded rows of this step; then the running layer input (residual)
  float* qkv = nullptr;        // [B,1536]
  float* part = nullptr;       // [B,16,nch,36] attention partials
  float* lnin = nullptr;       // [B,512] out-proj + residual (input of LN1)
  float* lnin2 = nullptr;      // [B,512] FFN2 + residual (input of LN2)
  float* h1 = nullptr;         // [B,512] LN1 output
  float* ff = nullptr;         // [B,2048]
  float* logits = nullptr; int ld_logits = 1025;
  void* kv = nullptr; int kv_f16 = 0;   // cache of floats or halves; strides in elements
  long long utt_stride = 0, layer_stride = 0, v_off = 0; int cap = 0;
  const int* kv_len = nullptr; const int* active = nullptr;
  unsigned* sync = nullptr;    // [0] barrier counter (zeroed per launch), [1] barrier-timeout flag
  int B = 1, nch = 1; float scale = 1.f;
};
int persistent_step_chunks(int B, int grid);
bool persistent_step_fits(int B, int grid);
void launch_t2s_step_persistent(const PersistentStep& a, int grid, cudaStream_t s);

struct SamplerArgs {
  const float* logits;     // [B, ld], row b
  int ld;
  int* hist;               // [slots, hist_ld] token history (prompt + generated)
  int hist_ld;
  int* hist_len;           // [slots]
  int* kv_len;             // [slots] incremented when advance_kv
  int* active;             // [slots] cleared on stop (when honour_stop) or when the budget hist_max is reached
  int* stop_step;          // [slots] history length at which the stop flag fired (or -1)
  const SlotParams* params;   // [slots]
  const int* slot_map;     // row b -> slot (prefill of newly admitted utterances); null = identity (decode)
  int B;
  int advance_kv; int check_stop;
  const float* dbg_noise;  // optional [B,1025] externally supplied noise (tests)
  int dbg_no_append;       // tests: write the token to dbg_tokens[b] instead of appending to the history
  int* dbg_tokens; int* dbg_stop;
};
void launch_sampler(const SamplerArgs& a, cudaStream_t s);

// ---- VITS helpers (channels-last fp32)
// out[r*repeat + j, :] = table[idx[r], :]; table_rows > 0: ids outside [0, table_rows) set *err = 2 and read row 0
void launch_gather_rows(float* out, int ldo, const float* table, int C, const long long* idx, int rows, int repeat,
                        cudaStream_t s, int table_rows = 0, int* err = nullptr);
void launch_gated_act(const float* x, int ldx, float* y, int ldy, int H, int rows, cudaStream_t s);  // tanh(a)*sigmoid(b)
void launch_glu_residual(const float* y2, int ld2, float* x, int ldx, int H, int rows, cudaStream_t s); // x += a*sigmoid(b)
void launch_flip_channels(const float* x, float* y, int C, int rows, cudaStream_t s);
// z[:, 96:] -= mean  (flow coupling reverse, logs == 0)
void launch_sub_cols(float* z, int ldz, int col0, const float* m, int ldm, int C, int rows, cudaStream_t s);
// zp = m + noise * exp(logs) * scale ; stats [rows, 384] = (m | logs); noise may be null (zeros)
void launch_zp(const float* stats, const float* noise, float* zp, float scale, int rows, cudaStream_t s);
// audio[t] = tanh( sum_{j<7,c<C} lrelu(x[t+j-3, c], 0.01) * w[j*C+c] )   per segment
void launch_conv_post_tanh(const float* x, int C, const float* w, float* audio, const int* off, int B, int maxT,
                           cudaStream_t s);
// spectrogram frames: reflect-pad 704, frame 2048, hop 640, periodic Hann -> frames [F, 2048]
void launch_stft_frames(const float* audio, int n, float* frames, int F, cudaStream_t s);
void launch_dft_matrix(float* w, cudaStream_t s);   // [1408, 2048]: rows 2b = cos, 2b+1 = -sin for bin b < 704
void launch_magnitude(const float* reim, float* mag, int F, cudaStream_t s);   // [F,1408] -> [F,704]
void launch_mean_rows(const float* x, int ld, int C, int rows, float* out, cudaStream_t s);
void launch_prelu_add(float* ge, const float* add, const float* slope, int C, cudaStream_t s);
// VQ: codes[r] = argmax_c -(|x_r|^2 - 2 x_r.e_c + |e_c|^2) given dots [rows, 1024] = (2*x).e
void launch_vq_argmax(const float* x, int ldx, const float* xe2, const float* e2, int rows, long long* codes,
                      cudaStream_t s);
void launch_row_sqnorm(const float* x, int ld, int C, int rows, float* out, cudaStream_t s);
void launch_transpose(const float* src, int rows, int cols, float* dst, cudaStream_t s);   // dst[c, r] = src[r, c]

}  // namespace genie
