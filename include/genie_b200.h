/* genie_b200 — C-ABI of the B200-native GPT-SoVITS synthesis hot path.
 *
 * Drop-in boundary: the reference (Genie-TTS) has no FFI of its own; its operator
 * boundary is the onnxruntime.InferenceSession duck type held in GSVModel
 * (/root/reference/src/genie_tts/ModelManager.py:48-56) and called at
 * src/genie_tts/Core/Inference.py:47,55,76,88,102 and
 * src/genie_tts/Audio/ReferenceAudio.py:72-73.  This library replaces those
 * sessions one level up: one entry point per *stage group*
 * instead of one run() per decode step with 48 KV tensors through numpy.
 *
 * Conventions: every function returns 0 on success, non-zero on failure and
 * never aborts; genie_last_error() returns the message of the last failure on
 * the calling thread.  Plain pointers and sizes only.  Pointers are HOST
 * pointers unless the `io_on_device` argument of the call is 1, in which case
 * the in/out payload pointers are device pointers on the model's GPU (used by
 * bench.py's HBM-resident `value` leg).  All integer token ids are int64 to
 * match the reference's numpy feeds.  One handle belongs to one GPU; calls on one
 * handle are serialised by the library (a second thread waits), calls on different
 * handles — including contexts cloned from one model — run concurrently.
 *
 * Streams: every handle issues its work on ONE main CUDA stream (cuBLAS-style
 * binding rather than a stream argument on every call): the library's own by
 * default, or a caller-owned cudaStream_t given to genie_context_create /
 * genie_set_stream.  Calls return after their results are complete on that stream.
 */
#ifndef GENIE_B200_H
#define GENIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct genie_model genie_model;
typedef struct genie_prompt genie_prompt;

/* which graph file of the model directory a tensor belongs to
 * (src/genie_tts/ModelManager.py:26-45: GSVModelFile) */
enum { GENIE_GRAPH_T2S_ENCODER = 0, GENIE_GRAPH_T2S = 1, GENIE_GRAPH_VITS = 2, GENIE_GRAPH_PROMPT_ENCODER = 3 };
enum { GENIE_F32 = 0, GENIE_F16 = 1 };

const char* genie_last_error(void);
int genie_version(void);
/* kernels launched by this library since load (bench.py: gpu_launches) */
unsigned long long genie_launch_count(void);
int genie_device_count(void);
/* page-locked host buffers for result payloads (the 1280 samples per token of genie_vits_decode cross PCIe at
 * DMA speed instead of through the driver's bounce buffers); the host side pools and reuses them */
int genie_host_alloc(size_t bytes, void** out);
int genie_host_free(void* p);

/* ---- model store: replaces ModelManager.load_character / load_session_with_fp16_conversion
 * (src/genie_tts/ModelManager.py:59-114, 231-310).  The host side parses the .onnx
 * initialiser tables and hands each tensor's bytes AS STORED (fp16 for the fp16 bins,
 * fp32 for t2s_encoder_fp32.bin); nothing is up-cast on the host. */
int genie_model_create(int device, genie_model** out);
int genie_model_add_tensor(genie_model* m, int graph, const char* name, const void* host_data, int dtype,
                           const int64_t* dims, int ndim);
/* architecture constants that live in Constant nodes of the graphs:
 * sinusoidal div_term (t2s_stage_decoder#[21]), top_k 15 (#[1788]), repetition penalty 1.35
 * (#[1780]), temperature 1.0 (#[1786]), vocoder noise scale 0.5 (vits#[6494]) */
int genie_model_set_constants(genie_model* m, const float* pe_div_term_256, int top_k, float repetition_penalty,
                              float temperature, float noise_scale);
/* folds weight-norm (vits#[6508-6510] x131, once instead of per call), repacks conv weights
 * channels-last, builds per-layer tables.  V2 vs V2ProPlus is decided by the presence of
 * prompt-encoder tensors (ModelManager.py:287-293). */
int genie_model_finalize(genie_model* m);
int genie_model_info(const genie_model* m, int* is_v2pp, long long* weight_bytes, long long* workspace_bytes);
/* Destroys the handle.  Weights are reference-counted: they are freed when the model AND every context cloned
 * from it are gone.  Prompts never dereference the model, so they may be destroyed before or after it. */
void genie_model_destroy(genie_model* m);

/* ---- execution contexts: a second (third, ...) independent handle onto the SAME weights, with its own stream,
 * workspace, T2S slot pool and captured graphs — ">1 session per model" (one scheduler thread per GPU
 * plus e.g. a latency-critical batch-1 caller beside it).  cuda_stream: a caller-owned cudaStream_t the context
 * issues its work on, or NULL for a library-owned stream.  Every entry point taking a genie_model* accepts a
 * context; prompts built through any handle of a model are valid for all of them.  Free with genie_model_destroy. */
int genie_context_create(genie_model* model, void* cuda_stream, genie_model** out);
/* re-bind a handle's main stream (NULL: back to a library-owned stream); drops captured graphs */
int genie_set_stream(genie_model* m, void* cuda_stream);

/* ---- prompt: replaces the per-reference-audio work the reference redoes in every call:
 * the VQ of ssl_content (t2s_encoder#[2-48]), the V2 ref_enc spectrogram path (vits#[3-271])
 * or the V2ProPlus prompt_encoder graph (ReferenceAudio.py:68-76), plus ge-only conditioning.
 *   ref_seq   int64[Lr]          ReferenceAudio.phonemes_seq
 *   ref_bert  f32[Lr,1024]|NULL  ReferenceAudio.text_bert (NULL == zeros, ja/en)
 *   ssl       f32[768,Ts]        ReferenceAudio.ssl_content[0]
 *   ref_audio f32[n_audio]       ReferenceAudio.audio_32k[0]
 *   sv_emb    f32[20480]|NULL    speaker-verification embedding (V2ProPlus only) */
int genie_prompt_create(genie_model* m, const int64_t* ref_seq, int Lr, const float* ref_bert, const float* ssl,
                        int Ts, const float* ref_audio, int n_audio, const float* sv_emb, genie_prompt** out);
/* V2ProPlus alternative: global embeddings already computed (the reference's vits graph inputs
 * `ge` f32[1024], `ge_advanced` f32[512]) */
int genie_prompt_create_with_ge(genie_model* m, const int64_t* ref_seq, int Lr, const float* ref_bert,
                                const float* ssl, int Ts, const float* ge, int ge_dim, const float* ge_advanced,
                                genie_prompt** out);
int genie_prompt_info(const genie_prompt* p, int* n_prompt_tokens, int* ge_dim, int* ref_len);
/* prompts int64[n_prompt_tokens], ge f32[ge_dim], ge_advanced f32[512] (V2ProPlus) — any may be NULL */
int genie_prompt_read(const genie_prompt* p, int64_t* prompts, float* ge, float* ge_advanced);
void genie_prompt_destroy(genie_prompt* p);

/* ---- T2S: replaces encoder.run + first_stage_decoder.run + <=500 x stage_decoder.run
 * (src/genie_tts/Core/Inference.py:76-106) for a batch of independent utterances. */
typedef struct genie_sampling {
  int top_k;                  /* <=0: model constant (15) */
  float temperature;          /* <=0: model constant (1.0) */
  float repetition_penalty;   /* <=0: model constant (1.35) */
  int greedy;                 /* 1: sampler noise == 1 (bit-comparable mode); 0: Philox N(0,1) */
  unsigned long long seed;
  int max_steps;              /* decode-loop bound, reference 500 (Inference.py:95); <=0: 500 */
  int fixed_steps;            /* >0: ignore stop flags and run exactly this many loop iterations */
  float top_p;                /* nucleus sampling, an EXTENSION: the reference graphs have no top-p node.
                                 <=0 or >=1: off (reference behaviour).  Applied to the penalised logits before
                                 temperature and top-k, upstream GPT-SoVITS order: in descending order drop every
                                 token whose inclusive cumulative probability exceeds top_p, except the first */
} genie_sampling;
/* All sampling parameters live in device memory per decode slot, so the captured decode-step graph is independent
 * of them: changing seed / temperature / top_k / top_p between calls costs nothing. */

/* text_seq: int64 concat over utterances (sum text_len); text_bert: f32 [sum text_len,1024] or NULL.
 * Outputs (host unless io_on_device): y int64[B, y_ld] = prompt tokens followed by every generated
 * token (the reference's `y` before `y[0,-1]=0; y[:, -idx:]`, Inference.py:108-109, which the host
 * mirror applies), y_len[B] valid entries, idx[B] = value of the reference's loop variable at exit.
 * cancel: optional host flag polled between decode steps (Inference.py:96-97); returns
 * GENIE_CANCELLED. */
#define GENIE_CANCELLED 2
int genie_t2s_generate(genie_model* m, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                       const int* text_len, const float* text_bert, const genie_sampling* sampling,
                       const volatile int* cancel, int io_on_device, int64_t* y, int y_ld, int* y_len, int* idx);

/* The same work in three calls (t2s_prefill / t2s_decode_steps), for callers that stream tokens, poll
 * between chunks of steps or interleave other work: prefill = encoder.run + first_stage_decoder.run
 * (Inference.py:76-93) for the batch, leaving it "in flight" inside the model (one batch in flight per model: the
 * KV cache and step buffers are the model's workspace); decode_steps = up to n_steps iterations of the
 * stage-decoder loop (Inference.py:95-106), stopping early when every utterance has stopped or the sampling's
 * max_steps / fixed_steps budget is used (n_active = utterances still decoding, steps_done = loop iterations so
 * far; GENIE_CANCELLED when the host flag fired); read = y / y_len / idx as documented above for what has been
 * generated so far.  genie_t2s_generate == prefill + decode_steps(budget) + read. */
int genie_t2s_prefill(genie_model* m, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                      const int* text_len, const float* text_bert, const genie_sampling* sampling, int io_on_device);
int genie_t2s_decode_steps(genie_model* m, int n_steps, const volatile int* cancel, int* n_active, int* steps_done);
int genie_t2s_read(genie_model* m, int io_on_device, int64_t* y, int y_ld, int* y_len, int* idx);

/* ---- continuous batching (BASELINE config 5): the T2S stage as a pool of decode SLOTS.  A slot owns a
 * KV slab of kv_capacity tokens and a token-history row; requests are admitted into free slots (the call runs their
 * prefill as one ragged batch, Inference.py:76-93) while other slots are mid-decode, every pool_step advances all
 * active slots together (Inference.py:95-106 per slot, each with its own sampling parameters, stop flag and step
 * budget), finished slots are read and released.  Per-slot results are independent of what shares the pool.
 * The batch calls above are the special case "pool sized for the batch, everything admitted at once"; on one
 * handle use either the batch calls or a pool (a prefill call replaces the pool).
 *   kv_capacity        tokens per slot: >= Lr + Lt + prompt tokens + max_steps + 1 of the longest request
 *   max_prompt_tokens  longest prompt-token sequence (Ts/2) that will be admitted
 *   max_steps          largest decode-loop bound that will be admitted (reference: 500) */
int genie_t2s_pool_create(genie_model* m, int n_slots, int kv_capacity, int max_prompt_tokens, int max_steps);
int genie_t2s_pool_info(genie_model* m, int* n_slots, int* kv_capacity, int* hist_ld);
/* admit n requests into the free slots slots[0..n); text_seq / text_len / text_bert as in genie_t2s_generate (host
 * pointers); sampling: array of n genie_sampling (NULL: model constants for all) */
int genie_t2s_admit(genie_model* m, int n, const int* slots, genie_prompt* const* prompts, const int64_t* text_seq,
                    const int* text_len, const float* text_bert, const genie_sampling* sampling);
/* up to n_steps decode steps over every active slot (returns early when none is left); n_active: slots still decoding */
int genie_t2s_pool_step(genie_model* m, int n_steps, int* n_active);
/* state[i]: 0 free, 1 decoding, 2 finished (stop flag or step budget); n_generated[i]: tokens generated so far
 * (the first-stage token included); arrays of n >= n_slots entries, either may be NULL */
int genie_t2s_pool_poll(genie_model* m, int* state, int* n_generated, int n);
/* y / y_len / idx of one slot, as documented for genie_t2s_generate (valid at any time while the slot is in use) */
int genie_t2s_pool_read(genie_model* m, int slot, int64_t* y, int y_capacity, int* y_len, int* idx);
/* free the slot (a slot released while still decoding — a cancelled request — stops at once) */
int genie_t2s_release(genie_model* m, int slot);

/* ---- SoVITS: replaces vocoder.run (Inference.py:46-61) for a batch.
 * sem: int64 concat of semantic tokens (sum sem_len, every id < 1024);
 * zp_noise: f32 concat per utterance of [192, 2*sem_len[b]] (the graph's RandomNormalLike
 * layout, vits#[6490]) or NULL -> Philox N(0,1) from `seed`; noise_scale < 0 -> model constant.
 * audio: f32 concat, 1280 samples per semantic token; audio_len[B].
 * noise_ids: int[B] or NULL — the Philox stream of utterance b is keyed by (seed, noise_ids[b]) instead of (seed, b),
 * so that several batches vocoded in ONE call draw the noise they would have drawn in separate calls (host int[B]). */
int genie_vits_decode(genie_model* m, genie_prompt* const* prompts, int B, const int64_t* text_seq,
                      const int* text_len, const int64_t* sem, const int* sem_len, const float* zp_noise,
                      unsigned long long seed, float noise_scale, int io_on_device, float* audio, int* audio_len,
                      const int* noise_ids);

/* ---- introspection for parity tests (host pointers only) */
/* logits of the last T2S call: f32[B, n_steps+1, 1025] when recording was enabled */
int genie_debug_record_logits(genie_model* m, int enable);
int genie_debug_read_logits(genie_model* m, float* out, int max_floats, int* n_floats);
/* intermediate tensors of the last T2S / VITS call by name ("x", "k0", "m_p", "z", ...) */
int genie_debug_read(genie_model* m, const char* what, float* out, long long max_floats, long long* n_floats);
int genie_debug_keep(genie_model* m, int enable);
/* the sampler kernel in isolation (stage#[1775-1821]): rows of host logits f32[rows,1025] with token histories
 * int64[rows,hist_ld] / hist_len[rows]; n_draws independent draws per row (Philox key (seed + draw, row, hist_len),
 * or externally supplied noise f32[n_draws, rows, 1025] replacing RandomNormalLike); tokens int64[rows, n_draws],
 * stop int[rows, n_draws] (may be NULL).  Histories are not modified. */
int genie_debug_sample(genie_model* m, const float* logits, int rows, const int64_t* hist, int hist_ld,
                       const int* hist_len, const genie_sampling* sampling, const float* noise, int n_draws,
                       int64_t* tokens, int* stop);
/* unit self-test of the tcgen05 implicit-GEMM conv kernel against the exact SIMT kernel on random data:
 * M rows in two ragged segments, ntaps taps with dilation dil; mode 1 = x_hi.w_hi, 2 = (x_hi+x_lo).w_hi,
 * 3 = 2 + x_hi.w_lo; exact_w = weights rounded to fp16 (the T2S case) */
int genie_debug_tc_selftest(int M, int Cin, int Cout, int ntaps, int dil, int mode, int exact_w, float* max_err,
                            float* ref_max);
/* cudaProfilerStart (1) / cudaProfilerStop (0): lets `ncu --profile-from-start off` capture one bench step */
int genie_profiler_range(int on);
/* stage timing of the last calls (CUDA events on the library's stream), up to 12 floats:
 * [0] prefill ms, [1] decode ms, [2] t2s total ms, [3] decode steps, [4] vits ms, [5] generator ms,
 * [6] generator launches, [7] latent rows, [8] decode-attention us per launch and [9] its KV MB per launch
 * (only after a genie_t2s_generate with option time_attention > 0), [10] ms of the generator stages with <= 32
 * channels (HBM-bound) and [11] their algorithmic MB */
int genie_last_timing(genie_model* m, float* ms, int n);
/* options: kv_fp16 (1, default: KV cache rows stored as fp16 — q, scores and accumulators stay fp32; 0: fp32 rows as
 * in the reference graphs; takes effect at the next prefill / pool_create),
 * use_graph (CUDA-graph replay of the decode step, default 1), use_tc / tc_vits (tensor-core paths),
 * skinny_max_rows, tc_min_rows, decode_split_min (path selection, debugging),
 * time_attention (n > 0: after the next t2s_generate replay the decode attention n x 24 times between events),
 * prefill_single (experiment, default 0: 1 = prefill linears as ONE fp16 product; measured 94 / 100 on the
 * acceptance set, so it is off), persistent_step, decode_branches, fuse_pairs, sm_partition (debugging / studies).
 * Environment: GENIE_TC_HALO=-1 disables the halo conv kernel, GENIE_HALO_PIPE=0 the persistent role-split form of
 * its Cin = 128 variant (4 = one CTA per SM with staged residual), GENIE_PDL=0 programmatic dependent launch,
 * GENIE_SYNC_DEBUG=1 synchronises after every launch and names the failing kernel. */
int genie_set_option(genie_model* m, const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* GENIE_B200_H */
