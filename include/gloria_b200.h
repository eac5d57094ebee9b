/*
 * gloria_b200.h -- C ABI of libgloria_b200.so: GLoRIA local/global contrastive similarity on B200 (sm_100a).
 *
 * The reference (strongbeamsprout/gloria-nlp-project) is pure Python/PyTorch and has NO FFI or plugin interface;
 * its boundary for this path is the set of Python functions in gloria/loss/gloria_loss.py and the loss methods of
 * gloria/models/gloria_model.py.  Each entry point below names the reference code it
 * replaces (file:line relative to the reference root).  INTEGRATION.md shows the ctypes binding and the
 * reference-side monkey patch.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator); the library never
 *     allocates, frees or retains memory, and keeps no mutable global state (re-entrant, device-explicit);
 *   - `stream` is the caller's cudaStream_t passed as void*; all work is enqueued there, nothing synchronises;
 *   - return value 0 = success, otherwise a gloria_status (or 1000 + cudaError_t); gloria_b200_last_error()
 *     gives a thread-local human-readable message.  No exceptions, no exit().
 *   - layouts are the reference's native ones: region features ctx [Bi, D, S] (S = H*W regions contiguous,
 *     gloria_loss.py:30), word features words [Bc, D, Lw] (word axis contiguous, text_model.py:126-131),
 *     cap_lens int32 [Bc].  The caption's words are columns [word_off, word_off + cap_len) of `words`
 *     (word_off 0: local_loss, gloria_loss.py:122; word_off 1: get_local_similarities, gloria_model.py:179).
 *   - sim is [Bi, Bc] row-major: sim[j, i] = pair (image j, caption i), the layout of `similarities`
 *     (gloria_loss.py:160-164) BEFORE the temp3 scale.
 */
#ifndef GLORIA_B200_H_
#define GLORIA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gloria_status {
  GLORIA_OK = 0,
  GLORIA_ERR_BAD_ARG = 1,       /* null pointer, non-positive size, unsupported shape */
  GLORIA_ERR_WORKSPACE = 2,     /* workspace too small */
  GLORIA_ERR_UNSUPPORTED = 3,   /* e.g. backward of agg = max */
  GLORIA_ERR_DRIVER = 4,        /* a driver entry point (tensor-map encode) is unavailable */
  GLORIA_ERR_CUDA_BASE = 1000   /* 1000 + cudaError_t */
} gloria_status;

/* aggregation over the caption's words of exp(temp2 * cos): gloria_loss.py:153-158 / gloria_model.py:198-201 */
#define GLORIA_AGG_SUM 0
#define GLORIA_AGG_MEAN 1
#define GLORIA_AGG_MAX 2

int gloria_b200_version(void);
/* sha256 prefix (24 hex digits) of the sources this binary was built from: csrc/* and this header, as hashed by
 * gloria_nlp_project_b200/build.py::source_id().  The Python loader refuses a library whose id differs from the
 * sources beside it. */
const char* gloria_b200_build_id(void);
const char* gloria_b200_last_error(void);
/* number of kernels launched by this thread's calls since the last reset (bench.py's gpu_launches) */
long long gloria_b200_launch_count(int reset);

/* Measurement hook (bench.py's live roofline timing): record the caller's cudaEvent_t `start_event` immediately
 * before and `stop_event` immediately after every launch of the named kernel, on the launch stream.  Pass NULLs to
 * clear.  Events stay owned by the caller. */
#define GLORIA_TIMER_TC_FWD 0        /* fused tcgen05 forward kernel                 */
#define GLORIA_TIMER_TC_BWD_PAIR 1   /* fused tcgen05 backward (per-pair recompute) */
#define GLORIA_TIMER_TC_BWD_GEMM 2   /* backward accumulation GEMMs                  */
#define GLORIA_TIMER_SLOTS 4
int gloria_b200_set_timer_events(int slot, void* start_event, void* stop_event);
/* cudaEventRecord(event, stream) for callers that hold only raw handles (see gloria_b200_tc_local_sim_bwd_train_ev). */
int gloria_b200_record_event(void* event, void* stream);
/* n host int32 values (caption lengths) -> device array, passed through the kernel parameter buffer (960 per launch):
 * asynchronous on `stream`, no copy engine, the host array may be reused as soon as the call returns. */
int gloria_b200_upload_ints(const int32_t* host, int n, int32_t* dev, void* stream);
/* Development aid (builds with -DGLORIA_PHASE_CLOCKS only): device buffer receiving 8 int64 phase clocks per CTA. */
void gloria_b200_debug_phase_clocks(void* device_buffer);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 mode (CUDA-core FFMA, fp32 accumulate): replaces attention_fn + cosine_similarity + the caption loop of
 * local_loss (gloria_loss.py:11-63, 116-162) for the un-autocast reference, logits within 1e-5 relative.
 * ---------------------------------------------------------------------------------------------------------- */

/* Bytes of workspace wanted for Bi images x Bc captions; `budget` caps it (captions are then processed in
 * chunks).  Always >= the minimum for one caption per chunk. */
size_t gloria_b200_local_f32_workspace(int Bi, int Bc, int D, int S, int Lw, int Lcap, size_t budget);

/* Forward.  Optional outputs (may be NULL):
 *   attn_diag [Bc, Lcap, S]  attention map A of the diagonal pair (i, i)   (att_maps, gloria_loss.py:141-143);
 *                            rows l >= cap_len are zero-filled.  Requires Bi == Bc.
 *   attn_mean [Bi, Bc, S]    word-mean attention of every pair (flattened_attn, gloria_loss.py:132).
 * Lcap is an upper bound on cap_lens (<= Lw - word_off).  */
int gloria_b200_local_sim_fwd_f32(const float* ctx, const float* words, const int32_t* cap_lens,
                                  int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                  float temp1, float temp2, int agg, float eps,
                                  float* sim, float* attn_diag, float* attn_mean,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above by recomputation (no activations are kept between forward and backward).
 *   dsim [Bi, Bc]; d_attn_diag [Bc, Lcap, S] or NULL; d_attn_mean [Bi, Bc, S] or NULL.
 *   d_ctx [Bi, D, S] and d_words [Bc, D, Lw] are fully overwritten (padded word columns get exactly 0). */
int gloria_b200_local_sim_bwd_f32(const float* ctx, const float* words, const int32_t* cap_lens,
                                  int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                  float temp1, float temp2, int agg, float eps,
                                  const float* dsim, const float* d_attn_diag, const float* d_attn_mean,
                                  float* d_ctx, float* d_words,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 mode on the tensor cores (tc_f32.cu): the same two entry points with every bmm of gloria_loss.py:40,59 (and of
 * their autograd) on the hand-written tcgen05 GEMM with split-precision operands (three bf16 pieces per fp32 value,
 * six piece products per GEMM in the forward: everything down to 2^-24, fp32 accumulation in tensor memory), softmaxes /
 * cosine / aggregation as streaming fp32 kernels.  Same arguments, results and gates as the _f32 entries above
 * (logits within 1e-5 relative).  `backward` selects the workspace layout of the backward (it holds more buffers).
 * gloria_b200_f32tc_supported: 0 when the shape is covered (D % 64 == 0, cap_len <= 128), else the _f32 entries serve it.
 * ---------------------------------------------------------------------------------------------------------- */
int gloria_b200_f32tc_supported(int D, int S, int Lcap);
size_t gloria_b200_local_f32tc_workspace(int Bi, int Bc, int D, int S, int Lw, int Lcap, size_t budget, int backward);
int gloria_b200_local_sim_fwd_f32tc(const float* ctx, const float* words, const int32_t* cap_lens,
                                    int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                    float temp1, float temp2, int agg, float eps,
                                    float* sim, float* attn_diag, float* attn_mean,
                                    void* workspace, size_t workspace_bytes, void* stream);
/* Training forward: the workspace is laid out for the backward (gloria_b200_local_f32tc_workspace(..., backward = 1) with
 * budget 0 = all captions in one chunk; GLORIA_ERR_WORKSPACE otherwise) and is left holding the forward's state -- operand
 * pieces, P, A and the contexts.  The backward below then takes the same buffer with state_from_forward = 1 and recomputes
 * nothing; it only reads that state, so a forward can be differentiated more than once.  With state_from_forward = 0 the
 * backward recomputes the forward chunk by chunk inside any workspace (no activations kept). */
int gloria_b200_local_sim_fwd_f32tc_train(const float* ctx, const float* words, const int32_t* cap_lens,
                                          int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                          float temp1, float temp2, int agg, float eps,
                                          float* sim, float* attn_diag, float* attn_mean,
                                          void* workspace, size_t workspace_bytes, void* stream);
int gloria_b200_local_sim_bwd_f32tc(const float* ctx, const float* words, const int32_t* cap_lens,
                                    int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                    float temp1, float temp2, int agg, float eps,
                                    const float* dsim, const float* d_attn_diag, const float* d_attn_mean,
                                    float* d_ctx, float* d_words,
                                    void* workspace, size_t workspace_bytes, int state_from_forward, void* stream);

/* Diagonal pairs only (B pairs instead of B^2): the attention maps A_ii that local_loss returns (att_maps,
 * gloria_loss.py:141-143; consumed by get_attn_maps, gloria_model.py:209-211) and the gradient flowing back through
 * them (supervised-attention term of GLoRIA.calc_loss, gloria_model.py:143-147).  attn_diag / d_attn_diag are
 * [B, Lcap, S]; rows l >= cap_len are zero.  With accumulate != 0 the backward adds into d_ctx / d_words (which then
 * already hold the similarity gradients), otherwise it overwrites them. */
size_t gloria_b200_diag_attn_workspace(int B, int D, int S, int Lw, int Lcap);
int gloria_b200_diag_attn_fwd_f32(const float* ctx, const float* words, const int32_t* cap_lens,
                                  int B, int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                  float* attn_diag, void* workspace, size_t workspace_bytes, void* stream);
int gloria_b200_diag_attn_bwd_f32(const float* ctx, const float* words, const int32_t* cap_lens,
                                  int B, int D, int S, int Lw, int Lcap, int word_off, float temp1,
                                  const float* d_attn_diag, float* d_ctx, float* d_words, int accumulate,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* attention_fn (gloria_loss.py:19-63) as a stand-alone paired operator, for callers that use it directly
 * (get_local_similarities, gloria_model.py:185; Retriver, retrival_model.py:147): query[b] [D, L] attends to
 * context[b] [D, S]; wctx [B, D, L], attn [B, L, S].  The backward takes d_wctx and / or d_attn (NULL = none) and
 * overwrites d_query [B, D, L] and d_ctx [B, D, S]. */
size_t gloria_b200_attention_workspace(int B, int D, int S, int L);
int gloria_b200_attention_fwd_f32(const float* query, const float* ctx, int B, int D, int S, int L, float temp1,
                                  float* wctx, float* attn, void* workspace, size_t workspace_bytes, void* stream);
int gloria_b200_attention_bwd_f32(const float* query, const float* ctx, int B, int D, int S, int L, float temp1,
                                  const float* d_wctx, const float* d_attn, float* d_query, float* d_ctx,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* cosine_similarity (gloria_loss.py:11-16) of the rows of two [N, D] matrices; stats [N, 3] (dot, |x1|, |x2|) is
 * saved by the forward for the backward. */
int gloria_b200_row_cosine_fwd(const float* x1, const float* x2, long long N, int D, float eps, float* out,
                               float* stats, void* stream);
int gloria_b200_row_cosine_bwd(const float* x1, const float* x2, const float* stats, const float* dout,
                               long long N, int D, float eps, float* dx1, float* dx2, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * bf16 tensor-core mode (tcgen05 / TMEM / TMA): the fused hot path.  Same math as above with bf16 operands and
 * fp32 accumulation and softmax (the analogue of the reference under Lightning AMP).
 * ---------------------------------------------------------------------------------------------------------- */

/* Padded sizes used by the packed bf16 layouts (S -> next multiple of 128 strictly above S: at least one padded
 * region row exists, the backward's Gram matrix keeps its row of ones there; L -> multiple of 16). */
int gloria_b200_tc_spad(int S);
int gloria_b200_tc_lpad(int Lcap);
/* Row pitch per caption of words_t and column pitch of the backward's operand matrices: round_up(Lcap, 8) <= Lpad. */
int gloria_b200_tc_lp(int Lcap);
/* Row pitch per image of ctx_t and of the backward's operand matrices: round_up(S, 16) <= Spad. */
int gloria_b200_tc_sp(int S);
/* 0 if this (D, S, Lcap) is supported by the tensor-core kernels, else GLORIA_ERR_UNSUPPORTED. */
int gloria_b200_tc_supported(int D, int S, int Lcap);

/* Cast + transpose into TMA-legal 16-bit layouts (native row pitches 1444 B / 388 B are not 16 B multiples):
 *   ctx_h   [Bi, Spad, D] fp16  region-major copy (d contiguous), rows s >= S zero   -> score GEMM (A operand)
 *   ctx_t   [Bi, Sp, D]   bf16  same rows, pitch Sp = round_up(S, 16)                 -> backward GEMMs, Gram matrix
 *   ctx_n   [Bi, D, Spad] bf16  channel-major copy (s contiguous), cols s >= S zero   -> context GEMM (B operand)
 *   words_h [Bc, Lpad, D] fp16  word-major copy of columns [word_off, word_off+cap_len), other rows zero
 *   words_t [Bc, Lp, D]   bf16  same rows, pitch Lp = round_up(Lcap, 8)               -> backward GEMMs
 *   wnorm   [Bc, Lpad]    fp32  |W_l| computed from the fp32 input (gloria_loss.py:14)
 * The score GEMM (K = D, feeds the word softmax, which amplifies operand rounding) runs on fp16 operands -- the dtype
 * the reference's AMP runs this bmm in; everything after the softmax is bf16 x bf16.  All accumulate in fp32. */
int gloria_b200_tc_prepack(const float* ctx, const float* words, const int32_t* cap_lens,
                           int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                           void* ctx_h, void* ctx_t, void* ctx_n, void* words_h, void* words_t, float* wnorm,
                           void* stream);
/* The two halves of the prepack on their own (ctx_n may be NULL: only the inference forward and the recompute backward
 * read it; ctx_t may be NULL: only the training path reads it). */
int gloria_b200_tc_prepack_ctx(const float* ctx, int Bi, int D, int S, void* ctx_h, void* ctx_t, void* ctx_n,
                               void* stream);
int gloria_b200_tc_prepack_words(const float* words, const int32_t* cap_lens, int Bc, int D, int Lw, int Lcap,
                                 int word_off, void* words_h, void* words_t, float* wnorm, void* stream);

/* Fused forward over all Bi x Bc pairs: scores on tcgen05 (K = D), both softmaxes, attention-weighted context
 * on tcgen05 (K = S), per-word cosine and the temp2 log-sum-exp; only sim[Bi, Bc] reaches HBM -- plus, when
 * `stats` is given, the two per-word scalars the backward needs (stats [Bi, Bc, 2, Lpad]: <W_l, C'_l> and
 * |C'_l|^2 of the un-normalised context C' = Z_l C_l).  Inputs are the prepacked buffers. */
int gloria_b200_tc_local_sim_fwd(const void* ctx_h, const void* ctx_n, const void* words_h, const float* wnorm,
                                 const int32_t* cap_lens, int Bi, int Bc, int D, int S, int Lcap,
                                 float temp1, float temp2, int agg, float eps,
                                 float* sim, float* stats, void* stream);

/* Packed prompts (zero-shot scoring: gloria_model.py:171-207 as driven by gloria.py:278-306 -- thousands of images
 * against a few short class prompts).  Up to 8 captions of at most 16 words (after word_off) share one word tile: the
 * kernel's cost is set by the image tiles it streams per (image, word tile), so 25 prompts cost 4 tiles instead of 25.
 * Bc captions -> gloria_b200_tc_packed_groups(Bc) tiles of 16 * gloria_b200_tc_packed_per(Bc) words;
 * words_h [groups, 16 * per, D] fp16 and wnorm [groups, 16 * per] fp32 come from gloria_b200_tc_prepack_words_packed.
 * Forward only (agg = sum / mean / max); sim [Bi, Bc]. */
int gloria_b200_tc_packed_groups(int Bc);
int gloria_b200_tc_packed_per(int Bc);
int gloria_b200_tc_prepack_words_packed(const float* words, const int32_t* cap_lens, int Bc, int D, int Lw,
                                        int word_off, void* words_h, float* wnorm, void* stream);
int gloria_b200_tc_local_sim_fwd_packed(const void* ctx_h, const void* ctx_n, const void* words_h, const float* wnorm,
                                        const int32_t* cap_lens, int Bi, int Bc, int D, int S,
                                        float temp1, float temp2, int agg, float eps, float* sim, void* stream);

/* Workspace of the backward.  `budget` (0 = unlimited) caps it: captions are then processed in chunks.  The two
 * bf16 operand matrices take 2 * Bi * Sp * Lp * 2 bytes per caption (about 153 KB per pair at the full sizes:
 * 40 GB for B = 512 -- this is what the 180 GB of HBM3e are used for). */
size_t gloria_b200_tc_bwd_workspace(int Bi, int Bc, int D, int S, int Lcap, int have_stats, size_t budget);

/* Backward (agg = sum / mean).  A fused tcgen05 kernel recomputes scores and both softmaxes per pair, obtains
 * <C_l, R_s> from the image's Gram matrix (one more score-shaped GEMM, K = S) and writes the per-pair operand rows
 * X^T, E^T, (beta/Z^2) E^T; the sums over images / captions are plain GEMMs.  `stats` is the forward's output
 * (NULL: recomputed with one extra forward pass).  d_ctx [Bi, D, S] and d_words [Bc, D, Lw] fp32 in the callers'
 * native layouts are fully overwritten (padded word columns get exactly 0).
 * d_attn_mean (optional, [Bi, Bc, S]): gradient w.r.t. the word-mean attention of every pair
 * (gloria_b200_tc_local_sim_fwd_mean) -- the entropy / symmetric-KL / no-attn regularisers of
 * gloria/loss/gloria_loss.py:108-139,173-199 back-propagate through it. */
int gloria_b200_tc_local_sim_bwd(const void* ctx_h, const void* ctx_t, const void* ctx_n, const void* words_h,
                                 const void* words_t, const float* wnorm, const int32_t* cap_lens, const float* stats, int Bi, int Bc, int D, int S, int Lw,
                                 int Lcap, int word_off, float temp1, float temp2, int agg, float eps,
                                 const float* dsim, const float* d_attn_mean, float* d_ctx, float* d_words,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Forward with the word-mean attention of every pair: attn_mean[j, i, s] = (1/L_i) sum_l A[s, l]  ([Bi, Bc, S] fp32;
 * gloria/loss/gloria_loss.py:125-127 `attn.mean(1)` collected over the caption loop), sim as above, and (optional)
 * `stats` [Bi, Bc, 2, Lpad] for gloria_b200_tc_local_sim_bwd.  One tcgen05 kernel per call (scores, both softmaxes,
 * Gram-form |C'|^2, Z from the Gram matrix's row of ones); workspace = Gram matrices. */
size_t gloria_b200_tc_mean_workspace(int Bi, int Bc, int D, int S, int Lcap);
int gloria_b200_tc_local_sim_fwd_mean(const void* ctx_h, const void* ctx_t, const void* words_h, const float* wnorm,
                                      const int32_t* cap_lens, int Bi, int Bc, int D, int S, int Lcap,
                                      float temp1, float temp2, int agg, float eps, float* sim, float* attn_mean,
                                      float* stats, void* workspace, size_t workspace_bytes, void* stream);

/* Fused training path (used when the workspace fits): every backward quantity of a pair is linear in
 * g = dsim[j, i], so ONE kernel computes sim and, for g = 1, the backward operand rows (X^T, E^T, f, gamma) during the
 * forward; the backward is a scale by g plus the accumulation GEMMs -- nothing is recomputed.  The same workspace
 * (gloria_b200_tc_train_workspace bytes; 41.7 GB at B = 512) is passed to both calls and must stay untouched in between;
 * the backward consumes it (X is scaled in place), so it can run once per forward. */
size_t gloria_b200_tc_train_workspace(int Bi, int Bc, int D, int S, int Lcap);
int gloria_b200_tc_local_sim_fwd_train(const void* ctx_h, const void* ctx_t, const void* words_h, const float* wnorm,
                                       const int32_t* cap_lens, int Bi, int Bc, int D, int S, int Lcap,
                                       float temp1, float temp2, int agg, float eps, float* sim,
                                       void* workspace, size_t workspace_bytes, void* stream);
int gloria_b200_tc_local_sim_bwd_train(const void* ctx_t, const void* words_t, const int32_t* cap_lens,
                                       int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                       const float* dsim, float* d_ctx, float* d_words,
                                       void* workspace, size_t workspace_bytes, void* stream);
/* Image range [j0, j0 + nj) of a (Bi x Bc) training forward: the workspace is laid out for all Bi images and this call
 * fills the rows of the range (ctx_h, ctx_t and sim are the base pointers of the full arrays).  One launch per
 * all_gather chunk lets a caption-sharded caller overlap the gather of the next chunk with this chunk's kernel. */
int gloria_b200_tc_local_sim_fwd_train_part(const void* ctx_h, const void* ctx_t, const void* words_h,
                                            const float* wnorm, const int32_t* cap_lens, int Bi, int j0, int nj,
                                            int Bc, int D, int S, int Lcap, float temp1, float temp2, int agg,
                                            float eps, float* sim, void* workspace, size_t workspace_bytes,
                                            void* stream);
/* Same for callers whose images of the range live in a buffer of their own: range_h / range_t point at the packed copies of
 * image j0 (nj images follow contiguously).  flags: 1 = record the bench timer's start before the launch, 2 = its stop
 * after it, 4 = first launch of this forward (resets the state's consumed flag).  A caption-sharded caller computes its
 * own images from its local pack while the all_gather of the other ranks' packed copies is in flight. */
int gloria_b200_tc_local_sim_fwd_train_range(const void* range_h, const void* range_t, const void* words_h,
                                             const float* wnorm, const int32_t* cap_lens, int Bi, int j0, int nj,
                                             int Bc, int D, int S, int Lcap, float temp1, float temp2, int agg,
                                             float eps, float* sim, void* workspace, size_t workspace_bytes,
                                             int flags, float* attn_diag_raw, int diag_lcap, void* stream);
/* Training forward that also returns the attention maps of the diagonal pairs (att_maps of local_loss,
 * gloria_loss.py:141-143; needs Bi == Bc): attn_diag [Bc, diag_lcap, S] (diag_lcap >= Lcap), rows beyond each caption's
 * length zero.  The maps come out of the fused kernel's own softmax -- the diagonal pairs store their fp32 numerators,
 * one small launch normalises them -- instead of a second pass over the features.  (attn_diag_raw / diag_lcap of the
 * range variant above: the un-normalised buffer, or NULL.) */
int gloria_b200_tc_local_sim_fwd_train_diag(const void* ctx_h, const void* ctx_t, const void* words_h,
                                            const float* wnorm, const int32_t* cap_lens, int Bi, int Bc, int D, int S,
                                            int Lcap, float temp1, float temp2, int agg, float eps, float* sim,
                                            void* workspace, size_t workspace_bytes, float* attn_diag, int diag_lcap,
                                            void* stream);
/* Backward with the image side done in n_parts equal image ranges (n_parts divides Bi); part_events[k] (cudaEvent_t or
 * NULL; the array itself may be NULL) is recorded on `stream` once the d_ctx rows of part k are final.  The
 * caption-side GEMM runs last.  d_ctx may be NULL: the image-side gradient is then left in the workspace in its packed
 * form dRt [Bi, sp, D] fp32 at gloria_b200_tc_train_drt_offset (final for part k at part_events[k]) and the caller
 * unpacks the images it wants with gloria_b200_tc_unpack_dctx -- a caption-sharded caller reduce_scatters the packed
 * rows and transposes only its own images (autograd of the all_gather in SURVEY 8e). */
int gloria_b200_tc_local_sim_bwd_train_parts(const void* ctx_t, const void* words_t, const int32_t* cap_lens,
                                             int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                             const float* dsim, float* d_ctx, float* d_words,
                                             void* workspace, size_t workspace_bytes, int n_parts,
                                             void* const* part_events, void* stream);
size_t gloria_b200_tc_train_drt_offset(int Bi, int Bc, int D, int S, int Lcap);
/* dRt [n, sp, D] fp32 -> d_ctx [n, D, S] (the layout of img_emb_l's gradient); D % 32 == 0. */
int gloria_b200_tc_unpack_dctx(const float* drt, float* d_ctx, int n, int D, int S, void* stream);
/* Same, with a caller-owned cudaEvent_t (or NULL) that is recorded on `stream` as soon as d_ctx is final -- before the
 * caption-side GEMM.  A caption-sharded caller waits on it to start the reduce_scatter of d_ctx while the
 * rest of the backward still runs. */
int gloria_b200_tc_local_sim_bwd_train_ev(const void* ctx_t, const void* words_t, const int32_t* cap_lens,
                                          int Bi, int Bc, int D, int S, int Lw, int Lcap, int word_off,
                                          const float* dsim, float* d_ctx, float* d_words,
                                          void* workspace, size_t workspace_bytes, void* d_ctx_ready_event,
                                          void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Accumulation GEMM of the backward (the sums over images / captions that autograd forms for the two bmm's of
 * attention_fn, gloria_loss.py:40,59): C[M, N] (fp32, row-major) = or += A B, bf16 operands in device memory, on a
 * persistent CTA-pair tcgen05 kernel (csrc/tc_gemm.cu).  Exported for the parity tests and for callers that want the
 * primitive alone; the training backward calls it internally.
 *   a_kmajor != 0: A is [M, K] row-major; else A is given transposed, [K, M] row-major.  B is [K, N] row-major.
 *   N % 16 == 0, K % 8 == 0 (and M % 8 == 0 for a transposed A).
 *   ksplit: 0 = library's choice, k > 1 cuts K in k parts whose tiles are added with red.global.add.
 *   accumulate != 0: C += A B (C holds the value to add to); otherwise C is overwritten.
 *   g (may be NULL): A[m, k] is multiplied by g[(m / m_div) * g_sm + (k / k_div) * g_sk] in fp32 and rounded to bf16 on
 *     its way to the tensor cores (through tensor memory for a row-major A, in place in shared memory for a transposed
 *     A); the divisor along the contiguous axis of A (k_div, resp. m_div) must be a multiple of 8.
 *   force_scaled_path: route an unscaled A through the scaling pipeline as well (weight 1; test hook).
 * ---------------------------------------------------------------------------------------------------------- */
int gloria_b200_acc_gemm(const void* A, const void* B, float* C, int M, int N, int K, int a_kmajor, int ksplit,
                         int accumulate, const float* g, int g_sm, int g_sk, int m_div, int k_div,
                         int force_scaled_path, void* stream);
/* The same GEMM with split-precision operands (the fp32 tensor-core mode): A and B are given as three dense planes of bf16
 * pieces of fp32 matrices (plane 0 = bf16(x), 1 = bf16(x - p0), 2 = bf16(x - p0 - p1)); nterms = 6 sums every piece
 * product down to 2^-24, nterms = 3 the three leading ones; nb independent problems are stacked along the rows of every
 * plane (A [3][nb*M, K] or, transposed, [3][nb*K, M]; B [3][nb*K, N]; C [nb][M, N]). */
int gloria_b200_acc_gemm_planes(const void* A, const void* B, float* C, int M, int N, int K, int a_kmajor, int nterms,
                                int nb, int ksplit, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Global similarity (global_loss, gloria_loss.py:75-80; get_global_similarities, gloria_model.py:164-169):
 *   cosm[a, b] = <x_a, y_b> / max(|x_a| |y_b|, eps),  x [Bi, D], y [Bc, D];  xn [Bi], yn [Bc] are saved norms.
 * ---------------------------------------------------------------------------------------------------------- */
int gloria_b200_global_sim_fwd(const float* x, const float* y, int Bi, int Bc, int D, float eps,
                               float* cosm, float* xn, float* yn, void* stream);
int gloria_b200_global_sim_bwd(const float* x, const float* y, const float* xn, const float* yn,
                               const float* dcos, int Bi, int Bc, int D, float eps,
                               float* dx, float* dy, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Bidirectional cross entropy with labels = arange(B) (gloria_loss.py:86-87 and :164-170):
 *   logits = scale * m (m [B, B]);  losses[0] = CE(logits), losses[1] = CE(logits^T), mean reduction.
 *   row_lse / col_lse [B] are saved for the backward.  g [2] = upstream gradients of the two losses (device).
 * ---------------------------------------------------------------------------------------------------------- */
int gloria_b200_ce_bidir_fwd(const float* m, int B, float scale, float* losses, float* row_lse, float* col_lse,
                             void* stream);
int gloria_b200_ce_bidir_bwd(const float* m, int B, float scale, const float* row_lse, const float* col_lse,
                             const float* g, float* dm, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Word-piece aggregation of the text encoder (BertEncoder.aggregate_tokens, gloria/models/text_model.py:32-90): the
 * step right before the loss path.  caption_ids [B, T] int64 (device); is_continuation [vocab] = 1 where the
 * vocabulary entry starts with "##"; sep_id = id of "[SEP]".
 *   gloria_b200_word_ranges: word_range [B, T, 2] (first token, one past the last token of word w; (0,0) beyond the
 *     caption's words), token_word [B, T] (word of token t, -1 if the token belongs to no emitted word), n_words [B].
 *   gloria_b200_aggregate_tokens_fwd: out[b, layer, w, :] = sum of embeddings[b, layer, t, :] over the tokens of word w,
 *     zero rows beyond n_words[b]; embeddings / out [B, layers, T, D] contiguous, dtype one of GLORIA_DTYPE_*
 *     (fp32 accumulation, one rounding).
 *   gloria_b200_aggregate_tokens_bwd: d_embeddings[b, layer, t, :] = d_out[b, layer, token_word[b, t], :] (or 0).
 * ---------------------------------------------------------------------------------------------------------- */
#define GLORIA_DTYPE_F32 0
#define GLORIA_DTYPE_F16 1
#define GLORIA_DTYPE_BF16 2
int gloria_b200_word_ranges(const long long* caption_ids, const unsigned char* is_continuation, int vocab,
                            long long sep_id, int B, int T, int32_t* word_range, int32_t* token_word,
                            int32_t* n_words, void* stream);
/* Same, plus the caption lengths the loss derives from the word strings (gloria_model.py:107-109: words that do not
 * start with '[' plus one) computed on the device: is_bracket [vocab] = 1 where the entry's text ("##" stripped) starts
 * with '['.  cap_lens [B] feeds local_loss / calc_loss without a host round trip. */
int gloria_b200_word_ranges_cap_lens(const long long* caption_ids, const unsigned char* is_continuation,
                                     const unsigned char* is_bracket, int vocab, long long sep_id, int B, int T,
                                     int32_t* word_range, int32_t* token_word, int32_t* n_words, int32_t* cap_lens,
                                     void* stream);
int gloria_b200_aggregate_tokens_fwd(const void* embeddings, int dtype, const int32_t* word_range, int B, int layers,
                                     int T, int D, void* out, void* stream);
int gloria_b200_aggregate_tokens_bwd(const void* d_out, int dtype, const int32_t* token_word, int B, int layers,
                                     int T, int D, void* d_embeddings, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLORIA_B200_H_ */
