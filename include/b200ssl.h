/*
 * b200ssl.h -- C ABI of libb200ssl.so: the FixMatch / CoMatch unlabeled-consistency
 * head and the EMA weight update as hand-written sm_100a CUDA kernels.
 *
 * The reference (taindp98/Endoscopy-Image-Classification) has no FFI of its
 * own: its boundary for this path is the Python call surface of code/loss.py,
 * code/ema.py and the inline head in code/comatch.py.  Each entry point below
 * replaces one eager-op sequence of that surface; the file:line it replaces is
 * cited on the declaration (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - return 0 on success, <0 for a rejected argument (B200SSL_E_*), >0 is a
 *     cudaError_t from the launch; b200ssl_last_error_string() describes it;
 *   - never allocate, never synchronise, never throw; re-entrant across streams
 *     as long as each stream uses its own workspace;
 *   - every pointer is a BORROWED DEVICE pointer (tensor.data_ptr()); matrices
 *     are dense row-major; `dtype` selects the storage type of the logits /
 *     embeddings / bank (B200SSL_F32 or B200SSL_BF16); all arithmetic is fp32;
 *   - scalars results (losses, means) are written to device memory;
 *   - `workspace` must be zero-filled once when it is allocated (it holds
 *     self-resetting ticket counters), >= b200ssl_workspace_bytes() long and
 *     256-byte aligned;
 *   - `stream` is a cudaStream_t passed as void*.
 */
#ifndef B200SSL_H_
#define B200SSL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SSL_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define B200SSL_API __attribute__((visibility("default")))
#else
#define B200SSL_API
#endif

enum b200ssl_dtype { B200SSL_F32 = 0, B200SSL_BF16 = 1, B200SSL_F16 = 2, B200SSL_I64 = 3, B200SSL_I32 = 4, B200SSL_U8 = 5 };

enum b200ssl_error {
  B200SSL_OK = 0,
  B200SSL_E_NULL = -1,      /* required pointer is NULL */
  B200SSL_E_SHAPE = -2,     /* rows/classes/dim outside the supported range */
  B200SSL_E_DTYPE = -3,     /* unsupported dtype for this entry point */
  B200SSL_E_ALIGN = -4,     /* pointer not aligned as required */
  B200SSL_E_WORKSPACE = -5, /* workspace too small */
  B200SSL_E_ARG = -6        /* other invalid argument */
};

/* limits */
#define B200SSL_MAX_CLASSES 1024
#define B200SSL_MAX_EMB_DIM 256
#define B200SSL_DA_WINDOW_MAX 64

B200SSL_API int b200ssl_version(void);
B200SSL_API const char* b200ssl_last_error_string(void);
/* Debug aid (tools/kernel_timeline.py): device buffer of uint64 [ctas * 16] that instrumented
 * kernels fill with clock64() stamps (4 regions of 4096*16, one per kernel); NULL (default) disables it. */
B200SSL_API void b200ssl_debug_set_timing_buffer(void* device_u64);
/* bytes of scratch any single call below may need for the given problem */
B200SSL_API size_t b200ssl_workspace_bytes(int64_t rows, int32_t classes, int64_t bank_rows);

/* ---------------------------------------------------------------- K1 ----
 * FixMatch unlabeled head, forward + backward in one launch.
 * Replaces code/loss.py:126-164 (consistency_loss, name='ce') incl. the
 * F.cross_entropy at :119 and its autograd backward; called from
 * code/fixmatch.py:116 and twice per step from code/semiformer.py:129-130
 * (pass logits_s2/grad_s2 to serve both strong heads with one read of w).
 *   p = softmax(w); (pmax, idx) = max(p) [first maximal index];
 *   mask = pmax >= p_cutoff;  loss = mean(CE(s, idx) * mask)
 *   grad_s = mask/rows * (softmax(s) - onehot(idx))            (hard labels)
 *   soft labels (use_hard_labels=0): targets softmax(w*inv_T).
 * out_scalars[0]=loss, [1]=mask mean, [2]=loss of the second strong head.
 * idx (int64[rows]) and mask (f32[rows]) are optional outputs.
 */
B200SSL_API int b200ssl_fixmatch_head_fwd_bwd(const void* logits_w, const void* logits_s, const void* logits_s2,
                                  void* grad_s, void* grad_s2, int64_t rows, int32_t classes,
                                  int32_t dtype, float p_cutoff, float inv_T, int32_t use_hard_labels,
                                  float* out_scalars, int64_t* idx, float* mask, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* grad[i] *= (*scale) * factor  -- chains the stashed gradient with autograd's
 * upstream gradient (device scalar) and a host constant (e.g. LAMBDA_U at
 * code/fixmatch.py:118). */
B200SSL_API int b200ssl_scale_inplace(void* grad, int64_t numel, int32_t dtype, const float* scale, float factor,
                                      void* stream);

/* ------------------------------------------------------------ f2 (next) --
 * Labeled-branch criterion: class-weighted CE or Poly-1 CE, forward+backward.
 * Replaces code/loss.py:103-119 + PolyLoss code/loss.py:308-364 (called at
 * code/fixmatch.py:114, code/comatch.py:156-160, code/semiformer.py:126-128).
 *   row_i = w[y_i]*CE_i + epsilon*(1 - softmax(x_i)[y_i])
 *   poly=1: loss = mean_i(row_i)               (plain mean, loss.py:355-356)
 *   poly=0: loss = sum_i w[y_i]*CE_i / sum_i w[y_i]   (F.cross_entropy 'mean')
 * class_weights may be NULL.  out_scalar[0] = loss.
 */
B200SSL_API int b200ssl_labeled_ce_fwd_bwd(const void* logits, const int64_t* targets, const float* class_weights,
                               void* grad, int64_t rows, int32_t classes, int32_t dtype, int32_t poly,
                               float epsilon, float* out_scalar, void* workspace, size_t workspace_bytes,
                               void* stream);

/* Un-reduced labeled criterion (code/loss.py:118-124; PolyLoss reduction='none', :357-359): loss_rows[i] and, in
 * grad_unit [rows, classes], the gradient of loss_rows[i] w.r.t. logits row i.  Exactly one of `targets` (int64 hard
 * labels; class_weights / poly as in b200ssl_labeled_ce_fwd_bwd) and `soft_targets` (fp32 [rows, classes];
 * loss.py:120-124: sum_c -t_c log_softmax(x)_c) is given.  Labels: -100 (F.cross_entropy's ignore_index) drops the
 * row; any other label outside [0, classes) also drops it and raises the sticky counter read by
 * b200ssl_bad_label_count (both labeled kernels) instead of reading out of bounds. */
B200SSL_API int b200ssl_ce_rows_fwd_bwd(const void* logits, const int64_t* targets, const float* soft_targets,
                            const float* class_weights, void* grad_unit, float* loss_rows, int64_t rows,
                            int32_t classes, int32_t dtype, int32_t poly, float epsilon, void* workspace,
                            size_t workspace_bytes, void* stream);

/* grad_out[i, :] = grad_in[i, :] * row_scale[i]  (autograd's per-row upstream gradient; out of place). */
B200SSL_API int b200ssl_scale_rows(const void* grad_in, void* grad_out, int64_t rows, int32_t classes, int32_t dtype,
                       const float* row_scale, void* stream);

/* Number of out-of-range labels the labeled kernels have seen on this workspace (synchronises; reset != 0 clears it). */
B200SSL_API int b200ssl_bad_label_count(void* workspace, uint32_t* count, int32_t reset);

/* ------------------------------------------------------------ f4 (next) --
 * Evaluation head: the per-batch tail of evaluate_one (code/fixmatch.py:154-168) in one launch.
 *   loss_out[0]          = F.cross_entropy(logits, targets, reduction='mean') of this batch (fixmatch.py:156)
 *   confusion[t*C + p]  += 1 for every row with target t and prediction p = argmax softmax(logits) (first index on
 *                          ties; :160,166) -- uint64 [classes, classes], accumulated over the batches of a pass
 *   pred[i]              = p (optional; classification_report)
 * Every figure of utils.calculate_metrics (code/utils.py:38-55) is a function of the confusion matrix, so a pass ends
 * with one device-to-host copy.  Labels as in b200ssl_ce_rows_fwd_bwd (-100 ignored, out of range counted and dropped). */
B200SSL_API int b200ssl_eval_head(const void* logits, const int64_t* targets, int64_t rows, int32_t classes, int32_t dtype,
                      uint64_t* confusion, float* loss_out, int64_t* pred, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---------------------------------------------------------------- K2 ----
 * CoMatch distribution alignment statistics.  Replaces code/comatch.py:167-173:
 * softmax(logits_u_w).mean(0) is pushed on a device-resident history ring
 * (da_ring f32[window*classes], da_state int32[2] = {count, head}) and
 * prob_avg f32[classes] = mean over the history, summed oldest -> newest.
 * col_mean_out (optional f32[classes]) receives this batch's column mean.
 */
B200SSL_API int b200ssl_comatch_da(const void* logits_u_w, int64_t rows, int32_t classes, int32_t dtype,
                       float* da_ring, int32_t* da_state, int32_t window, float* prob_avg,
                       float* col_mean_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------ SURVEY 8e ----
 * Memory bank of a multi-rank job, resident in NVLink peer memory (one node).
 * The reference keeps one bank per process and has no distributed code
 * (code/comatch.py:90-96); here every rank maps every rank's peer arena (see the
 * b200ssl_peer_* functions at the end of this file) and the bank lives inside the arenas:
 *   sharded     rank s keeps global rows [s*shard_rows, (s+1)*shard_rows); K3 reads ALL
 *               shards in place (TMA over NVLink for the remote ones) and the enqueue
 *               stores each row into the shard that owns it;
 *   replicated  every arena holds the whole ring (shard_rows = all rows); K3 reads only
 *               the local copy and the enqueue writes this rank's rows through into
 *               every copy -- the only NVLink traffic is the enqueued rows.
 * Either way a step has the launches of the single-GPU step plus one enqueue launch on
 * a side stream, and no collective.  Two epoch flags per step keep the ring consistent:
 *   - b200ssl_bank_smooth_partial(shards != NULL) waits until every rank's enqueue of
 *     the previous step is visible, and publishes "my reads are done" when it ends;
 *   - b200ssl_bank_enqueue_peer waits for every rank's "reads done" of this step
 *     before it writes rank `rank`'s block at global rows [ptr + rank*n, ptr +
 *     (rank+1)*n) (n = rows + n_x; ptr advances by world*n), then publishes "my rows
 *     are in".
 * Both flags have most of a step of slack, so the waits are normally free.
 * Requirements: bf16 bank, dim 64, classes <= 31, shard_rows a multiple of 8,
 * world <= 8; all ranks run the same call sequence, one enqueue per smoothing pass.
 */
typedef struct b200ssl_bank_shards {
  int32_t world, rank;
  int64_t shard_rows;             /* rows per shard; the global bank has world*shard_rows rows */
  const uint64_t* arenas_host;    /* host array [world]: arena base addresses as mapped in this process */
  void* const* arenas_dev;        /* the same table in device memory */
  uint64_t feats_offset;          /* byte offsets, identical in every arena, of queue_feats [shard_rows, 64], */
  uint64_t probs_offset;          /*   queue_probs [shard_rows, classes] and                                 */
  uint64_t probs_t_offset;        /*   queue_probs_t [32, shard_rows] (row `classes` = ones)                 */
  int32_t replicated;             /* != 0: every arena holds the whole ring (shard_rows = all bank rows) */
  int32_t reserved;
} b200ssl_bank_shards;

/* ---------------------------------------------------------------- K3 ----
 * Memory-smoothing partial sums against (a shard of) the bank.  Replaces the
 * two GEMMs + exp + row-sum of code/comatch.py:180-181 without materialising
 * A[rows, bank_rows]:
 *   rowsum[i]   = sum_k exp(<f_i, q_k> / temperature)
 *   numer[i,c]  = sum_k exp(<f_i, q_k> / temperature) * queue_probs[k,c]
 * feats / queue_* share `dtype`; rowsum/numer are fp32.  No running max (the
 * reference has none; |<f,q>|/tau <= 5 for unit-norm embeddings).
 * Code paths, chosen by storage type and shape (not a backend switch):
 *   - bf16 bank, dim == 64, classes <= 31, bank_rows % 8 == 0 and queue_probs_t
 *     given (bf16 [32, bank_rows]: rows 0..classes-1 = transposed copy of queue_probs
 *     maintained by b200ssl_bank_enqueue, row `classes` = all ones so that the second
 *     MMA also produces the row sums, remaining rows zero): tcgen05.mma with TMEM
 *     accumulators, operands staged by TMA, the exponentials kept in tensor memory as
 *     the A operand of the second MMA, split partials folded through thread-block
 *     cluster distributed shared memory (csrc/bank_tc.cu);
 *   - fp32 bank, dim == 64, classes <= 31 (any bank_rows; queue_probs_t unused): a pre-pass
 *     splits feats, queue_feats and queue_probs into bf16 hi + mid operands inside the
 *     workspace and the same tensor-core kernel runs three cross-term MMA groups per GEMM
 *     -- 1e-6 of fp64 on numer / rowsum (the reference's fp32 product: 3e-7);
 *   - otherwise exact-fp32 FFMA tiles (the reference's matrix product is true fp32).
 * b200ssl_workspace_bytes(rows, classes, bank_rows) covers the operand copies of the second form.
 * rowsum_ld / numer_ld are the row strides of the two outputs in floats (0 = dense: 1 and
 * `classes`); a packed [rows, W] buffer (numer at column 0, rowsum at column `classes`,
 * both strides W) lets the sharded bank reduce-scatter both with one collective.
 */
B200SSL_API int b200ssl_bank_smooth_partial(const void* feats_u_w, const void* queue_feats, const void* queue_probs,
                                const void* queue_probs_t, int64_t rows, int64_t bank_rows, int32_t dim, int32_t classes,
                                int32_t dtype, float temperature, float* rowsum, float* numer,
                                int32_t rowsum_ld, int32_t numer_ld, const struct b200ssl_bank_shards* shards,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------- K2b+K4+K7 ------
 * Per-row finalisation of the CoMatch pseudo-label and the focal soft-CE.
 * Replaces code/comatch.py:163,174-176 (DA divide + renormalise), :182
 * (alpha mix; rowsum/numer NULL => smoothing off), :184-185 (max / mask) and
 * :216-220 (focal soft cross-entropy) incl. its autograd backward:
 *   probs_orig = renorm(softmax(w) / prob_avg)
 *   probs      = alpha*probs_orig + one_minus_alpha*numer/rowsum
 *                (factors rounded to fp32 by the caller: (float)a, (float)(1.0-a))
 *   (scores, lbs) = max(probs);  mask = scores >= thr
 *   logp = -sum_c log_softmax(s0)_c*probs_c * mask;  p = exp(-logp)
 *   loss_u = mean((1-p)^gamma * logp)
 *   grad_s0 = dl * mask * (softmax(s0)*sum_c probs_c - probs),
 *             dl = (gamma*(1-p)^(gamma-1)*p*logp + (1-p)^gamma)/rows
 * probs / probs_orig are fp32 [rows, classes]; out_scalars[0]=loss_u,
 * [1]=mask mean.  scores/lbs/mask optional.  probs_hl (optional, classes <= 32)
 * receives bf16 [rows, 64] = [hi(probs) padded to 32 | lo = probs - hi padded to 32],
 * the operand of the tensor-core graph kernel (b200ssl_contrast_*).
 */
B200SSL_API int b200ssl_comatch_finalize(const void* logits_u_w, const void* logits_u_s0, const float* prob_avg,
                             const float* rowsum, const float* numer, int32_t rowsum_ld, int32_t numer_ld,
                             int64_t rows, int32_t classes,
                             int32_t dtype, float alpha, float one_minus_alpha, float thr, float gamma,
                             float* probs, float* probs_orig, void* probs_hl, float* scores, int64_t* lbs,
                             float* mask, void* grad_s0, float* out_scalars, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------- K2 + K2b/K4/K7 (+ K5) fused ----
 * The whole row phase of a CoMatch step in ONE thread-block-cluster launch, for small
 * batches (classes <= 32; intended for rows <= ~2048, larger batches use the separate
 * kernels above): b200ssl_comatch_da, then b200ssl_comatch_finalize, then -- when
 * queue_feats != NULL -- b200ssl_bank_enqueue of the unsharded bank (device write pointer
 * ptr_state, advanced by rows + n_x).  Same arithmetic, same outputs; the cross-row
 * reductions (DA column means, loss, mask mean) are exchanged through distributed shared
 * memory in rank order instead of global tickets.  Replaces code/comatch.py:163-176,
 * 182-196, 216-220.  onehot_tail != 0: probs_orig has rows + n_x rows and the tail receives
 * onehot(targets_x), i.e. the buffer becomes the probability block of the enqueue
 * (comatch.py:188-189) that a sharded bank all-gathers.
 */
B200SSL_API int b200ssl_comatch_rows_fused(const void* logits_u_w, const void* logits_u_s0, const float* rowsum,
                               const float* numer, int32_t rowsum_ld, int32_t numer_ld, int64_t rows,
                               int32_t classes, int32_t dtype, float alpha,
                               float one_minus_alpha, float thr, float gamma, float* da_ring, int32_t* da_state,
                               int32_t window, float* prob_avg, float* probs, float* probs_orig, void* probs_hl,
                               float* scores, int64_t* lbs, float* mask, void* grad_s0, float* out_scalars,
                               void* queue_feats, void* queue_probs, void* queue_probs_t, const void* feats_u_w,
                               const void* feats_x, const int64_t* targets_x, int64_t n_x, int32_t dim,
                               int64_t* ptr_state, int64_t bank_rows, int32_t onehot_tail, void* stream);

/* ---------------------------------------------------------------- K5 ----
 * Ring-buffer enqueue.  Replaces code/comatch.py:187-196: rows are
 * [unlabeled-weak (n_u) ; labeled (n_x)], probabilities [probs_orig ; onehot(targets_x)],
 * written at global rows (ptr + r) mod bank_rows_global.  Only rows that fall
 * into the local shard [shard_begin, shard_begin+shard_rows) are written (the
 * whole bank for shard_begin=0, shard_rows=bank_rows_global).  `block_offset`
 * is this block's row offset inside a multi-rank step (rank * n), 0 otherwise.
 * queue_probs_t (optional, may be NULL) is the transposed class-padded copy
 * [32, shard_rows] read by the tensor-core smoothing kernel.
 * The write pointer is either the host value `ptr` (ptr_state == NULL) or, for
 * CUDA-graph replay, device resident: ptr_state = int64[2] {write pointer, ticket
 * (zero-initialised)}; the kernel then ignores `ptr`, and after all rows are
 * written advances the device pointer by `advance` rows mod K (0 = leave it, for
 * all but the last block of a multi-rank step).
 */
B200SSL_API int b200ssl_bank_enqueue(void* queue_feats, void* queue_probs, void* queue_probs_t, const void* feats_u_w,
                         const void* feats_x,
                         const float* probs_orig, const int64_t* targets_x, int64_t n_u, int64_t n_x,
                         int32_t dim, int32_t classes, int32_t dtype, int64_t ptr, int64_t* ptr_state,
                         int64_t advance, int64_t block_offset, int64_t bank_rows_global,
                         int64_t shard_begin, int64_t shard_rows, void* stream);

/* ---------------------------------------------------------------- K6 ----
 * Graph-contrastive loss.  Replaces code/comatch.py:199-213 and its autograd
 * backward (closed form):
 *   P = rowsoftmax-without-max(F0 F1^T / tau);  Q = probs probs^T, diag 1,
 *   thresholded at contrast_th, row-normalised;  loss = mean_i(-sum_j log(P+1e-7) Q)
 * fwd writes loss to out_scalar[0]; when total_out != NULL it also writes
 * total_out[0] = lambda_u * loss_u[0] + lambda_c * loss (comatch.py:222 without
 * loss_x; loss_u is the device scalar produced by b200ssl_comatch_finalize).
 * Code paths by storage type and shape: bf16 embeddings with dim == 64, classes <= 32 and
 * probs_hl given run on tcgen05/TMEM/TMA (csrc/contrast_tc.cu, S = F0 F1^T and the
 * hi/lo-split Q = probs probs^T as MMAs, dZ staged through swizzled shared memory for the
 * two gradient GEMMs; the forward is ONE pass over the S/Q tiles: its second pass is closed
 * in per-row moments); fp32 embeddings with dim == 64, classes <= 32 run the
 * same kernels on bf16 hi / mid / lo operands split into the workspace by a pre-pass (probs_hl
 * unused; pairs of the pseudo-label graph take Q from the fp32 probabilities; 1e-5 parity);
 * everything else uses exact-fp32 FFMA tiles (csrc/contrast.cu).
 * fwd also stores the row statistics (rowsum, qsum, r) into
 * stats f32[3*rows]; bwd consumes them and writes grad_f0 / grad_f1 scaled by
 * (*upstream) * factor (upstream: device scalar from autograd, NULL => 1; factor:
 * host constant such as LAMBDA_C of comatch.py:222).  Optional piggy-back (scale_grad != NULL):
 * the same launch also performs scale_grad[i] *= (*scale_upstream) * scale_factor, i.e. the
 * b200ssl_scale_inplace of the stashed focal-CE gradient (saves one launch per backward).
 */
B200SSL_API int b200ssl_contrast_fwd(const void* feats_s0, const void* feats_s1, const float* probs, const void* probs_hl,
                         int64_t rows,
                         int32_t dim, int32_t classes, int32_t dtype, float temperature, float contrast_th,
                         float* stats, float* out_scalar, const float* loss_u, float lambda_u, float lambda_c,
                         float* total_out, void* workspace, size_t workspace_bytes, void* stream);
B200SSL_API int b200ssl_contrast_bwd(const void* feats_s0, const void* feats_s1, const float* probs, const void* probs_hl,
                         const float* stats, int64_t rows, int32_t dim, int32_t classes, int32_t dtype, float temperature,
                         float contrast_th, const float* upstream, float factor, void* grad_f0, void* grad_f1,
                         void* scale_grad, int64_t scale_numel, const float* scale_upstream, float scale_factor,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- K8 ----
 * Multi-tensor EMA.  Replaces the per-tensor loop of code/ema.py:51-59
 * (update) and :61-62 (set) with one launch over a device-resident table.
 *   update: e <- fl(fl(d*e) + fl((1-d)*m)), each op rounded in the tensor's
 *   dtype like the eager mul, mul, add (no FMA contraction); applied `repeat`
 *   times when the same storage appears several times in state_dict()
 *   (custom_model.py:194-200); int64 buffers are computed in fp32 and truncated.
 * The table is device resident: one 32-byte row per chunk (<= 4096 elements is
 * the shipped host policy) of one tensor, carrying that chunk's two base
 * pointers so a CTA needs one dependent load before it streams.  It must be
 * 16-byte aligned.  decay / one_minus_decay are (float)d and (float)(1.0-d).
 */
typedef struct b200ssl_ema_block {
  void* ema;          /* first element of the chunk in the EMA tensor */
  const void* model;  /* first element of the chunk in the live model tensor */
  int32_t count;      /* elements in this chunk */
  int32_t dtype;      /* enum b200ssl_dtype (F32, BF16, F16, I64, I32, U8) */
  int32_t repeat;     /* >= 1: multiplicity of this storage in state_dict() */
  int32_t reserved;
} b200ssl_ema_block;

/* One launch handles the rows whose dtype == float_dtype (F32, BF16 or F16) and,
 * when do_ints != 0, the integer rows; rows of another float type are skipped
 * (a mixed-precision model takes one launch per float type present). */
B200SSL_API int b200ssl_ema_multi_tensor(const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype,
                             int32_t do_ints, float decay, float one_minus_decay,
                             int32_t mode /*0 update, 1 set*/, void* stream);

/* The same update with at most `max_ctas` CTAs (0 = 4 per SM on every SM, as above).  For an update that runs on a side
 * stream next to other kernels (ema.ModelEMA(overlap=True)): its CTAs are persistent and four of them fill an SM's register
 * file, so a grid of 4 x (SMs - n) leaves n SMs -- the ones a kernel launched ahead of it already holds -- free for the whole
 * update; the SSL head's tensor-core kernels need whole SMs (160-225 KB of shared memory, 55 K registers). */
B200SSL_API int b200ssl_ema_multi_tensor_ctas(const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype,
                                  int32_t do_ints, float decay, float one_minus_decay,
                                  int32_t mode /*0 update, 1 set*/, int32_t max_ctas, void* stream);

/* The update with a set of SMs left alone: `sm_mask8` (8 x uint32, bit i = SM id i) marks SMs whose CTAs exit at once; the
 * other CTAs fetch the table's chunks from the device-side scheduler `sched2` (2 x uint32, zero before the first launch,
 * re-armed by every launch).  For an update that runs next to kernels which need whole SMs (ema.ModelEMA(overlap=True)): the
 * marked SMs are free for them whatever the order in which the launches were placed.  b200ssl_probe_sm_set finds a set in
 * which `n_clusters` thread-block clusters of `cluster` one-CTA-per-SM blocks fit side by side (it ORs their SM ids into
 * `mask8`; `counter` is one zeroed uint32 of scratch). */
B200SSL_API int b200ssl_probe_sm_set(int32_t n_clusters, int32_t cluster, uint32_t* mask8, uint32_t* counter, void* stream);
B200SSL_API int b200ssl_ema_multi_tensor_masked(const b200ssl_ema_block* blocks, int32_t n_blocks, int32_t float_dtype,
                                    int32_t do_ints, float decay, float one_minus_decay, int32_t mode /*0 update, 1 set*/,
                                    const uint32_t* sm_mask8, uint32_t* sched2, void* stream);

/* A one-thread kernel that holds `stream` for `nanoseconds` (<= 1 ms).  ema.ModelEMA(overlap=True) queues it ahead of the
 * capped update on the side stream: the head's first kernel, queued at the same moment on the other stream, then places
 * its CTAs first -- the order that keeps whole SMs free for the rest of the head. */
B200SSL_API int b200ssl_stream_delay(int64_t nanoseconds, void* stream);

/* ------------------------------------------------------------ SURVEY 8(f1) ----
 * Optimizer step + EMA in one multi-tensor pass over the trainable fp32 parameters.
 * Replaces `self.optimizer.step()` (code/fixmatch.py:123; the SGD-nesterov / Adam /
 * AdamW instances built by code/optimizer.py:43-51 with the no-decay group of
 * :13-27) followed by `self.ema_model.update(self.model)` (fixmatch.py:127,
 * ema.py:51-59) for those parameters; buffers and frozen parameters stay with
 * b200ssl_ema_multi_tensor.  Per element, the stock optimizers' single-tensor formulas:
 *   SGD    g += wd*p; buf = first_step ? g : momentum*buf + g; g = nesterov ? g + momentum*buf : buf; p -= lr*g
 *   ADAM   g += wd*p; m += (1-beta1)*(g-m); v = beta2*v + (1-beta2)*g*g; p -= step_size * m / (sqrt(v)/bias2_sqrt + eps)
 *   ADAMW  p *= decay_factor; then ADAM with wd = 0
 * then, when `ema` != NULL, `ema_repeat` times  e = decay*e + one_minus_decay*p  (rounded op by op).
 * The block table is DEVICE resident and 16-byte aligned; the group rows (1..8) are a HOST array read
 * at call time and passed to the kernel by value.  The host fills them every step:
 * step_size = lr/(1-beta1^t), bias2_sqrt = sqrt(1-beta2^t), decay_factor = 1-lr*wd,
 * one_minus_beta* = 1-beta*, all computed in double and rounded once.
 */
enum b200ssl_opt_kind { B200SSL_OPT_SGD = 0, B200SSL_OPT_ADAM = 1, B200SSL_OPT_ADAMW = 2 };

typedef struct b200ssl_opt_block {   /* one chunk (<= 4096 elements is the shipped host policy) of one parameter */
  void* param;
  const void* grad;
  void* state1;        /* SGD: momentum_buffer (NULL when momentum == 0); Adam: exp_avg */
  void* state2;        /* Adam: exp_avg_sq; SGD: NULL */
  void* ema;           /* EMA copy of the parameter, or NULL */
  int32_t count;       /* elements in this chunk */
  int32_t group;       /* row of the group table */
  int32_t ema_repeat;  /* multiplicity of the storage in state_dict() (code/ema.py loops over every entry) */
  int32_t reserved;
  int64_t reserved2;
} b200ssl_opt_block;

typedef struct b200ssl_opt_group {
  float lr, beta1, beta2, eps, weight_decay, step_size, bias2_sqrt, momentum;
  int32_t kind, nesterov, first_step, reserved;
  float one_minus_beta1, one_minus_beta2, decay_factor, pad;
} b200ssl_opt_group;

B200SSL_API int b200ssl_opt_ema_multi_tensor(const b200ssl_opt_block* blocks, int32_t n_blocks, const b200ssl_opt_group* groups,
                                 int32_t n_groups, float decay, float one_minus_decay, void* stream);

/* The same launch with the group rows in DEVICE memory (groups_dev: b200ssl_opt_group[n_groups], 16-byte aligned): the
 * launch has no step-dependent parameter any more, so a CUDA graph can replay it; the host refreshes the rows (learning rate,
 * bias corrections) with one small copy ahead of every replay (fused_step.FusedOptimizerEMA.capture). */
B200SSL_API int b200ssl_opt_ema_multi_tensor_dev(const b200ssl_opt_block* blocks, int32_t n_blocks,
                                     const b200ssl_opt_group* groups_dev, int32_t n_groups, float decay,
                                     float one_minus_decay, void* stream);

/* ------------------------------------------------------------ SURVEY 8e ----
 * Row exchanges of the rank-sharded memory bank over NVLink peer memory.  The
 * reference has no distributed code (one bank per process, code/comatch.py:90-96);
 * these entry points carry the three exchanges a sharded step needs (all-gather of
 * the enqueue blocks, reduce-scatter of the smoothing partial sums, all-gather of
 * the probability blocks) without a communication library on the data path.
 *
 * Arena: one device allocation per rank, exported through CUDA IPC and mapped by
 * every rank of the node.  Layout: b200ssl_peer_control_bytes() of flags/epochs
 * (zero-initialised by b200ssl_peer_alloc), then caller-defined staging regions;
 * the region of one exchange holds [2 parities][world][slot_bytes].
 *   arenas       device array [world] of arena base addresses AS MAPPED IN THIS
 *                PROCESS (entry `rank` = the own arena)
 *   exchange_id  0..7, one per call site; all ranks must issue the same sequence
 *                of calls (SPMD).  Epoch counters live in the arena: launches can
 *                be captured in a CUDA graph and replayed.
 * One launch = push to all peers + publish + wait + copy-out, `out` is ordinary
 * local memory.  A wait gives up after 20 s and counts a timeout instead of
 * hanging the device (the results of that step are then undefined).
 */
B200SSL_API size_t b200ssl_peer_control_bytes(void);
B200SSL_API int b200ssl_peer_alloc(size_t bytes, void** arena, void* ipc_handle_64 /* 64 bytes out */);
B200SSL_API int b200ssl_peer_open(const void* ipc_handle_64, void** arena);
B200SSL_API int b200ssl_peer_close(void* arena);
B200SSL_API int b200ssl_peer_free(void* arena);
B200SSL_API int b200ssl_peer_timeouts(const void* own_arena, uint32_t* count /* host out; synchronises */);

/* Enqueue (code/comatch.py:187-196) of this rank's block [feats_u_w ; feats_x] / [probs_orig ; onehot(targets_x)] into a
 * peer-memory resident bank (b200ssl_bank_shards): global rows [ptr + rank*n, ptr + (rank+1)*n) mod K go to the shard that
 * owns them or, replicated, into every rank's copy.  Waits for every rank's "reads done" flag of this step, advances the
 * device ring pointer by world*n and publishes "my rows are in".  A wide launch for a side stream: nothing else in the
 * step depends on it (b200ssl_comatch_rows_fused is then called with queue_feats = NULL and shards = NULL). */
B200SSL_API int b200ssl_bank_enqueue_peer(const void* feats_u_w, const void* feats_x, const float* probs_orig,
                              const int64_t* targets_x, int64_t n_u, int64_t n_x, int32_t dim, int32_t classes,
                              int32_t dtype, int64_t* ptr_state, const struct b200ssl_bank_shards* shards, void* stream);

/* out[r] = rank r's block [src0 ; src1] (bytes0 + bytes1 bytes, both multiples of 16), r = 0..world-1. */
B200SSL_API int b200ssl_peer_all_gather(const void* src0, size_t bytes0, const void* src1, size_t bytes1, void* out,
                            void* const* arenas, size_t region_offset, size_t slot_bytes, int32_t exchange_id,
                            int32_t rank, int32_t world, void* stream);

/* out[i] = sum over ranks r = 0..world-1 (in that order, fp32 round-to-nearest) of rank r's src[rank*count + i]. */
B200SSL_API int b200ssl_peer_reduce_scatter_f32(const float* src, float* out, int64_t count_per_rank, void* const* arenas,
                                    size_t region_offset, size_t slot_bytes, int32_t exchange_id, int32_t rank,
                                    int32_t world, void* stream);

/* Launch geometry of the tensor-core K3 (host only): out[0] = row tiles per CTA, out[1] = cluster size, out[2] = clusters
 * per group of row tiles.  remote_shards != 0: the plan of a directly addressed rank-sharded bank. */
/* ------------------------------------------------------- f4 (next), data side --
 * The pixel-exact tail of the reference's view transforms (code/dataset.py:24-109) on uint8 HWC images:
 * RandomHorizontalFlip -> RandomCrop(out_size, padding, padding_mode='reflect') -> ToTensor -> Normalize(mean, std), i.e.
 *   out[n, c, y, x] = ((u8 / 255) - mean[c]) / std[c]      (fp32, every operation rounded like ToTensor / Normalize do)
 * of the source pixel that the flip and the crop window select.  images_hwc: [n, height, width, 3] uint8 (device);
 * out_nchw: [n, 3, out_size, out_size] fp32 or bf16 (out_dtype), 16-byte aligned; flip: int32 [n] or NULL; crop_xy:
 * int32 [n, 2] = (left, top) of the window inside the reflect-padded image, or NULL for the centred window; mean3 /
 * std3: HOST arrays of 3 floats.  The random decisions stay with the caller (views.draw_view_params draws them in the
 * order the reference's Compose does). */
B200SSL_API int b200ssl_normalize_views(const uint8_t* images_hwc, void* out_nchw, int32_t n, int32_t height, int32_t width,
                            int32_t out_size, int32_t padding, const int32_t* flip, const int32_t* crop_xy,
                            const float* mean3, const float* std3, int32_t out_dtype, void* stream);

/* Tuning aid: force the row tiles per CTA, the cluster size and the clusters per group of row tiles (0 = planner) and the
 * number of exponentials out of 32 that the tensor-core K3 computes with the FMA-pipe polynomial (-1 = default). */
B200SSL_API void b200ssl_debug_set_k3(int32_t row_tiles_per_cta, int32_t cluster, int32_t clusters_per_row_group, int32_t poly_of_32);

/* SM budget of the SSL head (process-wide; 0 = none, the default; returns the previous value).  With a budget the launch
 * planners of K3 and K6 pick the best geometry of at most `sms` CTAs, as long as it costs no more than 1.5 x the unconstrained
 * one.  Set by ema.ModelEMA(overlap=True) to the SMs its capped update (b200ssl_ema_multi_tensor_ctas) leaves free: a head
 * kernel that took more SMs would hand the surplus to the update's pending CTAs when it retires, and the tensor-core kernels
 * after it -- which need whole SMs -- would wait for the update to finish. */
B200SSL_API int b200ssl_set_head_sm_budget(int32_t sms);

/* A/B aid: 1 = K3 and K6 with fp32 storage run the exact-fp32 FFMA tiles (csrc/bank.cu, csrc/contrast.cu) instead of the
 * tensor-core kernels on bf16 hi + mid operands (csrc/bank_tc.cu, csrc/contrast_tc.cu); 0 = default.  B200SSL_K3_F32_SIMT=1 in
 * the environment sets it at load time. */
B200SSL_API void b200ssl_debug_set_k3_f32_simt(int32_t on);

/* Clusters of `cluster` CTAs of the tensor-core K3 that the device runs at once (driver occupancy query; a table without a device). */
B200SSL_API int b200ssl_debug_max_active_clusters(int32_t cluster);

B200SSL_API int b200ssl_debug_smooth_plan(int64_t rows, int64_t bank_rows, int32_t remote_shards, int32_t* out_mt_cluster_nouter);

#ifdef __cplusplus
}
#endif
#endif /* B200SSL_H_ */
