/*
 * frx.h -- C ABI of the B200-native brand x post scoring + ranking library
 *          (libfrx_b200.so, sm_100a only).
 *
 * The reference (pinskyrobin/FancyRec) is pure Python/PyTorch and has no FFI of its
 * own; its callers bind by Python name.  This header is therefore the
 * boundary a maintainer binds with ctypes from the reference's own modules
 * (INTEGRATION.md shows the stubs).  Every entry point names the reference lines it
 * replaces.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name starts with "host_";
 *   - caller owns every buffer; the library never allocates (query
 *     frx_*_workspace_bytes and pass the scratch in);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and
 *     re-entrant across streams as long as buffers/workspaces are distinct;
 *   - return 0 on success, a negative FRX_E_* code otherwise; frx_last_error()
 *     returns a thread-local message for the last failing call;
 *   - there is NO CPU fallback: with no sm_100 device every compute call fails with
 *     FRX_E_DEVICE.
 *   - ordering everywhere: (score descending, post index ascending) -- the order
 *     Python's stable sorted(..., reverse=True) gives at evaluator.py:109.
 */
#ifndef FRX_H_
#define FRX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRX_ABI_VERSION 1

#define FRX_OK          0
#define FRX_E_ARG      -1   /* bad argument (null pointer, size, alignment)   */
#define FRX_E_DEVICE   -2   /* no sm_100 device / driver entry point missing  */
#define FRX_E_WORKSPACE -3  /* workspace too small                            */
#define FRX_E_CUDA     -4   /* a CUDA runtime/driver call failed              */
#define FRX_E_UNSUPPORTED -5

/* flags of frx_finalize_posts */
#define FRX_VISUAL_NORM 1   /* model.py:207-208  */
#define FRX_TEXT_NORM   2   /* model.py:301-302, :382-383 */
#define FRX_FINAL_NORM  4   /* evaluator.py:27-28 (cal_sim's l2norm of the post operand) */

int         frx_abi_version(void);
const char* frx_last_error(void);
/* 0 when device `ordinal` is an sm_100 part this library can run on, else FRX_E_DEVICE. */
int         frx_device_check(int ordinal);

/* ---------------------------------------------------------------------------------------------
 * A1-A3  post-embedding finalisation, one HBM pass.
 * Replaces: torch.mean(frames, 0) in util/data_provider.py:40,91,132 (mean over ALL frames of a
 * post, rows taken from the feature.bin matrix of util/imgbigfile.py:7-16), model.l2norm
 * (model.py:39-44) per branch, torch.cat((visual, text), 1) (model.py:482-485) and the row l2norm
 * of evaluator.py:14-19,27-28.
 *
 *   visual   [n_rows, dv] fp32 row-major.  With row_ptr == NULL it is one row per post.
 *   row_ptr  [n_posts + 1] int64 CSR offsets: post p owns frame rows row_ptr[p] .. row_ptr[p+1]-1
 *            (or entries of row_idx in that range), mean-pooled.  NULL = no pooling.
 *   row_idx  optional [row_ptr[n_posts]] int32 gather list into `visual` rows (frames of one video
 *            are not guaranteed contiguous, preprocess/get_frameInfo.py:55).  NULL = contiguous.
 *   text     [n_posts, dt] fp32 or NULL (dt = 0).
 *   out_f32  [n_posts, dv + dt] fp32 or NULL.
 *   out_bf16 [n_posts, ld_bf16] bf16 bits or NULL; columns dv+dt .. ld_bf16-1 are zero-filled
 *            (operand layout of frx_score_*: ld_bf16 % 8 == 0).
 * A zero row yields NaN exactly like the reference (no epsilon).
 */
int frx_finalize_posts(const float* visual, const int64_t* row_ptr, const int32_t* row_idx,
                       const float* text, int64_t n_posts, int dv, int dt, int flags,
                       float* out_f32, uint16_t* out_bf16, int64_t ld_bf16, void* stream);
/* Same pass with a bound on the thread blocks the kernel keeps resident per SM (0 = the measured optimum of the shape,
 * what frx_finalize_posts uses).  blocks_per_sm = 1 is the co-residency form: the contraction kernel (frx_score_*)
 * leaves room on every SM for exactly ONE finalise block, so a finalisation launched on a second stream with this bound
 * runs UNDER a contraction instead of queueing behind it or locking it out (fancyrec_b200/pipeline.py). */
int frx_finalize_posts_bounded(const float* visual, const int64_t* row_ptr, const int32_t* row_idx,
                               const float* text, int64_t n_posts, int dv, int dt, int flags,
                               float* out_f32, uint16_t* out_bf16, int64_t ld_bf16, int blocks_per_sm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A4  brand embedding without the [NB, A, D] intermediate.
 * Replaces: BrandAspects.forward in eval mode (model.py:419-428; dropout off, L1Penalty forward =
 * identity) followed by .permute(1,0,2).mean(0) (model.py:594, evaluator.py:93-94):
 *      out[i, :] = (1/A) * sum_a W[ids[i], a] * E[a, :]
 *   w [w_rows, a] fp32 (nn.Embedding table, brand_num + 1 rows), e [a, d] fp32,
 *   brand_ids [nb] int64 or NULL (= 0 .. nb-1), out_f32 [nb, d].
 * With a workspace of frx_brand_embed_workspace_bytes (a % 4 == 0) the product runs on the tensor cores as a 3xTF32
 * GEMM (operands split into tf32 hi + lo parts, K-concatenated, fp32 accumulation: ~1e-6 relative, fp32-grade);
 * with workspace == NULL it runs as an fp32 FMA GEMM on the CUDA cores.
 */
size_t frx_brand_embed_workspace_bytes(int nb, int a, int d);
int frx_brand_embed(const float* w, int64_t w_rows, const float* e, const int64_t* brand_ids,
                    int nb, int a, int d, float* out_f32,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A5 + A6  cosine-score contraction with the per-brand top-k fused into the GEMM epilogue.
 * Replaces: cal_sim's im.mm(s.t()) (evaluator.py:29), the device->host copy of the whole score
 * matrix (evaluator.py:96) and the per-brand sorted() / np.argsort (evaluator.py:108-109,124).
 * tcgen05.mma (bf16 in, fp32 accumulate in TMEM), TMA operand loads; the score matrix is never
 * written unless `dense_out` is given.
 *
 *   brand_bf16 [nb, ld_a], post_bf16 [n_posts, ld_b]: bf16 bits, rows already L2-normalised,
 *            ld_* % 8 == 0, columns d .. ld-1 zero, base pointers 16-byte aligned.
 *   k        1 .. 1024 (lists are padded with score = -inf, index = -1 when n_posts < k).
 *   labels   optional [n_posts] int32 brand label of each post; when given, pos_score[j] receives
 *            the score of post j against its own brand, S[labels[j], j] (0 <= labels[j] < nb), the
 *            only entries the "positive" side of AUC / first-positive-rank needs.  Posts whose
 *            label is outside [0, nb) get NaN.
 *   index_base  global index of post 0 of this shard; output indices are index_base + local and
 *            must fit int32.
 *   topk_scores [nb, k] fp32, topk_index [nb, k] int32, sorted by (score desc, index asc).
 *   dense_out optional [nb, ld_dense] fp32: the full score tile (debug / AUC path), same
 *            accumulators as the fused path, bit for bit.
 */
size_t frx_score_topk_workspace_bytes(int nb, int64_t n_posts, int d, int k);
int frx_score_topk(const uint16_t* brand_bf16, int64_t ld_a, const uint16_t* post_bf16, int64_t ld_b,
                   int nb, int64_t n_posts, int d, int k,
                   const int32_t* labels, int64_t index_base,
                   float* topk_scores, int32_t* topk_index, float* pos_score,
                   float* dense_out, int64_t ld_dense,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Dense score tile only (no top-k): out[b, j] = brand_b . post_j, for the score-tolerance tests and
 * the AUC row sweep.  Same kernel mainloop as frx_score_topk. */
int frx_score_dense(const uint16_t* brand_bf16, int64_t ld_a, const uint16_t* post_bf16, int64_t ld_b,
                    int nb, int64_t n_posts, int d, float* dense_out, int64_t ld_dense, void* stream);

/* Count pass: for every brand b, the number of posts that precede (thr_score[b], thr_index[b]) in
 * the stated order: #{ j : S[b,j] > thr_score[b]  or  (S[b,j] == thr_score[b] and index_base + j <
 * thr_index[b]) }.  With the threshold set to a brand's best positive this IS rank_of_first_pos
 * (evaluator.py:116) without materialising S.  count_out [nb] int64 is ACCUMULATED into (caller
 * zeroes it; shards add up).  Rows with thr_index[b] < 0 are skipped, and so is every 128-brand tile made only of such
 * rows (the launch returns immediately when no row has a threshold). */
int frx_score_count(const uint16_t* brand_bf16, int64_t ld_a, const uint16_t* post_bf16, int64_t ld_b,
                    int nb, int64_t n_posts, int d, int64_t index_base,
                    const float* thr_score, const int32_t* thr_index,
                    unsigned long long* count_out, void* stream);

/* Masked soft-max weighted pool (MultiHeadSelfAttention.forward, model.py:105-114, without its per-sample loop):
 *   out[b, :] = (1 / t_max) * sum_{t < lengths[b]} softmax(logits[b, :lengths[b]])[t] * x[b, t, :]
 * x [batch, t_max, d] fp32, logits [batch, t_max] fp32 (the head-averaged attention scores), lengths [batch] int64
 * (clamped to 0..t_max; a zero length gives a zero row), out [batch, d] fp32.  One HBM pass over the valid steps. */
int frx_softmax_pool(const float* x, const float* logits, const int64_t* lengths, int batch, int t_max, int d,
                     float* out, void* stream);

/* 3xTF32 operands (fp32-grade scores on the tensor cores, the mode that meets the 1e-5 score tolerance):
 * x = hi + lo with hi, lo exactly representable in tf32; the K-concatenated operands
 *   brand side (side 0): [hi | lo | hi]      post side (side 1): [hi | hi | lo]        (each [rows, 3 * cols] fp32)
 * make ONE tf32 contraction over K = 3 * cols equal to a_hi.b_hi + a_lo.b_hi + a_hi.b_lo = a.b up to O(2^-22).
 * Feed the results to the frx_score_*_tf32 entry points with d = 3 * cols: same kernels, same fused top-k,
 * ~6x the bf16 time.  Replaces nothing in the reference -- it is how cal_sim's fp32 mm (evaluator.py:29) is
 * matched to 1e-5 without leaving the tensor cores.  cols % 4 == 0, ld_x % 4 == 0, 16-byte aligned pointers. */
int frx_split_tf32x3(const float* x, int64_t rows, int cols, int64_t ld_x, int side, float* out, void* stream);

/* tf32 variants: identical contracts, operands are fp32 [rows, ld] (rows L2-normalised, ld % 4 == 0, 16-byte aligned base),
 * read by the tensor cores as tf32 (10-bit mantissa, tcgen05.mma.kind::tf32) with fp32 accumulation: scores within 2e-4 of
 * fp32 cal_sim at D >= 1024 on the cosine scale (observed 1.1e-4), at half the bf16 tensor throughput and twice the operand bytes. */
int frx_score_topk_tf32(const float* brand_f32, int64_t ld_a, const float* post_f32, int64_t ld_b,
                        int nb, int64_t n_posts, int d, int k,
                        const int32_t* labels, int64_t index_base,
                        float* topk_scores, int32_t* topk_index, float* pos_score,
                        float* dense_out, int64_t ld_dense,
                        void* workspace, size_t workspace_bytes, void* stream);
int frx_score_dense_tf32(const float* brand_f32, int64_t ld_a, const float* post_f32, int64_t ld_b,
                         int nb, int64_t n_posts, int d, float* dense_out, int64_t ld_dense, void* stream);
int frx_score_count_tf32(const float* brand_f32, int64_t ld_a, const float* post_f32, int64_t ld_b,
                         int nb, int64_t n_posts, int d, int64_t index_base,
                         const float* thr_score, const int32_t* thr_index,
                         unsigned long long* count_out, void* stream);

/* Merge G candidate lists per brand (the multi-GPU exchange step after the all-gather):
 * in_scores / in_index [g, nb, k_in] -> out [nb, k_out], same order.  Every input list must be sorted under the total
 * order (score desc, index asc) with its padding entries (index < 0) at the tail -- i.e. be a list frx_score_topk or this
 * function produced; a post appears in at most one list.  Requires g * k_in <= 16384. */
int frx_topk_merge(const float* in_scores, const int32_t* in_index, int g, int nb, int k_in,
                   float* out_scores, int32_t* out_index, int k_out, void* stream);
/* Same, reading shard r's [nb, k_in] lists at in_scores + r * shard_stride / in_index + r * shard_stride (elements):
 * the lists are merged in place out of the packed all-gather buffer of the multi-GPU exchange. */
int frx_topk_merge_strided(const float* in_scores, const int32_t* in_index, int g, int nb, int k_in,
                           int64_t shard_stride, float* out_scores, int32_t* out_index, int k_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A6-A9  rank statistics (integers; the float64 metric values are functions of these).
 *
 * frx_label_stats: n_pos[b] = #{j : labels[j] == b}; best_score/best_index[b] = the best positive
 * of brand b under the stated order (index = index_base + j; -1 when n_pos[b] == 0).
 * Replaces the pos/neg split of evaluator.py:111-112 and the argmax side of :116.
 */
int frx_label_stats(const int32_t* labels, const float* pos_score, int64_t n_posts, int nb,
                    int64_t index_base, int32_t* n_pos, float* best_score, int32_t* best_index,
                    void* workspace_nb_u64, void* stream);

/* frx_rank_from_topk: one warp per brand walks its sorted top-k list.
 *   hit_mask[b]  bit r (r < 64) set when the post at rank r is labelled b -- the relevance vector
 *                ndcg_at_k reads (evaluator.py:119-120, util/ndcg.py:37);
 *   first_rank[b] rank of the first positive inside the list, or -1 when the list holds none
 *                (then frx_score_count supplies it).
 * labels are indexed with (topk_index - index_base); entries outside [0, n_posts) count as misses. */
int frx_rank_from_topk(const int32_t* topk_index, int nb, int k, const int32_t* labels, int64_t n_posts,
                       int64_t index_base, unsigned long long* hit_mask, int32_t* first_rank, void* stream);

/* frx_reduce_shard_stats (multi-GPU): combine the per-shard label statistics after their all-gather.
 * Shard r's n_pos / best_score / best_index arrays ([nb] each) start r * stride_words 32-bit words after shard 0's
 * (so they can be read in place from one packed all-gather buffer).  n_pos = sum over shards; the best positive is
 * the maximum over shards under (score desc, index asc); (-inf, -1) when no shard has one. */
int frx_reduce_shard_stats(const int32_t* n_pos_g, const float* best_score_g, const int32_t* best_index_g, int g, int nb,
                           int64_t stride_words, int32_t* n_pos, float* best_score, int32_t* best_index, void* stream);

/* frx_missing_thresholds: thr_index[b] = best_index[b] for the brands whose first positive is NOT in the top-k list
 * (first_rank[b] < 0 and n_pos[b] > 0), else -1.  With thr_score = best_score this is the input of frx_score_count, which
 * skips every 128-brand tile that has no such row -- so the count pass can be enqueued unconditionally (no host
 * round trip to decide): it costs nothing when every first positive was found in the list (evaluator.py:116).
 * frx_pack_rank_stats: the per-brand integers an evaluation returns to the host as ONE [5|6, nb] int64 block
 * (n_pos, first rank in list, count before first positive, validity of that count, hit mask, AUC numerator if given). */
int frx_missing_thresholds(const int32_t* n_pos, const int32_t* first_in_list, const int32_t* best_index, int nb,
                           int32_t* thr_index, void* stream);
int frx_pack_rank_stats(const int32_t* n_pos, const int32_t* first_in_list, const unsigned long long* before_first,
                        const unsigned long long* hit_mask, const unsigned long long* auc_num, int nb, int all_valid,
                        long long* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A4, training path: BrandAspects.forward + mean over aspects WITH dropout(0.5) on the [B, A, D] products
 * (model.py:419-428, :594) and its backward incl. L1Penalty.backward (model.py:395-401), never materialising [B, A, D]:
 *      out[b, d]     = (2 / A) * sum_a w_rows[b, a] * E[a, d] * m(b, a, d)
 *      d_w_rows[b,a] = (2 / A) * sum_d g[b, d] * E[a, d] * m(b, a, d) + 1e-4 * sign(w_rows[b, a])
 *      d_e[a, d]     = (2 / A) * sum_b g[b, d] * w_rows[b, a] * m(b, a, d)
 * w_rows [b, ld_w] = the embedding rows of the batch (the caller gathers them; autograd scatters d_w_rows back).
 * m(b, a, d) in {0, 1} is a counter-based hash of (seed, b, a, d), identical in all three kernels for one seed;
 * frx_brand_dropout_mask writes it out as uint8 [b, a, d] (tests only). */
int frx_brand_train_fwd(const float* w_rows, int64_t ld_w, const float* e, int b, int a, int d, uint64_t seed, float* out,
                        void* stream);
int frx_brand_train_bwd(const float* grad_out, const float* w_rows, int64_t ld_w, const float* e, int b, int a, int d,
                        uint64_t seed, float* d_w_rows, float* d_e, void* stream);
int frx_brand_dropout_mask(int b, int a, int d, uint64_t seed, uint8_t* mask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Encoder-side Linear layers (model.py:59-83 MFC, :463-491 PrjHeadFusionEncoder):
 *      out[m, n] = act( (sum_k x[m, k] * w[n, k]) * col_scale[n] + col_shift[n] )
 * i.e. nn.Linear (w = weight [n, k], col_shift = bias), with an eval-mode BatchNorm1d folded into col_scale / col_shift
 * and an optional ReLU, as ONE fp32-grade tensor-core GEMM (3xTF32 split inside the kernel, tcgen05 kind::tf32, fp32
 * accumulation) -- no cuBLAS call, no intermediate activation tensor.  x [m, ld_x], w [n, ld_w] fp32 row-major, 16-byte
 * aligned, pitches multiples of 4 floats; col_scale / col_shift [n] or NULL; workspace only for small outputs whose K range
 * is split over idle SMs (frx_linear_workspace_bytes). */
size_t frx_linear_workspace_bytes(int m, int n, int k);
int frx_linear(const float* x, int64_t ld_x, const float* w, int64_t ld_w, const float* col_scale, const float* col_shift,
               int relu, int m, int n, int k, float* out, int64_t ld_out, void* workspace, size_t workspace_bytes,
               void* stream);
/* The same kernel as a plain product with either operand stored transposed -- the gradient products of the losses
 * (loss.py:87-143: dPost = dS . brand, dBrand = dS^T . post) read dS and the embeddings in place:
 *      out[m, n] = alpha * sum_k A(m, k) * B(n, k)
 * A is a[m, ld_a] (a_transposed = 0) or a[k, ld_a] with A(m, k) = a[k * ld_a + m] (a_transposed = 1); B likewise with n
 * rows.  Transposed operands whose row count is a multiple of 32 are fed to the tensor core as MN-major tiles (no
 * transposition anywhere); others are transposed in shared memory.  Workspace: frx_linear_workspace_bytes(m, n, k). */
int frx_matmul3x(const float* a, int64_t ld_a, int a_transposed, const float* b, int64_t ld_b, int b_transposed, int m,
                 int n, int k, float alpha, float* out, int64_t ld_out, void* workspace, size_t workspace_bytes,
                 void* stream);

/* frx_metric_scores (A10; util/metric.py:6-123): the reference's rank-metric scorers over a BATCH of sorted label lists,
 * one warp per list.  labels int32 [n_lists, ld] (graded relevance, list i occupies labels[i*ld .. i*ld + lengths[i]);
 * lengths NULL = every list has max_len entries).  kind: 0 P@k (PrecisionScorer), 1 AP@k (APScorer), 2 RR (RRScorer),
 * 3 NDCG@k (NDCGScorer), 4 DCG@k (DCGScorer); k = 0 means "whole list" exactly as MetricScorer.getLength does.
 * log2_table [max_len + 2] float64 with log2_table[i] = math.log(i, 2) for i >= 1, computed by the caller on the host
 * (CPython evaluates log(i) / log(2); the table makes the device result bit-identical).  out float64 [n_lists]:
 * bit for bit what `getScorer(name).score(list)` returns; where the reference raises (NDCG of an all-zero list:
 * ZeroDivisionError; empty list) the entry is NaN and the Python mirror raises the same exception. */
int frx_metric_scores(const int32_t* labels, int64_t ld, const int32_t* lengths, int n_lists, int max_len, int kind, int k,
                      const double* log2_table, double* out, void* stream);

/* frx_auc_rows: exact AUC numerators from dense score rows (evaluator.py:111-113):
 *   auc_num[row0 + r] += sum over positives e of brand (row0+r) of #{negatives el : e > el}
 *   before_first[row0 + r] += #{j : (S[r,j], j) precedes the brand's best positive}
 * scores [n_rows, ld] fp32 is the dense tile of brands row0 .. row0+n_rows-1 (frx_score_dense);
 * pos_sorted holds every brand's positive scores grouped by brand in ascending order, seg_ptr [nb+1]
 * the group offsets (frx_group_positives builds both).  Outputs are accumulated (caller zeroes). */
int frx_group_positives(const int32_t* labels, const float* pos_score, int64_t n_posts, int nb,
                        const int32_t* n_pos, int64_t* seg_ptr, float* pos_sorted,
                        void* workspace_nb_i64, void* stream);
int frx_auc_rows(const float* scores, int64_t ld, int row0, int n_rows, int64_t n_posts,
                 const int32_t* labels, const int64_t* seg_ptr, const float* pos_sorted,
                 const float* best_score, const int32_t* best_index, int64_t index_base,
                 unsigned long long* auc_num, unsigned long long* before_first, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A12  TripletLoss on the in-batch similarity tile (loss.py:87-143), forward + backward fused.
 *   S[i,j] = post_i . brand_j (fp32); rank weights (loss.py:96-105); hinge with same-brand mask
 *   (loss.py:107-129); weights broadcast along the column index (loss.py:131-132); sum | mean.
 *   Outputs: loss[1] fp32, d_brand [b, d], d_post [b, d] (gradients of the loss itself; scale by the
 *   upstream gradient on the host side).  mean_style: 0 = 'sum', 1 = 'mean'.
 */
size_t frx_triplet_workspace_bytes(int b, int d);
int frx_triplet_fwd_bwd(const int64_t* brand_ids, const float* brand, const float* post, int b, int d,
                        float margin, int mean_style, float* loss, float* d_brand, float* d_post,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Opt-in hardest-negative ("VSE++", max of violations) hinge on the same tile -- what loss.py's ignored `max_violation`
 * flag stands for; NOT what TripletLoss.forward computes:
 *   loss = sum_i max_j [m + S_ij - S_ii]_+  +  sum_j max_i [m + S_ij - S_jj]_+ , same-brand entries excluded
 *   (loss.py:116-119), no rank weights; mean_style 1 divides by b.  Same outputs and workspace as frx_triplet_fwd_bwd. */
int frx_vsepp_fwd_bwd(const int64_t* brand_ids, const float* brand, const float* post, int b, int d, float margin,
                      int mean_style, float* loss, float* d_brand, float* d_post, void* workspace, size_t workspace_bytes,
                      void* stream);

/* A13  ContrastiveLoss (loss_ctrs.py:179-214), forward + backward.
 *   keys [n_keys, d]: the queue AFTER enqueue (loss_ctrs.py:138-147) or NULL for the
 *   no_queue / no_intra variants (keys = normalised posts); mask_col0 = column masked for row 0
 *   (queue pointer after the move, loss_ctrs.py:149-159; row i masks column mask_col0 + i).
 *   post_norm_out [b, d] receives F.normalize(post) (what the caller enqueues).
 */
size_t frx_contrastive_workspace_bytes(int b, int d, int n_keys);
int frx_contrastive_fwd_bwd(const float* brand, const float* post, int b, int d,
                            const float* keys, int n_keys, int mask_col0, int no_intra,
                            float temperature, float negative_weight, int mean_style,
                            float* loss, float* d_brand, float* d_post,
                            void* workspace, size_t workspace_bytes, void* stream);
/* F.normalize(post) rows (eps 1e-12), what ContrastiveLoss enqueues (loss_ctrs.py:195,200). */
int frx_normalize_rows(const float* x, int rows, int d, float* out, void* stream);

/* A14  CrossCLR_onlyIntraModality (loss_ctrs.py:52-117), forward + backward.
 *   Row and column rank weights from the raw tile post.brand^T (loss_ctrs.py:62-77); F.normalize both; four
 *   B x B Gram matrices / T; the intra-modality logits * negative_weight with their diagonal entry multiplied to 0
 *   (it stays in the soft-max as a logit of 0, loss_ctrs.py:97-104); loss = (sum|mean_b + sum|mean_p) / 2.
 */
size_t frx_crossclr_workspace_bytes(int b, int d);
int frx_crossclr_fwd_bwd(const float* brand, const float* post, int b, int d,
                         float temperature, float negative_weight, int mean_style,
                         float* loss, float* d_brand, float* d_post,
                         void* workspace, size_t workspace_bytes, void* stream);

/* A14  LabLoss (loss.py:55-63): (sum exp(cos(brand_i, brand_j), diagonal filled with 0) - B) / B, forward + backward
 *   (d_brand may be NULL).  Rows are normalised without an epsilon, as loss.l2norm does (loss.py:20-24). */
size_t frx_lab_workspace_bytes(int b, int d);
int frx_lab_fwd_bwd(const float* brand, int b, int d, float* loss, float* d_brand,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement hook (bench.py roofline leg): when enabled, every MAIN launch of the fused score + top-k kernel
 * (not the sample pass, not dense / count launches) is bracketed by a pair of CUDA events on its own stream.  frx_probe_read synchronises
 * the recorded events and returns up to `max` per-launch durations in milliseconds (oldest first)
 * and clears the log.  Off by default; at most 4096 launches are logged.
 */
int frx_probe_enable(int on);

/* Kernel variant of every frx_score_* entry point: 0 = one CTA per SM (default), 1 = CTA pairs (tcgen05 cta_group::2:
 * the two SMs of a TPC share one 256 x 256 MMA tile, each staging its 128 brand rows and half of the post tile),
 * -1 = back to the FRX_PAIR environment variable.  Results are identical bit for bit (same accumulation order per
 * output element).  Returns the previous setting.  Process-wide; not for concurrent use with running launches. */
int frx_set_cta_pairs(int on);
/* Third kernel variant of frx_score_*: clusters of `cta_count` (2, 4 or 8) CTAs on consecutive 128-brand tiles share every
 * post tile through TMA multicast (each CTA fetches 1 / cta_count of it for all), bf16 operands.  0 or 1 = off (default), -1 =
 * back to the FRX_CLUSTER environment variable.  Returns the previous setting.  Results are bit-identical in every variant. */
int frx_set_cluster(int cta_count);
int frx_probe_read(float* host_ms_out, int max);

#ifdef __cplusplus
}
#endif
#endif /* FRX_H_ */
