/*
 * rz_b200.h -- C ABI of librz_b200.so: RadZero's VL-CABS similarity path on B200 (sm_100a).
 *
 * The reference (deepnoid-ai/RadZero) has no FFI: the path sits behind plain Python
 * callables.  This header is the boundary the new build introduces
 * below them; each entry point names the reference code it replaces.  radzero_b200's
 * Python mirror of the reference surface (losses.py / modeling.py / inference.py) binds
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors are dense row-major with the strides stated per function;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions enqueue work and return immediately: 0 (RZ_OK) or a negative RZ_ERR_*;
 *     nothing is allocated, no global state is kept besides a launch counter;
 *   - re-entrant; callers own every buffer (workspace sizes come from rz_*_workspace_bytes).
 *   - hidden size D is fixed at 768 (RadZeroLoss.hidden_dim, radzero.yaml:40).
 */
#ifndef RZ_B200_H_
#define RZ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RZ_OK 0
#define RZ_ERR_INVALID (-1)     /* bad shape / null pointer / unsupported size */
#define RZ_ERR_CUDA (-2)        /* a CUDA runtime call failed; see rz_last_cuda_error() */
#define RZ_ERR_UNSUPPORTED (-3) /* valid request outside what this build implements */
#define RZ_ERR_ALIGNMENT (-4)   /* pointer or stride not aligned as documented */

/* element types of the *input* tensors (tokens / text); math is fp32, MMA operands fp16 */
#define RZ_F32 0
#define RZ_BF16 1
#define RZ_F16 2

#define RZ_LN_EPS 1e-5f  /* nn.LayerNorm default, exp/cxr_pt/model/losses.py:51 */
#define RZ_L2_EPS 1e-12f /* F.normalize default, losses.py:212-213 */

/* ---- library ---------------------------------------------------------------------- */
int rz_version(void);                   /* 1000*major + minor */
const char* rz_strerror(int code);
const char* rz_last_cuda_error(void);   /* text of the last CUDA error seen, "" if none */
long long rz_launch_count(void);        /* kernels launched by this library so far */
int rz_device_sm_count(void);

/* ---- K1+K2: LayerNorm + L2 normalisation of rows -------------------------------------
 * Replaces nn.LayerNorm on text / vision tokens (losses.py:90-91, 163-164; the SAME
 * gamma/beta for both, losses.py:51) followed by F.normalize (losses.py:212-213).
 *   x        [rows, 768] of `dtype`, contiguous
 *   gamma/beta fp32 [768], or both NULL for no LayerNorm (use_layer_norm=False)
 *   out_f16  optional [groups, rows_per_group_padded, 768] fp16: row r of group g lands at
 *            (g*rows_per_group_padded + r); padding rows are zero-filled.  With
 *            rows_per_group = rows and rows_per_group_padded = rows this is plain [rows,768].
 *   out_f32  optional [rows, 768] fp32 (unpadded)
 *   stats    optional [rows, 3] fp32 = (mean, rstd, 1/max(|LN(x)|, eps)) kept for backward
 *   l2       0 = LayerNorm only (sim_op "dot"), 1 = LayerNorm then L2 (sim_op "cos")
 */
int rz_prep_rows(const void* x, int dtype, const float* gamma, const float* beta,
                 long long rows, int rows_per_group, int rows_per_group_padded,
                 void* out_f16, float* out_f32, float* stats, int l2, void* stream);

/* ---- K3-K6: fused similarity GEMM + softmax pooling + pooled logit (tcgen05 / TMA) ------
 * Replaces SimilarityLogit.forward (exp/cxr_pt/model/losses.py:187-240: bmm :219-221,
 * softmax :222, matmul :224, normalize + batched dot :226-233) and the CLS drop / transpose
 * of CxrAlignModel.compute_logits (exp/cxr_pt/model/modeling.py:311-328).  The (B,N,L)
 * probability tensor and the (B,N,768) expanded query of the reference are never formed.
 *   k_f16   [n_images, tokens_padded, 768] fp16 rows from rz_prep_rows (padding rows zero);
 *           tokens_padded is a multiple of 64
 *   q_f16   [n_text, 768] fp16 rows from rz_prep_rows
 *   scale   1/tau for sim_op "cos" (losses.py:221), 1/sqrt(768) for "dot" (:215);
 *   log_tau_scale optional DEVICE scalar: when given, scale = exp(-*log_tau_scale) is computed
 *           on the device from the learnable log-temperature (losses.py:54-62, 177-181), so no
 *           host synchronisation is needed to read the parameter
 *   q_inv_norm optional fp32 [n_text]: Z is multiplied by it (sim_op "dot": 1/|q|, the
 *           F.normalize(query) of losses.py:226; NULL for "cos" where |q| = 1)
 *   scores  optional fp32: scores[b*stride_image + n*stride_text + (l - drop_cls)] =
 *           scale * <k_bl, q_n> for tokens l >= drop_cls (drop_cls = 1 removes CLS as
 *           modeling.py:316-317 does); CLS still takes part in the softmax
 *   z       optional fp32: z[n*stride_text + b*stride_image] = f(z_scale * <q_n, o_bn/|o_bn|>)
 *           with f = identity (z_sigmoid = 0) or the logistic sigmoid (z_sigmoid = 1);
 *           log_tau_z (optional DEVICE scalar) overrides z_scale with exp(-*log_tau_z).
 *           t2i_logits is stride_text = B, stride_image = 1, z_scale = 1; the `logits` of
 *           compute_logits (modeling.py:324-328) is the transpose with z_scale = 1/tau, and
 *           similarity_prob the same with z_sigmoid = 1.
 *   lse / onorm optional fp32 [n_images, n_text]: log-sum-exp of the scores over tokens and
 *           |o| of the softmax-weighted token mean -- kept for the backward pass
 *   pooled_f16 optional fp16 [n_images, n_text, 768]: o_bn, kept for the backward pass
 */
int rz_sim_fwd(const void* k_f16, int n_images, int tokens, int tokens_padded,
               const void* q_f16, int n_text, float scale, const float* log_tau_scale,
               const float* q_inv_norm, float* scores, long long scores_stride_image,
               long long scores_stride_text, int drop_cls, float* z, long long z_stride_text,
               long long z_stride_image, float z_scale, const float* log_tau_z, int z_sigmoid,
               float* lse, float* onorm, void* pooled_f16, void* stream);

/* Same computation for small prompt sets (n_text <= 16: zero-shot classification, grounding,
 * model_inference), reading the RAW vision tokens exactly once: TMA streams the raw rows into a
 * shared-memory ring, converter warps apply LayerNorm (losses.py:90-91) + L2 normalisation (:213)
 * in registers and write the fp16 operand tile, tcgen05 computes S^T = q k^T and the pooled sums.
 * No fp16 copy of the tokens is written.  HBM-bound; this is the kernel behind compute_logits /
 * model_inference.  The (image, 32-token tile) sequence is cut into equal contiguous ranges, one
 * per SM, so any batch size (including ONE image) uses the whole GPU; images split across ranges
 * are finished by a small merge kernel through `workspace`.
 *   tokens_raw [n_images, tokens, 768] of `dtype` (RZ_F32 / RZ_BF16 / RZ_F16), contiguous
 *   gamma/beta fp32 [768] or both NULL; l2 as in rz_prep_rows
 *   q_f16      [n_text, 768] fp16 rows from rz_prep_rows (the prompts are tiny: 14 x 768)
 *   workspace  rz_sim_fwd_tokens_workspace_bytes(...) bytes, 16-byte aligned
 * Returns RZ_ERR_UNSUPPORTED for n_text > 16 (use rz_prep_rows + rz_sim_fwd).
  *   text_f32  optional fp32 [n_text, 768]: the prompt embeddings BEFORE LayerNorm + L2 (text_features_wo_l2_norm,
 *          losses.py:139-142).  When given, q_f16 may be NULL: every CTA normalises the rows in its prologue
 *          (compute_text_features' LayerNorm + F.normalize, losses.py:163-164, 212) with the same gamma / beta,
 *          which saves the rz_prep_rows launch of the prompts.  Not with q_inv_norm (sim_op "dot").
 */
size_t rz_sim_fwd_tokens_workspace_bytes(int n_images, int n_text);
int rz_sim_fwd_tokens(const void* tokens_raw, int dtype, const float* gamma, const float* beta,
                      int l2, int n_images, int tokens, const void* q_f16, int n_text,
                      float scale, const float* log_tau_scale, const float* q_inv_norm,
                      float* scores, long long scores_stride_image, long long scores_stride_text,
                      int drop_cls, float* z, long long z_stride_text, long long z_stride_image,
                      float z_scale, const float* log_tau_z, int z_sigmoid, const float* text_f32,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Same computation for LARGE prompt sets (open-vocabulary sweeps, the contrastive step), as
 * two full-rate tcgen05 GEMM passes: pass S computes the scores exactly once, one thread per
 * prompt row keeping a lazy running maximum, and writes the similarity map plus UNNORMALISED
 * probabilities P~ = exp(s - mref) as fp16; pass PK computes o = (P~ k) / lsum with |o| and
 * <q,o> in the epilogue.  Arguments as rz_sim_fwd, plus
 *   p_f16  optional fp16 [n_images, n_text, tokens_padded]: P~, kept for rz_sim_bwd (NULL: the
 *          workspace holds it)
 *   mref / lsum optional fp32 [n_images, n_text]: reference maximum and sum_l exp(s_l - mref)
 *          of every (image, prompt) row (lse = mref + log lsum), kept for rz_sim_bwd
 * tokens_padded must be a multiple of 128; want_pool = 0 with z = onorm = pooled = NULL stops
 * after the first pass (scores and/or lse only).  workspace: 256-byte aligned,
 * rz_sim_fwd_large_workspace_bytes(...) bytes.
 */
size_t rz_sim_fwd_large_workspace_bytes(int n_images, int n_text, int tokens_padded);
int rz_sim_fwd_large(const void* k_f16, int n_images, int tokens, int tokens_padded,
                     const void* q_f16, int n_text, float scale, const float* log_tau_scale,
                     const float* q_inv_norm, float* scores, long long scores_stride_image,
                     long long scores_stride_text, int drop_cls, float* z, long long z_stride_text,
                     long long z_stride_image, float z_scale, const float* log_tau_z, int z_sigmoid,
                     float* lse, float* onorm, void* pooled_f16, void* p_f16, float* mref,
                     float* lsum, int want_pool, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---- backward of K3-K6 -------------------------------------------------------------------
 * Replaces the autograd of SimilarityLogit.forward (losses.py:187-240: the backward of bmm,
 * softmax, matmul, normalize and the batched dot) in closed form: three tcgen05 GEMM passes
 * (recompute S and T = o.k^T with two accumulators -> fp16 coefficient matrices W1, W2;
 * dq = sum_b W1_b k_b; dk_b = W1_b^T q + W2_b^T o_b).  Inputs are what rz_sim_fwd produced.
 *   k_f16 [n_images, tokens_padded, 768], tokens_padded a multiple of 128
 *   z, dz  fp32 [n_text, ldz]: pooled logits and dL/dZ for the local images (columns)
 *   lse, onorm [n_images, n_text]; pooled_f16 [n_images, n_text, 768]
 *   p_f16, mref, lsum  optional: the unnormalised probabilities and row statistics kept by
 *          rz_sim_fwd_large.  When given, the scores are not recomputed (p = P~/lsum,
 *          s = mref + ln P~) and the coefficient pass is ONE GEMM (T = o.k^T) instead of two;
 *          lse may then be NULL.
 *   dq    fp32 [n_text, 768]  = dL/dq (normalised sentence embeddings), overwritten
 *   dk    fp32 [n_images, tokens_padded, 768] = dL/dk (normalised tokens), overwritten
 *   dlog_tau fp32 [1] = dL/dlog(tau_attn) through the softmax scores
 *   inv_tau / log_tau as in rz_sim_fwd (log_tau: optional device scalar)
 *   q_inv_norm  optional fp32 [n_text]: sim_op "dot" (losses.py:214-215, 226) -- q_f16 / k_f16 are the
 *          rows WITHOUT L2 normalisation, inv_tau = 1/sqrt(768), q_inv_norm = 1/|q_n|.  dq then lacks
 *          the radial term -(sum_b dZ Z) q_n / |q_n|^2, which the caller adds (radzero_b200/training.py);
 *          dlog_tau is meaningless in this mode.
 * workspace: rz_sim_bwd_workspace_bytes(...) bytes, 256-byte aligned.
 */
size_t rz_sim_bwd_workspace_bytes(int n_images, int n_text, int tokens_padded);
int rz_sim_bwd(const void* k_f16, int n_images, int tokens, int tokens_padded, const void* q_f16,
               int n_text, float inv_tau, const float* log_tau, const float* z, const float* dz,
               long long ldz, const float* lse, const float* onorm, const void* pooled_f16,
               const void* p_f16, const float* mref, const float* lsum, const float* q_inv_norm,
               float* dq, float* dk, float* dlog_tau, void* workspace, size_t workspace_bytes,
               void* stream);

/* ---- backward of K1+K2 --------------------------------------------------------------------
 * dL/dx and dL/dgamma, dL/dbeta from dL/d(normalised rows) (autograd of nn.LayerNorm +
 * F.normalize, losses.py:90-91, 163-164, 212-213).  Row statistics are recomputed from x.
 *   dnorm  fp32, padded group layout of rz_prep_rows' out_f16 ([groups, rows_per_group_padded, 768])
 *   dx     [rows, 768]: fp32, or -- dx_native = 1 and a 16-bit input dtype -- the input's own type
 *   partials fp32 [rz_prep_rows_bwd_blocks(rows), 2, 768] scratch (needed when gamma != NULL)
 *   dgamma/dbeta fp32 [768]: overwritten, or accumulated into when accumulate = 1 (tokens
 *   and text share one LayerNorm, losses.py:51); grad_scale multiplies the parameter grads.
 */
int rz_prep_rows_bwd_blocks(long long rows);
int rz_prep_rows_bwd(const void* x, int dtype, const float* gamma, const float* beta,
                     long long rows, int rows_per_group, int rows_per_group_padded,
                     const float* dnorm, int l2, void* dx, int dx_native, float* partials,
                     float* dgamma, float* dbeta, int accumulate, float grad_scale, void* stream);

/* ---- K8+K9: bilinear upsample of patch-grid similarity maps ----------------------------
 * Replaces F.interpolate(mode="bilinear", align_corners=False) in
 * interpolate_similarity_scores (exp/cxr_pt/inference/segmentation_utils.py:36-122) and
 * get_grounding_point (exp/cxr_pt/inference/grounding_utils.py:166-261), fused with the
 * consumers torch.sigmoid (segmentation_utils.py:225), `> t` (:258) and the global argmax
 * (grounding_utils.py:254-259).  One launch handles `maps` maps (the reference does one
 * F.interpolate per (image, prompt)).
 *   scores   [maps] grids of grid x grid fp32, map m at scores + m*map_stride (elements)
 *   The grid is resized to (interp_h, interp_w) and pasted with its top-left corner at
 *   (off_y, off_x) of the (out_h, out_w) canvas (offsets may be negative = crop); canvas
 *   pixels outside the pasted area get `fill` (the reference's -999).  The four image
 *   processor branches are parameter choices of this one kernel (see inference.py).
 *   mode     RZ_UP_RAW      out fp32 [maps,out_h,out_w] = interpolated score
 *            RZ_UP_SIGMOID  out fp32 = sigmoid(score)
 *            RZ_UP_MASK     out uint8 = sigmoid(score) > threshold
 *            RZ_UP_ARGMAX   out int64 [maps,2] = (x, y) of the first global maximum; no map
 *            RZ_UP_MASK_BITS out uint32 [maps,out_h,ceil(out_w/32)]: the same mask, one bit per pixel (bit x%32 of
 *                           word x/32, padding bits zero) -- 131 072 masks of 518 x 518 (BASELINE config 5 at
 *                           pixel level) are 4.6 GB instead of 35 GB
 *                           is written
 */
#define RZ_UP_RAW 0
#define RZ_UP_SIGMOID 1
#define RZ_UP_MASK 2
#define RZ_UP_ARGMAX 3
#define RZ_UP_MASK_BITS 4
int rz_upsample_maps(const float* scores, long long map_stride, int maps, int grid,
                     int out_h, int out_w, int interp_h, int interp_w, int off_y, int off_x,
                     float fill, int mode, float threshold, void* out, void* stream);

/* ---- fused consumer of the similarity map: threshold statistics ---------------------------
 * The reference evaluates segmentation by upsampling every map, taking the sigmoid, copying it to
 * the host and thresholding it 101 times (Dice sweep, exp/cxr_pt/inference/segmentation_utils.py:
 * 255-261; compute_specificity :136-158).  This entry point computes the sufficient statistics of
 * that sweep in one pass without writing the pixel map.  Geometry arguments as rz_upsample_maps.
 *   gt_masks          optional uint8 [maps, out_h, out_w] (non-zero = ground truth)
 *   thresholds_logit  fp32 [n_thresholds] ascending, in the SCORE domain (logit of the probability
 *                     thresholds; sigmoid(v) > t  <=>  v > logit(t)); n_thresholds <= 127
 *   hist_all, hist_gt uint32 [maps, n_thresholds + 1]: bin k counts the canvas pixels (all / ground
 *                     truth only) whose score exceeds exactly k thresholds, so
 *                     |P_t_j| = sum_{k > j} hist_all[k] and |P_t_j & G| = sum_{k > j} hist_gt[k]
 *   max_score         fp32 [maps]: maximum interpolated score (a map has a pixel above t iff
 *                     max_score > logit(t))
 */
int rz_map_threshold_stats(const float* scores, long long map_stride, int maps, int grid, int out_h,
                           int out_w, int interp_h, int interp_w, int off_y, int off_x, float fill,
                           const unsigned char* gt_masks, const float* thresholds_logit,
                           int n_thresholds, unsigned int* hist_all, unsigned int* hist_gt,
                           float* max_score, void* stream);

/* ---- K10: multi-positive NCE loss, forward + backward ----------------------------------
 * Replaces multi_positive_nce_loss + get_row_loss + get_col_loss (losses.py:243-344) and
 * their autograd, for the image-sharded layout of SURVEY.md section 8e: this rank holds
 * columns [col0, col0 + b_local) of the (n_total x b_global) logit matrix.
 * TWO launches (one persistent cooperative kernel each), one on either side of the cross-rank
 * reduction.  The temperature is exp(*log_tau) read ON THE DEVICE when `log_tau` is non-NULL
 * (the loss_temperature parameter: no host synchronisation), else 1 / inv_tau.
 *
 * Launch 1 (rz_mpnce_partials): E = exp(Z/tau); per-row local sums rowsum[i] = sum_b E_ib,
 *   pos[i] = E[i, group_map[i]] if that column is local else 0, and the column state
 *   colstate[5][b4] (b4 = b_local rounded up to 4): colneg[b] = sum_{i: g_i != b} E_ib,
 *   colpos[b] = sum_{i: g_i = b} E_ib, and the column coefficients acol / apos / lcol of the
 *   backward and the column loss terms (they depend on local quantities only), all in a fixed
 *   summation order (no float atomics) so that 1-GPU and N-GPU runs agree.  eps / col_sum / b_global
 *   must be the values launch 2 is given.
 *   With several ranks the caller all-reduces (sum) rowsum and pos between the launches.
 *   scratch1: fp32 [rz_mpnce_partials_scratch_floats(n_total, b_local)].
 * Launch 2 (rz_mpnce_finish): loss terms and dL/dZ for the local columns.
 *   loss_terms[0] = sum of the row terms this rank owns (rows whose positive column is
 *   local; images, when row_sum), loss_terms[1] = sum of its column terms,
 *   loss_terms[2] = sum_ib dZ_ib * Z_ib over the local block (= -dL/dlog(tau) share),
 *   loss_terms[3] = this rank's share of the loss = terms[0]/(2 n_row) + terms[1]/(2 n_col)
 *   (the loss itself on one rank; summed over ranks otherwise)
 *   with n_row = b_global if row_sum else n_total, n_col = b_global if col_sum else n_total;
 *   dZ already carries the 1/(2*n_row), 1/(2*n_col) factors.
 *   row_sum / col_sum select the MIL-NCE variants (losses.py:303-315, 331-336).
 *   z, dz    [n_total, ldz] fp32 (b_local valid columns per row); dz may be NULL (loss only)
 *   group_map int64 [n_total] GLOBAL image index
 *   colstate the block launch 1 wrote (16-byte aligned)
 *   scratch2 fp32 [rz_mpnce_finish_scratch_floats(n_total, b_local, b_global)]
 */
size_t rz_mpnce_partials_scratch_floats(int n_total, int b_local);
size_t rz_mpnce_finish_scratch_floats(int n_total, int b_local, int b_global);
int rz_mpnce_partials(const float* z, long long ldz, int n_total, int b_local, int b_global,
                      const long long* group_map, int col0, float inv_tau, const float* log_tau,
                      float eps, int col_sum, float* rowsum, float* pos, float* colstate,
                      float* scratch1, void* stream);
int rz_mpnce_finish(const float* z, long long ldz, int n_total, int b_local, int b_global,
                    const long long* group_map, int col0, float inv_tau, const float* log_tau,
                    float eps, int row_sum, int col_sum,
                    const float* rowsum, const float* pos, const float* colstate,
                    float* scratch2, float* dz, float* loss_terms, void* stream);

/* ---- group_map ---------------------------------------------------------------------------
 * group_map[j] = first_image + (image of sentence j), from the per-image sentence counts
 * (compute_text_features, losses.py:131-151: `global_index = i + local_rank * B_local`).
 * counts_host is a HOST array; the counts reach the device as kernel parameters, so nothing queues on
 * the copy engine.  out: int64 [sum(counts)].  counts must be in [0, 65535].
 */
int rz_group_map(const int* counts_host, int n_images, long long first_image, long long* out, void* stream);

/* ---- K11: image preprocessing of the zero-shot evaluators ------------
 * Replaces, for `images` same-sized raw images resident in device memory, the host-side chain
 *   collate_fn (exp/cxr_pt/inference/dataset.py:31-51): cv2.normalize(img, None, 0, 255, NORM_MINMAX, CV_8U)
 *   image_processor (exp/cxr_pt/model/processing.py:85-101, BlipImageProcessor at 518): convert to RGB,
 *     PIL bicubic resize, * rescale_factor, (x - mean) / std, channels first
 * with BIT-IDENTICAL results (oracle/preprocess.py pins the arithmetic against cv2 / Pillow / transformers).
 *   raw          [images, height, width, channels] of `raw_dtype`, contiguous; channels = 1 (grey, replicated
 *                to RGB) or 3 (interleaved RGB); min / max are taken over all channels, as cv2 does
 *   mean_host, std_host  HOST pointers to 3 floats (image_mean / image_std as float32)
 *   pixel_values [images, 3, out_h, out_w] of `out_dtype` (RZ_F32 is the reference's; RZ_F16 / RZ_BF16 round it)
 *   workspace    rz_preprocess_workspace_bytes(...) bytes, 256-byte aligned
 */
#define RZ_IMG_U8 0
#define RZ_IMG_U16 1
#define RZ_IMG_I16 2
#define RZ_IMG_I32 3
#define RZ_IMG_F32 4
size_t rz_preprocess_workspace_bytes(int images, int height, int width, int channels, int out_h, int out_w);
int rz_preprocess_images(const void* raw, int raw_dtype, int images, int height, int width, int channels,
                         int out_h, int out_w, const float* mean_host, const float* std_host,
                         double rescale_factor, void* pixel_values, int out_dtype, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- T0: text side ------------------------------------------
 * Masked mean pooling of the text encoder's token embeddings (exp/cxr_pt/model/modeling.py:147-156,
 * text_encoders.py:32-41) fused with the path's LayerNorm + L2 normalisation of the pooled vector
 * (losses.py:163-164, 212-213), for ALL prompts / sentences in one launch (the reference calls the
 * text model once per prompt, modeling.py:290-298, and normalises afterwards).
 *   hidden          [n_sentences, tokens, 768] of `dtype` (last_hidden_state, padded batch)
 *   attention_mask  int64 [n_sentences, tokens]; tokens whose mask is 0 are never read
 *   feats_f32       optional fp32 [n_sentences, 768] = text_features_wo_l2_norm
 *   q_f16           optional fp16 [n_sentences, 768] = rows as rz_prep_rows would write them from
 *                   feats_f32 (gamma/beta NULL: no LayerNorm; l2 as in rz_prep_rows)
 */
int rz_text_pool(const void* hidden, int dtype, const long long* attention_mask, int n_sentences,
                 int tokens, const float* gamma, const float* beta, int l2, float* feats_f32,
                 void* q_f16, void* stream);

/* ---- A0-A2: the AlignTransformer in front of the path ------
 * The vision tokens the VL-CABS path consumes are produced by AlignTransformer.forward
 * (exp/cxr_pt/model/align_transformers.py:37-45): a transformers `Dinov2Encoder` of two layers
 * (radzero.yaml:29-33), each   h = x + ls1 * attn(norm1(x));  y = h + ls2 * fc2(gelu(fc1(norm2(h)))).
 * These three entry points are the kernels of that forward (inference); the layer sequencing is
 * host code (radzero_b200/align.py).  All operands are fp16 with fp32 accumulation; the residual
 * stream stays fp32.
 *
 * rz_ln_rows   nn.LayerNorm with the caller's eps (Dinov2Layer.norm1 / norm2, eps 1e-6):
 *              x [rows, 768] of `dtype` -> out_f16 [rows, 768].
 * rz_linear    out = epilogue(a . w^T + bias): a fp16 [m, k], w fp16 [n, k] (nn.Linear.weight
 *              layout), bias fp32 [n] or NULL; k % 64 == 0, n % 256 == 0.
 *                RZ_LIN_BIAS      out fp16 [m, n] = acc + bias            (query/key/value fused)
 *                RZ_LIN_GELU      out fp16 [m, n] = gelu_erf(acc + bias)  (mlp.fc1 + activation)
 *                RZ_LIN_RESIDUAL  out fp32 [m, n] = residual + scale * (acc + bias)
 *                                 (attention.output.dense / mlp.fc2 + Dinov2LayerScale + the
 *                                 residual add; scale fp32 [n] or NULL; out may alias residual)
 * rz_attention softmax(q k^T) v per (image, head), head dim 64, no mask (Dinov2SelfAttention):
 *              qkv fp16 [n_images, tokens, 3 * heads * 64] = [q | k | v] column blocks with the
 *              1/sqrt(64) scale already folded into q; out fp16 [n_images, tokens, heads * 64].
 */
#define RZ_LIN_BIAS 0
#define RZ_LIN_GELU 1
#define RZ_LIN_RESIDUAL 2
#define RZ_LIN_RESIDUAL_F16 3 /* residual + scale * (acc + bias) written ONLY as fp16 [m, n]: the last layer's
                                 tokens handed to rz_sim_fwd_tokens at half the bytes (out must not alias residual) */
int rz_ln_rows(const void* x, int dtype, const float* gamma, const float* beta, float eps,
               long long rows, void* out_f16, void* stream);
int rz_linear(const void* a_f16, long long m, int k, const void* w_f16, int n, const float* bias,
              int epilogue, const float* scale, const float* residual, void* out, void* stream);
int rz_attention(const void* qkv_f16, int n_images, int tokens, int heads, void* out_f16,
                 void* stream);

/* ---- backward of the AlignTransformer layers ------------------------------------------------
 * Autograd of transformers' Dinov2Layer as AlignTransformer.forward runs it
 * (exp/cxr_pt/model/align_transformers.py:37-45; trained by radzero.yaml `module_to_update`).  The
 * GEMM-shaped products go through rz_linear (dX = dY . W with the weight pre-transposed, dW = dY^T . X
 * with both operands transposed by rz_transpose_pad and the fp32 RZ_LIN_RESIDUAL epilogue accumulating
 * into the gradient); these entry points are the steps between them.  All fp16 gradients carry one
 * power-of-two factor 2^k chosen on the device from max|dL/dtokens| (rz_grad_scale); every fp32
 * result is multiplied by 2^-k.
 *
 * rz_grad_scale     sc fp32 [rz_grad_scale_floats()]: sc[0] = 2^k, sc[1] = 2^-k, sc[2] = max|grad|,
 *                   sc[3] = k, sc[4 ..] = 2^-k repeated (the `scale` vector of rz_linear).
 * rz_ls_cast_bwd    Dinov2LayerScale: do_f16 [rows, 768] = fp16(2^k ls dy); dls [768] += sum_rows dy * o
 *                   (o_f16 = the recomputed output of the scaled linear layer; o_f16 / dls may be NULL).
 *                   ls NULL = 1 (a plain scaled cast).
 * rz_ls_weight_bwd  Dinov2LayerScale backward WITHOUT recomputing the scaled product o = x W^T + b: with
 *                   g fp32 [n, k] = dy^T x (the weight gradient of the unscaled product) and colsum [n] = sum_rows dy,
 *                   in place g[n, :] *= ls[n] (= dW), colsum[n] *= ls[n] (= db), and
 *                   dls[n] = sum_k w[n, k] g[n, k] + bias[n] colsum[n] (overwritten).  w fp32 [n, k] = the weight.
 * rz_transpose_pad  in fp16 [rows, cols] -> out fp16 [cols, rows_padded] (zero for rows >= `rows`;
 *                   cols % 64 == 0, rows_padded % 64 == 0); colsum [cols] += 2^-k sum_rows in
 *                   (the bias gradient; out or colsum may be NULL).
 * rz_gelu_bwd       du = dg * gelu_erf'(u), n fp16 elements (n % 8 == 0).
 * rz_ln_rows_bwd    nn.LayerNorm(768, eps): dx fp32 [rows, 768] = dres + 2^-k LN'(dh_f16) (dres NULL = 0, dx
 *                   may alias dres); dx_f16 (optional) = fp16(2^k dx), the operand of the next product;
 *                   dgamma / dbeta [768] += the parameter gradients (NULL = skipped).
 * rz_attention_bwd  backward of rz_attention: qkv / out as there, dout fp16 [n_images, tokens, heads * 64];
 *                   dqkv fp16 like qkv, its q block multiplied by q_scale (the 1/sqrt(64) the host folded
 *                   into the query projection: dqkv is then the gradient of the UNSCALED projections);
 *                   lse, delta fp32 [n_images * heads * tokens rounded up to 64] scratch, 16-byte aligned.
 */
size_t rz_grad_scale_floats(void);
int rz_grad_scale(const float* grad, long long n, float* sc, void* stream);
int rz_ls_cast_bwd(const float* dy, const float* ls, const void* o_f16, const float* sc, long long rows,
                   void* do_f16, float* dls, void* stream);
int rz_ls_weight_bwd(float* g, const float* w, const float* bias, float* colsum, const float* ls, int n, int k,
                     float* dls, void* stream);
int rz_transpose_pad(const void* in_f16, long long rows, int cols, long long rows_padded, void* out_f16,
                     float* colsum, const float* sc, void* stream);
int rz_gelu_bwd(const void* dg_f16, const void* u_f16, long long n, void* du_f16, void* stream);
int rz_ln_rows_bwd(const float* x, const void* dh_f16, const float* gamma, float eps, const float* dres,
                   const float* sc, long long rows, float* dx, void* dx_f16, float* dgamma, float* dbeta,
                   void* stream);
int rz_attention_bwd(const void* qkv_f16, const void* out_f16, const void* dout_f16, int n_images,
                     int tokens, int heads, float q_scale, float* lse, float* delta, void* dqkv_f16,
                     void* stream);

/* ---- diagnostics: one tcgen05.mma probe -------------------------------------------------
 * Copies caller-built shared-memory images of A and B into smem, issues `k_steps`
 * tcgen05.mma (kind::f16, fp32 accumulate) with the given descriptors and dumps all 128
 * TMEM lanes x `ncols` columns.  Used by tests/test_umma_probe.py to pin the descriptor and
 * TMEM layouts the production kernels rely on.  Not part of the reference surface.
 */
int rz_umma_probe(const void* a_image, int a_bytes, const void* b_image, int b_bytes,
                  unsigned long long a_desc, unsigned long long b_desc,
                  int a_step_bytes, int b_step_bytes, int k_steps, unsigned int idesc,
                  unsigned int d_tmem_offset, int ncols, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RZ_B200_H_ */
