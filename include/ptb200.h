/* ptb200 -- C ABI of the B200-native Probabilistic Teacher hot path (libptb200.so).
 *
 * Conventions
 *  - Every entry point takes plain device pointers, sizes and a cudaStream_t (as void*), returns 0 on
 *    success or a non-zero error code (cudaError_t values < 1000, argument errors >= 1000), and never
 *    synchronises the stream. Pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *  - Activations are fp16 "NHWC-flat": [N][H][Wp][C] with Wp = W + 1; the pad column x = W is kept at
 *    zero so that the 3x3 taps of a convolution are plain row shifts of the flattened [H*Wp] axis.
 *  - Variable-length results live in fixed-capacity buffers with a device-side count per image.
 *  - Each declaration cites the reference interface it replaces (paths relative to the reference
 *    checkout of hikvision-research/ProbabilisticTeacher; "d2" = detectron2 v0.5, un-vendored).
 */
#ifndef PTB200_H
#define PTB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Epilogues of ptb200_gemm_tn_f16. */
#define PTB200_EPI_BIAS_RELU_F16 0
#define PTB200_EPI_BIAS_F16 1
#define PTB200_EPI_F32_SPLIT 2
#define PTB200_EPI_MASK_F16 3
#define PTB200_EPI_ATOMIC_F32 4 /* split-K partials: d0[row][n] += acc (fp32 atomics) */
#define PTB200_EPI_SPLIT3_RELU_F16 5 /* f16x3 only: relu(alpha*acc + bias) -> [hi | lo | hi] triple */
#define PTB200_EPI_SPLIT3_F16 6      /* f16x3 only: alpha*acc + bias -> [hi | lo | hi] triple */
#define PTB200_EPI_SPLIT3_MASK_F16 7 /* f16x3 only: (aux > 0 ? alpha*acc : 0) -> triple (ReLU backward) */
#define PTB200_EPI_F32_STORE 8       /* f16x3 only: alpha*acc + bias -> fp32 d0[row][n] */

/* ---- dense contractions (tcgen05 tensor cores, TMA-staged tiles) ------------------------------ */

/* Implicit GEMM  D[b][p][n] = sum_t sum_k A[b][p + shifts[t]][k] * B[n][t*k_per_tap + k]  (+ epilogue).
 * fp16 operands, fp32 accumulation. Replaces the cuDNN / cuBLAS calls behind d2 Conv2d and nn.Linear:
 *   pt/modeling/backbone/vgg.py:45-53,65-72 (3x3 conv + bias + ReLU),
 *   pt/modeling/proposal_generator/rpn.py:96 (RPN head), pt/modeling/roi_heads/roi_heads.py:127-128
 *   (box head), pt/modeling/roi_heads/fast_rcnn.py:157-169 (predictor), and their data-gradients.
 * Rows outside [0, rows) read as zero (conv zero padding). ksplit > 1 (with PTB200_EPI_ATOMIC_F32)
 * splits the reduction across CTAs for skinny problems (fc1 forward): with split == 0 the partial sums are
 * added into d0 with fp32 atomics; with split == 1 d0 is [ksplit][batch][rows][ld0] and every K split stores its
 * own slice (ptb200_bias_act_cast_f16 adds them in a fixed order: a bit-reproducible forward). seg_counts (may be NULL, batch == 1):
 * rows form segments of seg_cap rows of which only the first seg_counts[s] are live (fixed-capacity roi
 * buffers); 128-row tiles without a live row are skipped and their outputs left untouched. */
int ptb200_gemm_tn_f16(const void* A, int batch, int rows, int k_per_tap, int64_t lda,
                       int64_t a_batch_stride, int taps, const int* shifts_host, const void* B,
                       int n_total, int bn, int epi, const float* bias, int n_bias, void* D,
                       int64_t ldd, int64_t d_batch_stride, const void* aux, int w_valid, int wp,
                       float* d0, int ld0, float* d1, int ld1, int split, int n_valid, int max_ctas,
                       int ksplit, const int* seg_counts, int seg_cap, void* stream);

/* Split-fp16 ("f16x3") fp32-equivalent variant of ptb200_gemm_tn_f16, forward AND backward: the reference
 * computes these contractions and their autograd in fp32 (AMP off, pt/engine/trainer.py:271-277,383-386 run
 * without autocast), so the 1e-3 parity claim of losses AND parameter gradients is checked -- and a training
 * step is timed -- in this mode. Every fp32 value x travels as two fp16 numbers hi = fp16(x), lo = fp16(x - hi);
 * an activation (or output-gradient) row is the K-concatenation [hi | lo | hi] (3 * channels wide), a weight
 * row per tap is [Wh | Wh | Wl] of W * 2^s (ptb200_split3_pack_f16 / ptb200_transpose_pack_f16x3), so the
 * tensor-core main loop accumulates hi*Wh + lo*Wh + hi*Wl in fp32 (the lo*Wl term, 2^-22 relative, is
 * dropped). alpha = 2^-s is applied to the accumulator before the bias. k3_per_tap = 3 * K.
 * The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of chained
 * MMAs; the kernel therefore runs the reduction in chunks of `chunk` k-iterations of 64 (0 = default 4; the
 * row-window / pair modes promote per filter-row pair of windows -- at most 36 chained MMAs -- unless chunk >= 64) on
 * alternating TMEM accumulator stages and its epilogue warps promote every finished chunk into fp32 register
 * accumulators (round to nearest) while the next chunk runs.
 * Epilogues: PTB200_EPI_SPLIT3_{RELU_,}F16 (D3 is [batch][rows][3*n_total]); PTB200_EPI_SPLIT3_MASK_F16 (ReLU
 * backward fused into a data-gradient GEMM: aux3 = the forward activation triples, same geometry as D3);
 * PTB200_EPI_F32_STORE (fp32 d0[row][n], ld0 = row pitch; with ksplit > 1 -- skinny
 * problems -- d0 is [ksplit][batch][rows][ld0], one partial sum per K split, finished in a fixed order by
 * ptb200_bias_act_split3_f16: no atomics, the forward is bit-reproducible); PTB200_EPI_ATOMIC_F32 with ksplit > 1
 * (red.add partial sums); PTB200_EPI_F32_SPLIT (narrow fp32 heads, single chain).
 * bn must be 64, 128 or 256 except for PTB200_EPI_F32_SPLIT. */
int ptb200_gemm_tn_f16x3(const void* A3, int batch, int rows, int k3_per_tap, int64_t lda,
                         int64_t a_batch_stride, int taps, const int* shifts_host, const void* B3,
                         int n_total, int bn, int epi, const float* bias, int n_bias, void* D3, int64_t ldd,
                         int64_t d_batch_stride, int w_valid, int wp, float* d0, int ld0, float* d1, int ld1,
                         int split, int n_valid, int max_ctas, int ksplit, const int* seg_counts, int seg_cap,
                         float alpha, const void* aux3, int chunk, void* stream);

/* Bias + ReLU of the same layers (pt/modeling/backbone/vgg.py:65-72, the box head of roi_heads.py:127-128) applied to
 * the fp32 partial sums of a split-K f16x3 GEMM:
 * out3 = triple(act(alpha * sum_s in[s] + bias)), fp32 [slices][rows][n] added in slice order -> fp16 [rows][3n];
 * wp > 0: rows with (row % wp) >= w_valid are written as zero (pad column of the flat activation layout). */
int ptb200_bias_act_split3_f16(const float* in, int slices, const float* bias, int relu, float alpha, int64_t rows,
                               int n, int wp, int w_valid, void* out3, void* stream);

/* Operand preparation that has no counterpart in the reference (it feeds fp32 tensors to cuDNN / cuBLAS at
 * vgg.py:45-53, rpn.py:96, fast_rcnn.py:157-169 directly):
 * fp32 [rows][k] -> f16x3 operand [rows][3k] of src * scale: order 0 = activation triple [hi | lo | hi],
 * order 1 = weight triple [hi | hi | lo]. (rows counts (output channel, tap) pairs for conv weights.) */
int ptb200_split3_pack_f16(const float* src, void* dst, int64_t rows, int k, float scale, int order,
                           void* stream);

/* f16x3 triple [rows][3k] -> fp32 [rows][k] (hi + lo); inspection / tests (gives back the fp32 tensor the
 * reference's layer would have produced, e.g. `features["vgg_block5"]` of pt/modeling/meta_arch/rcnn.py:45). */
int ptb200_split3_unpack_f32(const void* src, float* dst, int64_t rows, int k, void* stream);

/* Weight gradient  out[m][t*n_total + n] += scale * sum_b sum_p G[b][p][m] * X[b][p + shifts[t]][n]
 * (split-K, fp32 atomic accumulation). Replaces the wgrad kernels autograd issues at
 * pt/engine/trainer.py:384 (`losses.backward()`). m_total % 128 == 0, n_total % 64 == 0.
 * bias_out (may be NULL): bias_out[m] += scale * sum_b sum_p G[b][p][m] (bias gradient, fused). */
int ptb200_gemm_wgrad_f16(const void* G, int64_t ldg, int64_t g_batch_stride, const void* X,
                          int64_t ldx, int64_t x_batch_stride, int batch, int rows, int m_total,
                          int n_total, int taps, const int* shifts_host, float* out, int64_t ld_out,
                          float scale, int ksplit, float* bias_out, const int* seg_counts, int seg_cap,
                          void* stream);

/* f16x3 (fp32-equivalent) weight gradient, the reference's fp32 autograd of the same layers
 * (pt/engine/trainer.py:383-386 through vgg.py:45-53, rpn.py:44-55, roi_heads.py:127-128, fast_rcnn.py:157-169):
 * G3 [batch][rows][3*m_total] and X3 [batch][rows][3*n_total] are [hi | lo | hi] triples (ldg / ldx = their row
 * pitches); out += scale * (Gh'Xh + Gl'Xh + Gh'Xl) in ONE launch (a pipeline stage holds the hi and lo tiles of
 * both operands), bias_out += scale * colsum(Gh + Gl). Other arguments as ptb200_gemm_wgrad_f16. */
int ptb200_gemm_wgrad_f16x3(const void* G3, int64_t ldg, int64_t g_batch_stride, const void* X3, int64_t ldx,
                            int64_t x_batch_stride, int batch, int rows, int m_total, int n_total, int taps,
                            const int* shifts_host, float* out, int64_t ld_out, float scale, int ksplit,
                            float* bias_out, const int* seg_counts, int seg_cap, void* stream);

/* ---- image / activation helpers ---------------------------------------------------------------- */

/* d2 GeneralizedRCNN.preprocess_image (called at pt/modeling/meta_arch/rcnn.py:38-43): uint8 CHW
 * images -> (x - mean) / std, zero-padded to the batch max size, emitted directly as the im2col
 * operand [n][hmax*(wmax+1)][64] (27 live columns, order (ky*3+kx)*3+c) of the first VGG conv. */
int ptb200_preprocess_im2col(const uint8_t* images, const int* hw_dev, int n, int hmax, int wmax,
                             int64_t image_stride, const float* mean3_host, const float* std3_host,
                             void* out_f16, void* stream);

/* Same pre-processing (d2 preprocess_image, rcnn.py:38-43) fused with the first VGG conv (vgg_block1.conv1 of
 * pt/modeling/backbone/vgg.py:45-53,65-72, 3 -> 64, bias + ReLU): uint8 CHW
 * images -> fp16 NHWC-flat [n][hmax*(wmax+1)][64]. wpack_f16 = weights [64][32] fp16, k = (ky*3+kx)*3+c. */
int ptb200_conv1_u8_f16(const uint8_t* images, const int* hw_dev, int n, int hmax, int wmax,
                        int64_t image_stride, const float* mean3_host, const float* std3_host,
                        const void* wpack_f16, const float* bias, void* out_f16, void* stream);

/* f16x3 variant of ptb200_conv1_u8_f16 (rcnn.py:38-43 + vgg.py:45-53,65-72): the 3 -> 64 conv is evaluated in fp32
 * on the CUDA cores (K = 27),
 * output [n][hmax*(wmax+1)][192] triples. w_f32 = [64][27] (k = (ky*3+kx)*3+c). */
int ptb200_conv1_u8_f16x3(const uint8_t* images, const int* hw_dev, int n, int hmax, int wmax,
                          int64_t image_stride, const float* mean3_host, const float* std3_host,
                          const float* w_f32, const float* bias, void* out_f16x3, void* stream);

/* Tensor-core version of the same fused call (rcnn.py:38-43 + vgg.py:45-53,65-72) in the f16x3 precision: raw pixel
 * values (exact in fp16) are the A operand, [p | p], against wpack3 = fp16 [64][64] rows
 * [Wh(27) 0(5) | Wl(27) 0(5)] of w / std * 2^s (alpha = 2^-s); the mean enters through bias_table = fp32 [10][64]:
 * rows 0..8 = S_t[co] = sum_c w[co][t][c] / std_c * mean_c, row 9 = bias - sum_t S_t (pixels with taps outside their
 * image add the S_t of the absent taps back: zero padding of the NORMALISED image, as d2 ImageList pads). */
int ptb200_conv1_u8_f16x3_tc(const uint8_t* images, const int* hw_dev, int n, int hmax, int wmax,
                             int64_t image_stride, const void* wpack3_f16, const float* bias_table, float alpha,
                             void* out_f16x3, void* stream);

/* PTrainer.resize (pt/engine/trainer.py:557-590) for one CHW uint8 image: bilinear down-scale to
 * (dh, dw) pasted at (x1, y1) on a canvas filled with int(pixel_mean). */
int ptb200_resize_paste_u8(const uint8_t* src, uint8_t* dst, int h, int w, int dh, int dw, int x1, int y1,
                           int m0, int m1, int m2, void* stream);

/* Same (pt/engine/trainer.py:557-590), with the geometry {dh, dw, x1, y1} read from device memory (CUDA-graph
 * replay: the random ratio of :561 is drawn on the host and copied into a persistent device buffer). */
int ptb200_resize_paste_u8_dev(const uint8_t* src, uint8_t* dst, int h, int w, const int* params_dev, int m0,
                               int m1, int m2, void* stream);

/* F.max_pool2d(2, 2) of pt/modeling/backbone/vgg.py:59,71 (floor mode). */
int ptb200_maxpool2x2_f16(const void* in, void* out, int n, int h, int w, int c, void* stream);

/* f16x3 variant of F.max_pool2d(2, 2) (vgg.py:59,71; c = channels, rows are 3*c wide): the max is taken over
 * hi + lo. */
int ptb200_maxpool2x2_f16x3(const void* in, void* out, int n, int h, int w, int c, void* stream);

/* Backward of ReLU followed by the 2x2 max pool (autograd of vgg.py:65-72). */
int ptb200_maxpool2x2_relu_bwd_f16(const void* x, const void* dpooled, void* dz, int n, int h, int w,
                                   int c, void* stream);

/* f16x3 counterpart (autograd of vgg.py:59,65-72 in the reference's fp32): x3 = pre-pool activation triples
 * [n][h*(w+1)][3c], dpooled = fp32 gradient of the pooled map [n][(h/2)*(w/2+1)][c], dz3 = gradient triples. */
int ptb200_maxpool2x2_relu_bwd_f16x3(const void* x3, const float* dpooled, void* dz3, int n, int h, int w, int c,
                                     void* stream);

/* fp32 master weights -> fp16 GEMM operands: the cast the reference never needs (fp32 nn.Parameter tensors go to
 * cuDNN / cuBLAS as they are; vgg.py:45-53, rpn.py:44-55, fast_rcnn.py:157-169). transpose_pack builds the
 * [Cin][tap][Cout] operand of the data-gradient GEMM that autograd derives from the same weights
 * (pt/engine/trainer.py:384); cast_pad_rows pads the K = 27 first-conv filter to 32. */
int ptb200_cast_f32_f16(const float* src, void* dst, int64_t n, void* stream);
int ptb200_transpose_pack_f16(const float* src, void* dst, int rows, int cols, int taps, int flip,
                              int64_t ld_dst, void* stream);
int ptb200_cast_pad_rows_f16(const float* src, void* dst, int rows, int cols, int ld_dst, void* stream);
/* f16x3 data-gradient operand of the same weights (autograd of vgg.py:45-53, rpn.py:44-55, roi_heads.py:127-128,
 * fast_rcnn.py:157-169 in fp32): src fp32 [rows][taps][cols] -> dst3[col][tap (flipped when flip)][Wh | Wh | Wl]
 * (3*rows wide) of src * scale. */
int ptb200_transpose_pack_f16x3(const float* src, void* dst3, int rows, int cols, int taps, int flip, float scale,
                                void* stream);

/* Bias gradient (autograd of the `+ bias` in d2 Conv2d / nn.Linear, run by `losses.backward()` at
 * pt/engine/trainer.py:384): out[c] += scale * sum_rows in[row][c]. */
int ptb200_colsum_f16(const void* in, int64_t rows, int c, int64_t ld, float scale, float* out,
                      void* stream);

/* Head of the backward chain (autograd of the loss sums at pt/engine/trainer.py:364-384 w.r.t. the two head
 * outputs: objectness + anchor deltas of rpn.py:96, cls_score + bbox_pred of fast_rcnn.py:157-169):
 * packs the unit gradients of two loss terms (n0 and n1 columns) scaled by the upstream gradients
 * g0[0], g1[0] and the loss scale into the fp16 [rows][ld] operand of the backward GEMMs. */
int ptb200_pack_grad2_f16(const float* d0, int n0, const float* d1, int n1, const float* g0,
                          const float* g1, float lscale, int64_t rows, int ld, void* out, void* stream);

/* f16x3 counterpart of ptb200_pack_grad2_f16 (same autograd entry, pt/engine/trainer.py:364-384, in the
 * reference's fp32): out3 = gradient triples [rows][3*ld]. */
int ptb200_pack_grad2_f16x3(const float* d0, int n0, const float* d1, int n1, const float* g0, const float* g1,
                            float lscale, int64_t rows, int ld, void* out3, void* stream);

/* Finishes a split-K GEMM (fc1 of the box head, roi_heads.py:127-128): out = half(act(sum_s in[s] + bias)),
 * in = fp32 [slices][rows][n] added in slice order. */
int ptb200_bias_act_cast_f16(const float* in, int slices, const float* bias, int relu, int64_t rows, int n,
                             void* out, void* stream);

/* out = half(a + scale * b): autograd's accumulation of two gradient paths into one tensor (the RPN and ROI
 * branches both consume `features["vgg_block5"]`, pt/modeling/meta_arch/rcnn.py:45-61). */
int ptb200_add_f32_to_f16(const void* a, const float* b, float scale, void* out, int64_t n,
                          void* stream);

/* out = (aux > 0) ? a + scale * b : 0: the same accumulation fused with the backward of the last backbone ReLU
 * (vgg.py:65-72; sums the RPN and ROI gradient paths into the backbone output, rcnn.py:45-61). */
int ptb200_add_mask_f16(const void* a, const float* b, float scale, const void* aux, void* out,
                        int64_t n, void* stream);

/* f16x3 counterpart (rcnn.py:45-61 + the last ReLU of vgg.py:65-72, fp32 autograd): out3 = triple of
 * ((aux_hi + aux_lo) > 0 ? a + b : 0); a (may be NULL), b fp32 [rows][c]; aux3, out3 triples [rows][3c]. */
int ptb200_add_mask_f16x3(const float* a, const float* b, const void* aux3, void* out3, int64_t rows, int c,
                          void* stream);

/* ---- sorting ----------------------------------------------------------------------------------- */

/* Segmented stable radix sort, ascending uint32 keys with uint32 payload; replaces torch.sort at
 * pt/modeling/proposal_generator/proposal_utils.py:87, the sort inside torchvision nms and
 * torch.randperm in d2 subsample_labels. (end_bit - begin_bit) must be a multiple of 16. */
int ptb200_segmented_sort_u32(uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp,
                              int segments, int64_t seg_stride, const int* seg_len_dev, int max_len,
                              int begin_bit, int end_bit, void* stream);

/* ---- anchors / proposals / NMS ------------------------------------------------------------------ */

/* pt/modeling/anchor_generator.py:145-148 (cell anchors from the learnable (w,h)) and :108-122 (grid). */
int ptb200_cell_anchors_from_wh(const float* wh, int num_cell, float* cell, void* stream);
int ptb200_anchor_grid(const float* cell, int num_cell, int h, int w, float stride, float offset,
                       float* anchors, void* stream);

/* Sort keys (descending logit, stable) for proposal_utils.py:87. */
int ptb200_rpn_make_keys(const float* logits, int ld, int n, int h, int w, int num_cell,
                         uint32_t* keys, uint32_t* vals, void* stream);

/* proposal_utils.py:92-138 for the k best anchors of every image: decode (box_regression.py:101-139),
 * finite check, clip, non-empty filter, sigma re-scoring (sigma read at the UNSORTED index, :94). */
int ptb200_rpn_topk_decode(const uint32_t* sorted_idx, int64_t idx_stride, const float* logits,
                           int ld_logit, const float* deltas, int ld_delta, const float* anchors, int n,
                           int h, int w, int num_cell, int k, const float* img_hw, float min_size,
                           float* boxes, float* scores, uint32_t* keys2, uint32_t* vals2,
                           int* valid_count, int* nonfinite_flag, void* stream);

/* d2 batched_nms -> torchvision nms (proposal_utils.py:140, fast_rcnn.py:104): greedy NMS over the
 * candidates in `order` (descending score); class_mod > 0 restricts suppression to candidates with
 * equal (order value % class_mod). keep_idx holds positions in `order`; the scan stops after max_keep survivors
 * and the suppression bit-mask is only built for the bands of candidate rows the scan reaches. mask_scratch needs
 * n * wpad + n * band_rows * wpad 64-bit words, wpad = roundup2(ceil(cap/64)), band_rows = 64 * max(first band,
 * later bands) as computed by ptb200_nms; passing n * (cap + 1) * wpad words is always enough. */
int ptb200_nms(const float* boxes, int64_t box_stride, const uint32_t* order, int64_t order_stride,
               const int* counts, int n, int cap, float thresh, int class_mod, int max_keep,
               unsigned long long* mask_scratch, int* keep_idx, int* keep_count, void* stream);

int ptb200_rpn_gather(const float* boxes, const float* scores, int k, const uint32_t* order,
                      int64_t order_stride, const int* keep_idx, const int* keep_count, int max_keep,
                      int n, float* out_boxes, float* out_scores, void* stream);

/* Teacher pseudo-label filter, pt/modeling/roi_heads/fast_rcnn.py:34-120,338-409. */
int ptb200_roi_infer_candidates(const float* scores, const float* deltas, const float* props,
                                const int* prop_count, int n, int cap, int num_classes,
                                const float* img_hw, float score_thresh, const float* weights4_host,
                                float* cboxes, float* cscores, uint32_t* keys, uint32_t* vals,
                                int* cand_count, void* stream);
int ptb200_roi_infer_gather(const float* cboxes, const float* cscores, const float* scores,
                            const float* deltas, const uint32_t* order, const int* keep_idx,
                            const int* keep_count, int n, int cap, int num_classes, int topk,
                            float* out_boxes, float* out_scores, int64_t* out_classes, float* out_logits,
                            float* out_sigma, int* out_src_roi, void* stream);

/* ---- matching / sampling ------------------------------------------------------------------------ */

/* d2 pairwise_iou + Matcher([lo, hi], [0,-1,1], allow_low_quality_matches=True), as called at
 * pt/modeling/proposal_generator/rpn.py:414-415. labels in {-1, 0, 1} (before sub-sampling). */
int ptb200_rpn_match(const float* gt_boxes, const int* gt_count, int gt_cap, const float* anchors,
                     int num_anchors, int n, float iou_lo, float iou_hi, float* max_iou_scratch,
                     int* best_per_gt_scratch, int* matched_idx, int* labels, void* stream);

/* d2 subsample_labels building blocks (rpn.py:433, roi_heads.py:223-225). */
int ptb200_compact_pos_neg(const int* labels, int64_t stride, const int* seg_len, int fixed_len,
                           int segments, int bg_label, int* pos_list, int* neg_list, int* counts,
                           void* stream);
int ptb200_prio_keys(const float* prio_pos, const float* prio_neg, int64_t stride, const int* counts,
                     int n, uint32_t* keys, uint32_t* vals, void* stream);
int ptb200_rpn_sample_apply(const int* pos_list, const int* neg_list, const uint32_t* perm,
                            int64_t stride, const int* counts, int n, int num_anchors,
                            int batch_per_image, int max_pos, signed char* labels_out, void* stream);

/* pt/modeling/roi_heads/roi_heads.py:192-255 (supervised) with add_ground_truth_to_proposals
 * (proposal_utils.py:157-224) and d2 _sample_proposals. */
int ptb200_roi_label(const float* gt_boxes, const int* gt_classes, const int* gt_count, int gt_cap,
                     const float* props, const int* prop_count, int prop_cap, int n, int num_classes,
                     float iou_thr, int* cls, int* matched, int* cand_count, void* stream);
int ptb200_roi_sample_apply(const int* pos_list, const int* neg_list, const uint32_t* perm,
                            int64_t stride, const int* counts, const int* cls, const int* matched,
                            const float* gt_boxes, const int* gt_count, int gt_cap, const float* props,
                            const int* prop_count, int prop_cap, int n, int batch_per_image, int max_fg,
                            int num_classes, float* out_rois, int* out_cls, float* out_gt,
                            int* out_count, int* out_src, void* stream);

/* pt/modeling/roi_heads/roi_heads.py:257-291 (_sample_proposals_unsup). */
int ptb200_roi_match_unsup(const float* pseudo_boxes, const float* pseudo_logits,
                           const float* pseudo_sigma, const int* pseudo_count, int pseudo_cap,
                           const float* props, const int* prop_count, int prop_cap, int n,
                           int num_classes_plus1, float iou_thr, float* out_rois, float* out_pseudo,
                           float* out_logits, float* out_sigma, int* out_count, void* stream);

/* ---- ROIAlign ------------------------------------------------------------------------------------ */

/* d2 ROIPooler -> torchvision roi_align(aligned=True, sampling_ratio=0), roi_heads.py:68-73,126.
 * out: fp16 [n*cap][pooled*pooled][c] (the K-major fc1 operand); bwd accumulates into fp32 dfeat. */
int ptb200_roi_align_fwd_f16(const void* feat, int n, int h, int w, int c, const float* rois,
                             const int* roi_count, int cap, float spatial_scale, int pooled, void* out,
                             void* stream);
/* f16x3 variant of the same call (roi_heads.py:68-73,126): feat rows are [hi | lo | hi] triples (3*c wide); a roi's
 * output row is the triple [hi | lo | hi] of the plain [pooled*pooled][c] row (3*pooled*pooled*c wide), i.e. the
 * K-major operand of fc1; bilinear sums in fp32. */
int ptb200_roi_align_fwd_f16x3(const void* feat3, int n, int h, int w, int c, const float* rois,
                               const int* roi_count, int cap, float spatial_scale, int pooled, void* out3,
                               void* stream);
/* Backward of the same call (torchvision roi_align's autograd, run by `losses.backward()` at
 * pt/engine/trainer.py:384): scatters dout [n*cap][pooled*pooled][c] into the fp32 feature gradient dfeat
 * [n][h*(w+1)][c] with vector atomics. */
int ptb200_roi_align_bwd_f16(const void* dout, int n, int h, int w, int c, const float* rois,
                             const int* roi_count, int cap, float spatial_scale, int pooled,
                             float* dfeat, void* stream);

/* Same backward (torchvision roi_align's autograd behind roi_heads.py:68-73,126) with an fp32 output gradient
 * (f16x3 training path: the fc1 data-gradient GEMM stores fp32). */
int ptb200_roi_align_bwd_f32(const float* dout, int n, int h, int w, int c, const float* rois,
                             const int* roi_count, int cap, float spatial_scale, int pooled,
                             float* dfeat, void* stream);

/* ---- losses (fused forward + unit-gradient backward) --------------------------------------------- */

/* pt/modeling/proposal_generator/rpn.py:191-255 (+ box_regression.py:33-35,142-176). loss2 = {cls, loc}. */
int ptb200_rpn_loss_sup(const float* logits, int ld_logit, const float* deltas, int ld_delta,
                        const signed char* labels, const int* matched, const float* gt_boxes, int gt_cap,
                        const float* anchors, int n, int h, int w, int num_cell, float norm, float* loss2,
                        float* dlogits, float* ddeltas, void* stream);

/* pt/modeling/proposal_generator/rpn.py:257-361; danchor_wh (may be NULL) receives d(loss_loc)/d(anchor w,h). */
int ptb200_rpn_loss_unsup(const float* logits, int ld_logit, const float* deltas, int ld_delta,
                          const int* labels, const int* matched, const float* pseudo_boxes,
                          const float* pseudo_logits, const float* pseudo_sigma, int pseudo_cap,
                          const float* anchors, int n, int h, int w, int num_cell, int num_classes_plus1,
                          int efl, float lam0, float lam1, float tau0, float tau1, float norm,
                          float* loss2, float* dlogits, float* ddeltas, float* danchor_wh, void* stream);

/* d2 FastRCNNOutputLayers.losses (mean cross-entropy) + pt/modeling/roi_heads/fast_rcnn.py:265-336. */
int ptb200_roi_loss_sup(const float* scores, const float* deltas, const int* gt_classes,
                        const float* props, const float* gt_boxes, const int* counts, int n, int cap,
                        int num_classes, const float* weights4_host, float* loss2, float* dscores,
                        float* ddeltas, void* stream);

/* pt/modeling/roi_heads/fast_rcnn.py:179-263 + roi_heads.py:131-172. */
int ptb200_roi_loss_unsup(const float* scores, const float* deltas, const float* soft_logits,
                          const float* sigma_t, const float* props, const float* pseudo_boxes,
                          const int* counts, int n, int cap, int num_classes, int efl, float lam0,
                          float lam1, float tau0, float tau1, const float* weights4_host, int* totals2,
                          float* loss2, float* dscores, float* ddeltas, void* stream);

int ptb200_axpy_dev(const float* alpha_dev, float scale, const float* x, float* y, int n, void* stream);

/* ---- input pipeline: strong augmentation on the device ----------------------------- */

/* The per-pixel operations of `build_strong_augmentation` (pt/data/detection_utils.py:38-60: torchvision ColorJitter /
 * RandomGrayscale on a PIL image, Solarize of pt/data/transforms/augmentation_impl.py:40-53), bit-exact with Pillow's
 * Image.blend / Image.convert arithmetic, over a planar uint8 image [3][h][w]. ops_host[i] in {0 brightness,
 * 1 contrast, 2 saturation, 3 hue, 4 grayscale (3 equal channels), 5 solarize(128)}; factors_host[i] = the blend factor
 * (ops 0-2) or the uint8 hue shift np.uint8(hue_factor * 255) (op 3). A contrast op must be the FIRST of its call: its
 * degenerate image is the mean luma of the image at that point of the chain, read from gray_sum_in (sum of the luma of
 * the input, written by the previous call through gray_sum_out; may be NULL when not needed). in == out is allowed. */
int ptb200_aug_pointwise_u8(const uint8_t* in, uint8_t* out, int h, int w, int n_ops, const int* ops_host,
                            const float* factors_host, const unsigned long long* gray_sum_in,
                            unsigned long long* gray_sum_out, void* stream);

/* PIL ImageFilter.GaussianBlur as the reference's GaussianBlur applies it (pt/data/transforms/augmentation_impl.py:22-37):
 * `passes` box-blur passes along x then along y with the 8.24 fixed-point weights ww (full taps) / fw (the two
 * fractional outer taps) of libImaging/BoxBlur.c and edge extension; planes x [h][w] uint8, tmp = scratch of the
 * same size. (radius, ww, fw) from the Gaussian sigma: probabilisticteacher_b200/data_aug.py. */
int ptb200_aug_boxblur_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int planes, int h, int w, int radius,
                          int ww, int fw, int passes, void* stream);

/* Weak augmentation of the same mapper (d2 v0.5 `utils.build_augmentation`: ResizeShortestEdge + RandomFlip, used at
 * pt/data/dataset_mapper.py:67,104-106): one pass of Pillow's `Image.resize(..., BILINEAR)` (libImaging/Resample.c,
 * 8 bpc) along x (along_x = 1: in_h == out_h) or y over planes x [in_h][in_w] uint8; bounds_dev int32 [out][2] = (first
 * input index, tap count), coeffs_dev int32 [out][ksize] = 22-bit fixed-point coefficients (computed on the host as
 * precompute_coeffs / normalize_coeffs_8bpc do: probabilisticteacher_b200/data_aug.py). A resize is the x pass
 * followed by the y pass, each rounded to uint8. */
int ptb200_aug_resample_u8(const uint8_t* in, uint8_t* out, int planes, int in_h, int in_w, int out_h, int out_w,
                           int along_x, const int* bounds_dev, const int* coeffs_dev, int ksize, void* stream);

/* d2 HFlipTransform.apply_image of RandomFlip(horizontal): out[.., x] = in[.., w - 1 - x]. */
int ptb200_aug_hflip_u8(const uint8_t* in, uint8_t* out, int planes, int h, int w, void* stream);

/* ---- optimiser arena ------------------------------------------------------------------------------ */

/* pt/engine/trainer.py:431-449: teacher = keep * teacher + (1 - keep) * student over the flat arena. */
int ptb200_ema_update(float* teacher, const float* student, int64_t n, float keep_rate, void* stream);
/* Same update with the keep rate read from device memory (CUDA-graph replays of the step): the host stages
 * EMA_KEEP_RATE on the iterations where (iter - BURN_UP_STEP) % TEACHER_UPDATE_ITER == 0 and 1.0 on the others
 * (pt/engine/trainer.py:296-298); keep == 1 leaves the teacher untouched bit for bit. */
int ptb200_ema_update_dev(float* teacher, const float* student, int64_t n, const float* keep_rate_dev,
                          void* stream);

/* pt/engine/trainer.py:592-603 (global L2 norm) and torch.optim.SGD(momentum, weight_decay) step
 * (trainer.py:386) fused; the clip coefficient clip/max(norm, clip) is evaluated on device. */
int ptb200_grad_sumsq(const float* grads, int64_t n, float pre_scale, float* sumsq_out, void* stream);
int ptb200_clip_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr,
                         float momentum, float weight_decay, float clip_norm, float pre_scale,
                         const float* sumsq_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
