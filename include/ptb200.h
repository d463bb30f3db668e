/* ptb200 -- C ABI of the B200-native Probabilistic Teacher hot path.
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t (as void*), returns 0 on
 * success or a non-zero error code (cudaError_t values < 1000, argument errors >= 1000), and never
 * synchronises the stream. Each declaration cites the reference interface it replaces
 * (paths relative to the reference checkout of hikvision-research/ProbabilisticTeacher).
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

/* Implicit GEMM  D[b][p][n] = sum_t sum_k A[b][p + shifts[t]][k] * B[n][t*k_per_tap + k]  (+ epilogue).
 * fp16 operands, fp32 accumulation on tcgen05 tensor cores, TMA-staged tiles.
 * Replaces the cuDNN / cuBLAS calls behind detectron2 Conv2d and nn.Linear on the hot path:
 *   pt/modeling/backbone/vgg.py:45-53,65-72 (3x3 conv + bias + ReLU),
 *   pt/modeling/proposal_generator/rpn.py:96 (RPN head), pt/modeling/roi_heads/roi_heads.py:127-128
 *   (box head), pt/modeling/roi_heads/fast_rcnn.py:157-169 (predictor), and their data-gradients.
 * Rows outside [0, rows) read as zero (conv zero padding); `shifts` is a HOST array of `taps` ints. */
int ptb200_gemm_tn_f16(const void* A, int batch, int rows, int k_per_tap, int64_t lda,
                       int64_t a_batch_stride, int taps, const int* shifts, const void* B,
                       int n_total, int bn, int epi, const float* bias, int n_bias, void* D,
                       int64_t ldd, int64_t d_batch_stride, const void* aux, int w_valid, int wp,
                       float* d0, int ld0, float* d1, int ld1, int split, int n_valid, int max_ctas,
                       void* stream);

/* Weight gradient  out[m][t*n_total + n] += scale * sum_b sum_p G[b][p][m] * X[b][p + shifts[t]][n]
 * (fp16 operands, fp32 accumulation, split-K with fp32 atomic accumulation into `out`).
 * Replaces the cuDNN wgrad / cuBLAS calls autograd issues for the layers listed above
 * (pt/engine/trainer.py:384 `losses.backward()`). m_total % 128 == 0, n_total % 64 == 0.
 * ksplit <= 0 selects the split automatically. */
int ptb200_gemm_wgrad_f16(const void* G, int64_t ldg, int64_t g_batch_stride, const void* X,
                          int64_t ldx, int64_t x_batch_stride, int batch, int rows, int m_total,
                          int n_total, int taps, const int* shifts, float* out, int64_t ld_out,
                          float scale, int ksplit, void* stream);

#ifdef __cplusplus
}
#endif
#endif
