// Internal (C++) interface of the tcgen05 implicit-GEMM kernels. The public C ABI is include/ptb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptb {

enum GemmEpilogue {
  EPI_BIAS_RELU_F16 = 0,  // D = relu(acc + bias) -> fp16 via TMA store (pad column forced to 0)
  EPI_BIAS_F16 = 1,       // D = acc + bias -> fp16
  EPI_F32_SPLIT = 2,      // acc + bias -> fp32, columns [0,split) to d0, [split,n_valid) to d1
  EPI_MASK_F16 = 3,       // D = (aux > 0) ? acc : 0 -> fp16 (ReLU backward fused into dgrad)
  EPI_ATOMIC_F32 = 4,     // split-K partial: d0[row][n] += acc (fp32 red.add), bias/activation applied later
  // split-fp16 ("f16x3") parity precision: x = alpha*acc + bias (ReLU for 5) is written as the K-concatenated
  // triple [hi | lo | hi] (hi = fp16(x), lo = fp16(x - hi)) into D of width 3*n_total, so that the next
  // GEMM against [Wh | Wh | Wl] evaluates hi*Wh + lo*Wh + hi*Wl with fp32 accumulation
  EPI_SPLIT3_RELU_F16 = 5,
  EPI_SPLIT3_F16 = 6,
  // gemm_tn_x3.cu (in-kernel fp32 promotion) only:
  EPI_SPLIT3_MASK_F16 = 7,  // (aux > 0 ? alpha*acc : 0) -> triple; aux = forward activation triple (hi + lo)
  EPI_F32_STORE = 8,        // alpha*acc + bias -> fp32 d0[row][n] (pad-column rows written as zero)
};

struct GemmTnParams {
  int batch, rows, k_per_tap, taps;
  int shifts[9];
  int n_total, bn;
  int w_valid, wp;
  int epi;
  int stages;
  int ksplit;
  int b_resident;  // ROWWIN only: whole 3x3 filter (9 x bn x 64) stays in shared memory
  int staging_bufs;  // 1 or 2 epilogue staging buffers of 16 KB
  float alpha;     // accumulator scale applied before the bias (power-of-two weight scaling of the f16x3 mode)
  const float* bias;
  int n_bias;
  float* d0;
  int ld0;
  float* d1;
  int ld1;
  int split;
  int n_valid;
  const int* seg_counts;  // optional: rows are [segments][seg_cap], only the first seg_counts[s] rows of a
  int seg_cap;            // segment are live; 128-row tiles without any live row are skipped entirely
  int chunk = 0;          // gemm_tn_x3.cu: k-iterations (of 64) per fp32 promotion of the TMEM partial sums
  int hi_share = 0;       // gemm_tn_x3.cu row-window mode: one `hi` A window serves the Wh AND the Wl products
  int rw_ny = 3, rw_nx = 3;  // gemm_tn_x3.cu pair mode: 3 x 3 taps (conv) or 1 x 1 (plain GEMM through the same rings)
};

struct GemmTnArgs {
  // A: fp16 [batch][rows][lda] (k_per_tap <= lda), row shift per tap
  const void* A;
  int batch, rows, k_per_tap;
  int64_t lda, a_batch_stride;  // in elements
  int taps;
  int shifts[9];
  // B: fp16 [n_total][taps * k_per_tap]
  const void* B;
  int n_total, bn;
  // epilogue
  int epi;
  const float* bias;
  int n_bias;
  void* D;  // fp16 [batch][rows][ldd]
  int64_t ldd, d_batch_stride;
  const void* aux;  // fp16, same geometry as D (EPI_MASK_F16)
  int w_valid, wp;  // wp > 0: rows with (row % wp) >= w_valid are written as zero
  float* d0;
  int ld0;
  float* d1;
  int ld1;
  int split, n_valid;
  int max_ctas;  // 0 = one per SM
  int ksplit;    // > 1 only with EPI_ATOMIC_F32
  const int* seg_counts;
  int seg_cap;
  float alpha = 1.0f;
};

int gemm_tn_launch(const GemmTnArgs& a, cudaStream_t stream);
// f16x3 GEMM with in-kernel fp32 promotion (gemm_tn_x3.cu); a.D / a.aux are [batch][rows][3*n_total] triples
int gemm_tn_promote_launch(const GemmTnArgs& a, int chunk, cudaStream_t stream);

// dW[m][t*n_total + n] += scale * sum_{b,p} G[b][p][m] * X[b][p + shifts[t]][n]   (fp32 atomics)
int gemm_wgrad_launch(const void* G, int64_t ldg, int64_t g_batch_stride, const void* X, int64_t ldx,
                      int64_t x_batch_stride, int batch, int rows, int m_total, int n_total, int taps,
                      const int* shifts, float* out, int64_t ld_out, float scale, int ksplit, float* bias_out,
                      const int* seg_counts, int seg_cap, cudaStream_t stream, bool x3 = false);

}  // namespace ptb
