// C ABI: dense contractions (see include/ptb200.h for the contract of every entry point).
#include "gemm_tn.h"
#include "../../include/ptb200.h"

using namespace ptb;

static int gemm_tn_capi(const void* A, int batch, int rows, int k_per_tap, int64_t lda,
                        int64_t a_batch_stride, int taps, const int* shifts, const void* B,
                        int n_total, int bn, int epi, const float* bias, int n_bias, void* D,
                        int64_t ldd, int64_t d_batch_stride, const void* aux, int w_valid,
                        int wp, float* d0, int ld0, float* d1, int ld1, int split,
                        int n_valid, int max_ctas, int ksplit, const int* seg_counts, int seg_cap,
                        float alpha, void* stream) {
  GemmTnArgs a;
  a.A = A;
  a.batch = batch;
  a.rows = rows;
  a.k_per_tap = k_per_tap;
  a.lda = lda;
  a.a_batch_stride = a_batch_stride;
  a.taps = taps;
  for (int i = 0; i < 9; ++i) a.shifts[i] = (shifts != nullptr && i < taps) ? shifts[i] : 0;
  a.B = B;
  a.n_total = n_total;
  a.bn = bn;
  a.epi = epi;
  a.bias = bias;
  a.n_bias = n_bias;
  a.D = D;
  a.ldd = ldd;
  a.d_batch_stride = d_batch_stride;
  a.aux = aux;
  a.w_valid = w_valid;
  a.wp = wp;
  a.d0 = d0;
  a.ld0 = ld0;
  a.d1 = d1;
  a.ld1 = ld1;
  a.split = split;
  a.n_valid = n_valid;
  a.max_ctas = max_ctas;
  a.ksplit = ksplit;
  a.seg_counts = seg_counts;
  a.seg_cap = seg_cap;
  a.alpha = alpha;
  return gemm_tn_launch(a, static_cast<cudaStream_t>(stream));
}

extern "C" int ptb200_gemm_tn_f16(const void* A, int batch, int rows, int k_per_tap, int64_t lda,
                                  int64_t a_batch_stride, int taps, const int* shifts, const void* B,
                                  int n_total, int bn, int epi, const float* bias, int n_bias, void* D,
                                  int64_t ldd, int64_t d_batch_stride, const void* aux, int w_valid,
                                  int wp, float* d0, int ld0, float* d1, int ld1, int split,
                                  int n_valid, int max_ctas, int ksplit, const int* seg_counts, int seg_cap,
                                  void* stream) {
  if (epi == EPI_SPLIT3_RELU_F16 || epi == EPI_SPLIT3_F16) return 1020;  // f16x3 epilogues: use ptb200_gemm_tn_f16x3
  return gemm_tn_capi(A, batch, rows, k_per_tap, lda, a_batch_stride, taps, shifts, B, n_total, bn, epi, bias,
                      n_bias, D, ldd, d_batch_stride, aux, w_valid, wp, d0, ld0, d1, ld1, split, n_valid, max_ctas,
                      ksplit, seg_counts, seg_cap, 1.0f, stream);
}

extern "C" int ptb200_gemm_tn_f16x3(const void* A3, int batch, int rows, int k3_per_tap, int64_t lda,
                                    int64_t a_batch_stride, int taps, const int* shifts, const void* B3,
                                    int n_total, int bn, int epi, const float* bias, int n_bias, void* D3,
                                    int64_t ldd, int64_t d_batch_stride, int w_valid, int wp, float* d0, int ld0,
                                    float* d1, int ld1, int split, int n_valid, int max_ctas, int ksplit,
                                    const int* seg_counts, int seg_cap, float alpha, const void* aux3, int chunk,
                                    void* stream) {
  if (k3_per_tap % 3 != 0) return 1022;
  if (epi == EPI_F32_SPLIT) {
    // narrow fp32 heads (N = 96 / 128, K <= 3 * 1024: at most 192 chained MMAs): one accumulation chain
    return gemm_tn_capi(A3, batch, rows, k3_per_tap, lda, a_batch_stride, taps, shifts, B3, n_total, bn, epi, bias,
                        n_bias, D3, ldd, d_batch_stride, nullptr, w_valid, wp, d0, ld0, d1, ld1, split, n_valid,
                        max_ctas, ksplit, seg_counts, seg_cap, alpha, stream);
  }
  if (epi != EPI_SPLIT3_RELU_F16 && epi != EPI_SPLIT3_F16 && epi != EPI_SPLIT3_MASK_F16 && epi != EPI_F32_STORE &&
      epi != EPI_ATOMIC_F32)
    return 1021;
  GemmTnArgs a;
  a.A = A3;
  a.batch = batch;
  a.rows = rows;
  a.k_per_tap = k3_per_tap;
  a.lda = lda;
  a.a_batch_stride = a_batch_stride;
  a.taps = taps;
  for (int i = 0; i < 9; ++i) a.shifts[i] = (shifts != nullptr && i < taps) ? shifts[i] : 0;
  a.B = B3;
  a.n_total = n_total;
  a.bn = bn;
  a.epi = epi;
  a.bias = bias;
  a.n_bias = n_bias;
  a.D = D3;
  a.ldd = ldd;
  a.d_batch_stride = d_batch_stride;
  a.aux = aux3;
  a.w_valid = w_valid;
  a.wp = wp;
  a.d0 = d0;
  a.ld0 = ld0;
  a.d1 = nullptr;
  a.ld1 = 0;
  a.split = 0;
  a.n_valid = n_total;
  a.max_ctas = max_ctas;
  a.ksplit = ksplit;
  a.seg_counts = seg_counts;
  a.seg_cap = seg_cap;
  a.alpha = alpha;
  return gemm_tn_promote_launch(a, chunk, static_cast<cudaStream_t>(stream));
}

extern "C" int ptb200_gemm_wgrad_f16(const void* G, int64_t ldg, int64_t g_batch_stride, const void* X,
                                     int64_t ldx, int64_t x_batch_stride, int batch, int rows,
                                     int m_total, int n_total, int taps, const int* shifts, float* out,
                                     int64_t ld_out, float scale, int ksplit, float* bias_out,
                                     const int* seg_counts, int seg_cap, void* stream) {
  return gemm_wgrad_launch(G, ldg, g_batch_stride, X, ldx, x_batch_stride, batch, rows, m_total,
                           n_total, taps, shifts, out, ld_out, scale, ksplit, bias_out, seg_counts, seg_cap,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int ptb200_gemm_wgrad_f16x3(const void* G3, int64_t ldg, int64_t g_batch_stride, const void* X3,
                                       int64_t ldx, int64_t x_batch_stride, int batch, int rows, int m_total,
                                       int n_total, int taps, const int* shifts, float* out, int64_t ld_out,
                                       float scale, int ksplit, float* bias_out, const int* seg_counts, int seg_cap,
                                       void* stream) {
  return gemm_wgrad_launch(G3, ldg, g_batch_stride, X3, ldx, x_batch_stride, batch, rows, m_total, n_total, taps,
                           shifts, out, ld_out, scale, ksplit, bias_out, seg_counts, seg_cap,
                           static_cast<cudaStream_t>(stream), true);
}
