// Strong augmentation of the input pipeline on the device: the torchvision / Pillow chain of
// pt/data/detection_utils.py:38-60 (ColorJitter -> RandomGrayscale -> GaussianBlur -> Solarize, applied to a PIL
// image at pt/data/dataset_mapper.py:159-164) over uint8 CHW images, BIT-EXACT with Pillow 12.2 / torchvision 0.26:
//   * Image.blend (ImageEnhance Brightness / Contrast / Color): single-precision d + a * (i - d), no FMA
//     contraction, truncation (clip first when a is outside [0, 1])           (libImaging/Blend.c)
//   * Image.convert L / HSV / RGB: 16.16 fixed-point luma; rgb2hsv / hsv2rgb with the float / double mix of
//     libImaging/Convert.c (pinned exhaustively over all 2^24 inputs by the oracle's CPU test)
//   * ImageFilter.GaussianBlur(radius): 3 box-blur passes per axis with 8.24 fixed-point weights and edge extension,
//     every pass rounded to uint8                                               (libImaging/BoxBlur.c)
//   * ImageOps.solarize(threshold 128).
// HBM-bound byte work: one pointwise kernel per contiguous run of per-pixel operations (the contrast op needs the mean
// luma of the image AS IT IS at that point of the chain, accumulated by the preceding run), one kernel for the three
// horizontal blur passes (a row lives in shared memory) and one for the three vertical passes (a 32-column strip).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

enum AugOp { OP_BRIGHTNESS = 0, OP_CONTRAST = 1, OP_SATURATION = 2, OP_HUE = 3, OP_GRAY = 4, OP_SOLARIZE = 5 };

struct AugChain {
  int n;
  int op[8];
  float f[8];  // blend factor (brightness / contrast / saturation); hue: the uint8 shift stored as float
};

__device__ __forceinline__ int luma(int r, int g, int b) {
  return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16;
}

// Image.blend of one channel: degenerate d, image i, factor a
__device__ __forceinline__ int blend1(float d, float i, float a, bool inside) {
  const float t = __fadd_rn(d, __fmul_rn(a, __fsub_rn(i, d)));
  if (inside) return static_cast<int>(static_cast<uint8_t>(t));
  if (t <= 0.f) return 0;
  if (t >= 255.f) return 255;
  return static_cast<int>(static_cast<uint8_t>(t));
}

__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) {
    uh = 0;
    us = 0;
    return;
  }
  const float cr = static_cast<float>(maxc - minc);
  const float s = __fdiv_rn(cr, static_cast<float>(maxc));
  const double rc = static_cast<double>(__fdiv_rn(static_cast<float>(maxc - r), cr));
  const double gc = static_cast<double>(__fdiv_rn(static_cast<float>(maxc - g), cr));
  const double bc = static_cast<double>(__fdiv_rn(static_cast<float>(maxc - b), cr));
  float h;
  if (r == maxc)
    h = static_cast<float>(__dsub_rn(bc, gc));
  else if (g == maxc)
    h = static_cast<float>(__dsub_rn(__dadd_rn(2.0, rc), bc));
  else
    h = static_cast<float>(__dsub_rn(__dadd_rn(4.0, gc), rc));
  const double x = __dadd_rn(__ddiv_rn(static_cast<double>(h), 6.0), 1.0);
  const float hh = static_cast<float>(x - floor(x));  // fmod(x, 1.0) for x > 0 (exact)
  int ih = static_cast<int>(__dmul_rn(static_cast<double>(hh), 255.0));
  int is = static_cast<int>(__dmul_rn(static_cast<double>(s), 255.0));
  uh = min(max(ih, 0), 255);
  us = min(max(is, 0), 255);
}

__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) {
    r = g = b = v;
    return;
  }
  const double hd = static_cast<double>(h), vd = static_cast<double>(v);
  const double h6 = __ddiv_rn(__dmul_rn(hd, 6.0), 255.0);
  const double fi = floor(h6);
  const double f = static_cast<double>(static_cast<float>(__dsub_rn(h6, fi)));
  const double fs = static_cast<double>(static_cast<float>(__ddiv_rn(static_cast<double>(s), 255.0)));
  const int p = min(max(static_cast<int>(round(__dmul_rn(vd, __dsub_rn(1.0, fs)))), 0), 255);
  const int q = min(max(static_cast<int>(round(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fs, f))))), 0), 255);
  const int t = min(max(static_cast<int>(round(__dmul_rn(vd, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f)))))), 0), 255);
  switch (static_cast<int>(fi) % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// Applies chain.op[0..n) to every pixel of a planar uint8 image. gray_sum_in: sum of the luma of the INPUT image
// (consumed by a contrast op, which must be first in its run); gray_sum_out (may be null): receives the sum of the
// luma of the OUTPUT (for the contrast op that opens the next run).
__global__ void aug_pointwise_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t npix,
                                     AugChain chain, const unsigned long long* __restrict__ gray_sum_in,
                                     unsigned long long* __restrict__ gray_sum_out) {
  int mean = 0;
  if (chain.n > 0 && chain.op[0] == OP_CONTRAST) {
    // int(ImageStat.Stat(L).mean[0] + 0.5): Python float (double) arithmetic
    mean = static_cast<int>(__dadd_rn(__ddiv_rn(static_cast<double>(gray_sum_in[0]), static_cast<double>(npix)), 0.5));
  }
  unsigned long long lsum = 0ull;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int r = in[i], g = in[npix + i], b = in[2 * npix + i];
    for (int k = 0; k < chain.n; ++k) {
      const float a = chain.f[k];
      const bool inside = a >= 0.f && a <= 1.f;
      switch (chain.op[k]) {
        case OP_BRIGHTNESS:
          r = blend1(0.f, static_cast<float>(r), a, inside);
          g = blend1(0.f, static_cast<float>(g), a, inside);
          b = blend1(0.f, static_cast<float>(b), a, inside);
          break;
        case OP_CONTRAST: {
          const float d = static_cast<float>(mean);
          r = blend1(d, static_cast<float>(r), a, inside);
          g = blend1(d, static_cast<float>(g), a, inside);
          b = blend1(d, static_cast<float>(b), a, inside);
          break;
        }
        case OP_SATURATION: {
          const float d = static_cast<float>(luma(r, g, b));
          r = blend1(d, static_cast<float>(r), a, inside);
          g = blend1(d, static_cast<float>(g), a, inside);
          b = blend1(d, static_cast<float>(b), a, inside);
          break;
        }
        case OP_HUE: {
          int h, s, v;
          rgb2hsv(r, g, b, h, s, v);
          h = (h + static_cast<int>(a)) & 255;  // uint8 wrap-around
          hsv2rgb(h, s, v, r, g, b);
          break;
        }
        case OP_GRAY:
          r = g = b = luma(r, g, b);
          break;
        default:  // OP_SOLARIZE, threshold 128
          r = r < 128 ? r : 255 - r;
          g = g < 128 ? g : 255 - g;
          b = b < 128 ? b : 255 - b;
          break;
      }
    }
    out[i] = static_cast<uint8_t>(r);
    out[npix + i] = static_cast<uint8_t>(g);
    out[2 * npix + i] = static_cast<uint8_t>(b);
    if (gray_sum_out != nullptr) lsum += static_cast<unsigned long long>(luma(r, g, b));
  }
  if (gray_sum_out != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0 && lsum) atomicAdd(gray_sum_out, lsum);
  }
}

// One BoxBlur.c pass over a line held in shared memory: out[x] = (ww * sum_{|d|<=radius} in[clamp(x+d)]
//   + fw * (in[clamp(x-radius-1)] + in[clamp(x+radius+1)]) + 2^23) >> 24
__device__ __forceinline__ uint8_t box_tap(const uint8_t* line, int n, int x, int radius, uint32_t ww, uint32_t fw) {
  uint32_t acc = 0;
  for (int d = -radius; d <= radius; ++d) acc += line[min(max(x + d, 0), n - 1)];
  const uint32_t far = static_cast<uint32_t>(line[min(max(x - radius - 1, 0), n - 1)]) + line[min(max(x + radius + 1, 0), n - 1)];
  const unsigned long long bulk = static_cast<unsigned long long>(acc) * ww + static_cast<unsigned long long>(far) * fw;
  return static_cast<uint8_t>((bulk + (1ull << 23)) >> 24);
}

// `passes` horizontal passes of one image row (one CTA per (plane, row)); two ping-pong line buffers in shared memory
__global__ void aug_boxblur_rows_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int radius,
                                        uint32_t ww, uint32_t fw, int passes) {
  extern __shared__ uint8_t lines[];
  uint8_t* a = lines;
  uint8_t* b = lines + ((W + 15) & ~15);
  const int64_t off = static_cast<int64_t>(blockIdx.x) * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) a[x] = in[off + x];
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    for (int x = threadIdx.x; x < W; x += blockDim.x) b[x] = box_tap(a, W, x, radius, ww, fw);
    __syncthreads();
    uint8_t* t = a;
    a = b;
    b = t;
  }
  for (int x = threadIdx.x; x < W; x += blockDim.x) out[off + x] = a[x];
}

// `passes` vertical passes of a strip of 32 columns of one plane: thread (cx, ry) walks rows ry, ry + 8, ...
__global__ void aug_boxblur_cols_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W,
                                        int radius, uint32_t ww, uint32_t fw, int passes) {
  extern __shared__ uint8_t strip[];  // 2 x [32][Hp] (column-major: a column is a contiguous line)
  const int Hp = (H + 15) & ~15;
  uint8_t* a = strip;
  uint8_t* b = strip + 32 * Hp;
  const int plane = blockIdx.y, x0 = blockIdx.x * 32;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5, nry = blockDim.x >> 5;
  const int64_t base = static_cast<int64_t>(plane) * H * W;
  const bool live = x0 + cx < W;
  for (int y = ry; y < H; y += nry) a[cx * Hp + y] = live ? in[base + static_cast<int64_t>(y) * W + x0 + cx] : 0;
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    for (int y = ry; y < H; y += nry) b[cx * Hp + y] = box_tap(a + cx * Hp, H, y, radius, ww, fw);
    __syncthreads();
    uint8_t* t = a;
    a = b;
    b = t;
  }
  if (live)
    for (int y = ry; y < H; y += nry) out[base + static_cast<int64_t>(y) * W + x0 + cx] = a[cx * Hp + y];
}

// One pass of Pillow's ImagingResample (libImaging/Resample.c, 8 bits per channel) along x or y of a planar image:
// out[o] = clip8((2^21 + sum_j in[xmin_o + j] * kk[o][j]) >> 22) with the per-output tap window (bounds) and 22-bit
// fixed-point coefficients precomputed on the host exactly as precompute_coeffs / normalize_coeffs_8bpc do.
__global__ void aug_resample_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int planes, int in_h,
                                    int in_w, int out_h, int out_w, int along_x, const int* __restrict__ bounds,
                                    const int* __restrict__ kk, int ksize) {
  const int64_t total = static_cast<int64_t>(planes) * out_h * out_w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % out_w);
    const int64_t r = i / out_w;
    const int y = static_cast<int>(r % out_h);
    const int pl = static_cast<int>(r / out_h);
    const int o = along_x ? x : y;
    const int xmin = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = kk + static_cast<int64_t>(o) * ksize;
    const uint8_t* src = in + static_cast<int64_t>(pl) * in_h * in_w + (along_x ? static_cast<int64_t>(y) * in_w + xmin
                                                                               : static_cast<int64_t>(xmin) * in_w + x);
    const int stride = along_x ? 1 : in_w;
    int ss = 1 << 21;
    for (int j = 0; j < cnt; ++j) ss += static_cast<int>(src[static_cast<int64_t>(j) * stride]) * k[j];
    ss >>= 22;
    out[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

__global__ void aug_hflip_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t rows, int w) {
  const int64_t total = rows * w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    out[i] = in[i - x + (w - 1 - x)];
  }
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ptb200_aug_pointwise_u8(const uint8_t* in, uint8_t* out, int h, int w, int n_ops, const int* ops_host,
                                       const float* factors_host, const unsigned long long* gray_sum_in,
                                       unsigned long long* gray_sum_out, void* stream) {
  if (n_ops < 0 || n_ops > 8) return 1601;
  AugChain c;
  c.n = n_ops;
  for (int i = 0; i < 8; ++i) {
    c.op[i] = i < n_ops ? ops_host[i] : 0;
    c.f[i] = i < n_ops ? factors_host[i] : 0.f;
    if (i < n_ops && (c.op[i] < 0 || c.op[i] > OP_SOLARIZE)) return 1602;
    if (i > 0 && i < n_ops && c.op[i] == OP_CONTRAST) return 1603;  // contrast must open its run (it needs the mean)
  }
  if (n_ops > 0 && c.op[0] == OP_CONTRAST && gray_sum_in == nullptr) return 1604;
  if (gray_sum_out != nullptr) cudaMemsetAsync(gray_sum_out, 0, sizeof(unsigned long long), STREAM);
  const int64_t npix = static_cast<int64_t>(h) * w;
  aug_pointwise_kernel<<<grid_for(npix), 256, 0, STREAM>>>(in, out, npix, c, gray_sum_in, gray_sum_out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_aug_boxblur_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int planes, int h, int w,
                                     int radius, int ww, int fw, int passes, void* stream) {
  if (radius < 0 || passes < 1) return 1605;
  const size_t smem_r = 2 * static_cast<size_t>((w + 15) & ~15);
  const size_t smem_c = 2 * 32 * static_cast<size_t>((h + 15) & ~15);
  if (smem_r > 200 * 1024 || smem_c > 200 * 1024) return 1606;
  static size_t conf_r = 48 * 1024, conf_c = 48 * 1024;
  if (smem_r > conf_r) {
    if (cudaFuncSetAttribute(aug_boxblur_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 1607;
    conf_r = 200 * 1024;
  }
  if (smem_c > conf_c) {
    if (cudaFuncSetAttribute(aug_boxblur_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 1607;
    conf_c = 200 * 1024;
  }
  aug_boxblur_rows_kernel<<<planes * h, 256, smem_r, STREAM>>>(in, tmp, w, radius, static_cast<uint32_t>(ww),
                                                              static_cast<uint32_t>(fw), passes);
  dim3 grid((w + 31) / 32, planes);
  aug_boxblur_cols_kernel<<<grid, 256, smem_c, STREAM>>>(tmp, out, h, w, radius, static_cast<uint32_t>(ww),
                                                        static_cast<uint32_t>(fw), passes);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_aug_resample_u8(const uint8_t* in, uint8_t* out, int planes, int in_h, int in_w, int out_h,
                                      int out_w, int along_x, const int* bounds_dev, const int* coeffs_dev, int ksize,
                                      void* stream) {
  if ((along_x && in_h != out_h) || (!along_x && in_w != out_w) || ksize < 1) return 1608;
  const int64_t total = static_cast<int64_t>(planes) * out_h * out_w;
  aug_resample_kernel<<<grid_for(total), 256, 0, STREAM>>>(in, out, planes, in_h, in_w, out_h, out_w, along_x,
                                                          bounds_dev, coeffs_dev, ksize);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_aug_hflip_u8(const uint8_t* in, uint8_t* out, int planes, int h, int w, void* stream) {
  const int64_t rows = static_cast<int64_t>(planes) * h;
  aug_hflip_kernel<<<grid_for(rows * w), 256, 0, STREAM>>>(in, out, rows, w);
  return static_cast<int>(cudaGetLastError());
}
