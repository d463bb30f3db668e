// Helpers of the split-fp16 ("f16x3") fp32-equivalent precision, forward and backward (see ptb200_gemm_tn_f16x3 in
// include/ptb200.h). An fp32 value x is carried as hi = fp16(x), lo = fp16(x - hi); activation rows are
// the K-concatenation [hi | lo | hi] (3C wide), weight rows [Wh | Wh | Wl], so the tensor-core main loop
// needs no change. Everything here is HBM-bound streaming work with 16-byte accesses, plus the first
// VGG conv (K = 27) evaluated in fp32 on the CUDA cores.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t n, int per_block = kThreads) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return static_cast<int>(g);
}

__device__ __forceinline__ void split_hl(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}

// src fp32 [rows][k] -> dst fp16 [rows][3k]; order 0: [hi | lo | hi], order 1: [hi | hi | lo]
__global__ void split3_pack_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t rows, int k,
                                   float scale, int order) {
  const int64_t total = rows * k;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / k;
    const int c = static_cast<int>(i - r * k);
    __half h, l;
    split_hl(src[i] * scale, h, l);
    __half* d = dst + r * 3 * k + c;
    d[0] = h;
    d[k] = order == 0 ? l : h;
    d[2 * k] = order == 0 ? h : l;
  }
}

// same, 8 consecutive elements per thread (k % 8 == 0): two 16-byte loads, three 16-byte stores. The scalar kernel
// above (2-byte stores, one division per element) ran the per-step weight re-packing at 1.5 TB/s.
__global__ void split3_pack_vec8_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t rows, int k,
                                        float scale, int order) {
  const int k8 = k >> 3;
  const int64_t total = rows * k8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / k8;
    const int c = static_cast<int>(i - r * k8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(src + r * k + c);
    const float4 b = *reinterpret_cast<const float4*>(src + r * k + c + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_hl(v[e] * scale, h[e], l[e]);
    __half* d = dst + r * 3 * k + c;
    const uint4 hv = *reinterpret_cast<const uint4*>(h), lv = *reinterpret_cast<const uint4*>(l);
    *reinterpret_cast<uint4*>(d) = hv;
    *reinterpret_cast<uint4*>(d + k) = order == 0 ? lv : hv;
    *reinterpret_cast<uint4*>(d + 2 * k) = order == 0 ? hv : lv;
  }
}

__global__ void split3_unpack_kernel(const __half* __restrict__ src, float* __restrict__ dst, int64_t rows, int k) {
  const int64_t total = rows * k;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / k;
    const int c = static_cast<int>(i - r * k);
    const __half* s = src + r * 3 * k + c;
    dst[i] = __half2float(s[0]) + __half2float(s[k]);
  }
}

// 2x2 / stride 2 max pool over triples: the arg-max is taken over hi + lo (exact in fp32 for |lo| <= ulp(hi)/2),
// its (hi, lo) pair is copied. One thread = (output pixel, 8 channels). Pad column written as zero.
__global__ void maxpool2x2_x3_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W,
                                     int C) {
  const int Wp = W + 1, Ho = H / 2, Wo = W / 2, Wop = Wo + 1, C8 = C / 8;
  const int64_t LD = 3 * static_cast<int64_t>(C);
  const int64_t total = static_cast<int64_t>(N) * Ho * Wop * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    int64_t r = i / C8;
    const int xo = static_cast<int>(r % Wop);
    r /= Wop;
    const int yo = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    __align__(16) __half oh[8], ol[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) oh[e] = ol[e] = __float2half(0.f);
    if (xo < Wo) {
      float best[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __half* base =
            in + ((static_cast<int64_t>(n) * H + 2 * yo + (q >> 1)) * Wp + 2 * xo + (q & 1)) * LD + c8 * 8;
        const uint4 vh = *reinterpret_cast<const uint4*>(base);
        const uint4 vl = *reinterpret_cast<const uint4*>(base + C);
        const __half* hh = reinterpret_cast<const __half*>(&vh);
        const __half* ll = reinterpret_cast<const __half*>(&vl);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = __half2float(hh[e]) + __half2float(ll[e]);
          if (q == 0 || v > best[e]) {
            best[e] = v;
            oh[e] = hh[e];
            ol[e] = ll[e];
          }
        }
      }
    }
    __half* o = out + ((static_cast<int64_t>(n) * Ho + yo) * Wop + xo) * LD + c8 * 8;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(oh);
    *reinterpret_cast<uint4*>(o + C) = *reinterpret_cast<const uint4*>(ol);
    *reinterpret_cast<uint4*>(o + 2 * C) = *reinterpret_cast<const uint4*>(oh);
  }
}

// Pre-processing + first VGG conv in fp32: one thread = one pixel x 16 output channels (4 threads per pixel);
// the 27 normalised inputs come from the L1-resident neighbourhood, the 64 x 27 filter sits in shared memory.
__global__ void __launch_bounds__(256)
conv1_u8_x3_kernel(const uint8_t* __restrict__ img, const int* __restrict__ hw, int N, int Hmax, int Wmax,
                   int64_t img_stride, float m0, float m1, float m2, float is0, float is1, float is2,
                   const float* __restrict__ w /* [64][27] */, const float* __restrict__ bias,
                   __half* __restrict__ out /* [N][Hmax*Wp][192] */) {
  __shared__ float ws[27][64];
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) ws[i % 27][i / 27] = w[i];
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int Wp = Wmax + 1;
  const int64_t total = static_cast<int64_t>(N) * Hmax * Wp * 4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int part = static_cast<int>(i & 3);
    const int64_t pix = i >> 2;
    const int x = static_cast<int>(pix % Wp);
    const int64_t ny = pix / Wp;
    const int y = static_cast<int>(ny % Hmax);
    const int n = static_cast<int>(ny / Hmax);
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    if (x < Wmax) {
      const int h = hw[2 * n], wd = hw[2 * n + 1];
      const uint8_t* ib = img + n * img_stride;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        const bool ok = yy >= 0 && yy < h && xx >= 0 && xx < wd;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ok) {
            const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
            const float istd = c == 0 ? is0 : (c == 1 ? is1 : is2);
            v = (static_cast<float>(__ldg(ib + (static_cast<int64_t>(c) * h + yy) * wd + xx)) - mean) * istd;
          }
          const float* wr = &ws[t * 3 + c][part * 16];
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[e] = fmaf(v, wr[e], acc[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = fmaxf(acc[e] + bs[part * 16 + e], 0.f);
    }
    __align__(16) __half oh[16], ol[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) split_hl(acc[e], oh[e], ol[e]);
    __half* o = out + pix * 192 + part * 16;
    const uint4* ph = reinterpret_cast<const uint4*>(oh);
    const uint4* pl = reinterpret_cast<const uint4*>(ol);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      *reinterpret_cast<uint4*>(o + q * 8) = ph[q];
      *reinterpret_cast<uint4*>(o + 64 + q * 8) = pl[q];
      *reinterpret_cast<uint4*>(o + 128 + q * 8) = ph[q];
    }
  }
}

// Finishes a split-K f16x3 GEMM: `slices` fp32 partial sums [slices][rows][n] (one per K split, written by
// EPI_F32_STORE) are added IN A FIXED ORDER -- the forward stays bit-reproducible run to run, which the discrete
// proposal / pseudo-label decisions downstream need -- then x = act(alpha * sum + bias) -> triple.
// wp > 0: rows with (row % wp) >= w_valid (the pad column of the flat activation layout) are written as zero.
__global__ void bias_act_split3_kernel(const float* __restrict__ in, int slices, const float* __restrict__ bias,
                                       int relu, float alpha, int64_t rows, int n, int wp, int w_valid,
                                       __half* __restrict__ out) {
  const int n4 = n / 4;
  const int64_t total4 = rows * n4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / n4;
    const int c = static_cast<int>(i - r * n4) * 4;
    float4 v4 = reinterpret_cast<const float4*>(in)[i];
    for (int sl = 1; sl < slices; ++sl) {
      const float4 w4 = reinterpret_cast<const float4*>(in)[sl * total4 + i];
      v4.x += w4.x;
      v4.y += w4.y;
      v4.z += w4.z;
      v4.w += w4.w;
    }
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
    const bool live = wp <= 0 || static_cast<int>(r % wp) < w_valid;
    __align__(8) __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x = v[e] * alpha + (bias != nullptr ? bias[c + e] : 0.f);
      if (relu) x = fmaxf(x, 0.f);
      if (!live) x = 0.f;
      split_hl(x, h[e], l[e]);
    }
    __half* o = out + r * 3 * n + c;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(o + n) = *reinterpret_cast<const uint2*>(l);
    *reinterpret_cast<uint2*>(o + 2 * n) = *reinterpret_cast<const uint2*>(h);
  }
}


// ------------------------------------------------------------------------------------------ backward helpers
// Unit gradients of two loss terms -> output-gradient triples [rows][3*ld] of the head GEMM (f16x3 counterpart of
// pack_grad2_kernel): x = lscale * (c < n0 ? g0*d0[r][c] : g1*d1[r][c-n0]), zero padded to ld.
__global__ void pack_grad2_x3_kernel(const float* __restrict__ d0, int n0, const float* __restrict__ d1, int n1,
                                     const float* __restrict__ g0, const float* __restrict__ g1, float lscale,
                                     int64_t rows, int ld, __half* __restrict__ out) {
  const int64_t total = rows * ld;
  const float w0 = g0 != nullptr ? g0[0] * lscale : lscale;
  const float w1 = g1 != nullptr ? g1[0] * lscale : lscale;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ld);
    const int64_t r = i / ld;
    float v = 0.f;
    if (c < n0)
      v = w0 * d0[r * n0 + c];
    else if (c < n0 + n1)
      v = w1 * d1[r * n1 + (c - n0)];
    __half h, l;
    split_hl(v, h, l);
    __half* o = out + r * 3 * ld + c;
    o[0] = h;
    o[ld] = l;
    o[2 * ld] = h;
  }
}

// Backward of (ReLU -> 2x2 max pool) in the f16x3 precision: X = pre-pool activation triples, dP = fp32 gradient
// w.r.t. the pooled map [N][Ho*(Wo+1)][C] -> dZ triples [N][H*(W+1)][3C]. dZ[pos] = dP[pooled] if pos is the
// first arg-max (row-major scan, as ATen) of its window over hi + lo and X[pos] > 0, else 0.
__global__ void maxpool2x2_relu_bwd_x3_kernel(const __half* __restrict__ x, const float* __restrict__ dp,
                                              __half* __restrict__ dz, int N, int H, int W, int C) {
  const int Wp = W + 1, Ho = H / 2, Wo = W / 2, Wop = Wo + 1, C8 = C / 8;
  const int Hc = (H + 1) / 2, Wc = (Wp + 1) / 2;
  const int64_t LD = 3 * static_cast<int64_t>(C);
  const int64_t total = static_cast<int64_t>(N) * Hc * Wc * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % C8);
    int64_t r = i / C8;
    const int xo = static_cast<int>(r % Wc);
    r /= Wc;
    const int yo = static_cast<int>(r % Hc);
    const int n = static_cast<int>(r / Hc);
    const bool covered = (yo < Ho) && (xo < Wo);
    float g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = 0.f;
    int best[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) best[e] = -1;
    if (covered) {
      const float* gp = dp + ((static_cast<int64_t>(n) * Ho + yo) * Wop + xo) * C + c8 * 8;
      const float4 g0 = *reinterpret_cast<const float4*>(gp), g1 = *reinterpret_cast<const float4*>(gp + 4);
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w;
      g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
      float bv[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half* base = x + ((static_cast<int64_t>(n) * H + 2 * yo + (k >> 1)) * Wp + 2 * xo + (k & 1)) * LD + c8 * 8;
        const uint4 vh = *reinterpret_cast<const uint4*>(base);
        const uint4 vl = *reinterpret_cast<const uint4*>(base + C);
        const __half* hh = reinterpret_cast<const __half*>(&vh);
        const __half* ll = reinterpret_cast<const __half*>(&vl);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = __half2float(hh[e]) + __half2float(ll[e]);
          if (k == 0 || v > bv[e]) {
            bv[e] = v;
            best[e] = k;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (!(bv[e] > 0.f)) best[e] = -1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = 2 * yo + (k >> 1), xx = 2 * xo + (k & 1);
      if (yy >= H || xx >= Wp) continue;
      __align__(16) __half oh[8], ol[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_hl(best[e] == k ? g[e] : 0.f, oh[e], ol[e]);
      __half* o = dz + ((static_cast<int64_t>(n) * H + yy) * Wp + xx) * LD + c8 * 8;
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(oh);
      *reinterpret_cast<uint4*>(o + C) = *reinterpret_cast<const uint4*>(ol);
      *reinterpret_cast<uint4*>(o + 2 * C) = *reinterpret_cast<const uint4*>(oh);
    }
  }
}

// out triple = mask ? (a + b) : 0 with mask = (aux_hi + aux_lo) > 0; a, b fp32 [rows][c] (a may be null), aux / out
// triples [rows][3c]. Joins the RPN and ROI data gradients at the backbone output (ReLU of the last conv).
__global__ void add_mask_x3_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const __half* __restrict__ aux, __half* __restrict__ out, int64_t rows, int c) {
  const int c4 = c / 4;
  const int64_t total = rows * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / c4;
    const int cc = static_cast<int>(i - r * c4) * 4;
    const float4 bv = reinterpret_cast<const float4*>(b)[i];
    float v[4] = {bv.x, bv.y, bv.z, bv.w};
    if (a != nullptr) {
      const float4 av = reinterpret_cast<const float4*>(a)[i];
      v[0] += av.x; v[1] += av.y; v[2] += av.z; v[3] += av.w;
    }
    __align__(8) __half h[4], l[4];
    const __half* xh = aux + r * 3 * c + cc;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xv = __half2float(xh[e]) + __half2float(xh[c + e]);
      split_hl(xv > 0.f ? v[e] : 0.f, h[e], l[e]);
    }
    __half* o = out + r * 3 * c + cc;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<const uint2*>(l);
    *reinterpret_cast<uint2*>(o + 2 * c) = *reinterpret_cast<const uint2*>(h);
  }
}

// Data-gradient weight operand in the f16x3 precision: src fp32 [rows][taps][cols] ->
// dst[c][td][Wh(rows) | Wh(rows) | Wl(rows)] of src * scale, td = taps-1-t when flip (conv) else t.
__global__ void transpose_pack_x3_kernel(const float* __restrict__ src, __half* __restrict__ dst, int rows, int cols,
                                         int taps, int flip, float scale) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int td = flip ? (taps - 1 - t) : t;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(static_cast<int64_t>(r) * taps + t) * cols + c] * scale : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      __half h, l;
      split_hl(tile[threadIdx.x][j], h, l);
      __half* o = dst + (static_cast<int64_t>(c) * taps + td) * 3 * rows + r;
      o[0] = h;
      o[rows] = h;
      o[2 * rows] = l;
    }
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ptb200_split3_pack_f16(const float* src, void* dst, int64_t rows, int k, float scale, int order,
                                      void* stream) {
  if (k % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
    split3_pack_vec8_kernel<<<grid_for(rows * (k / 8)), kThreads, 0, STREAM>>>(src, static_cast<__half*>(dst), rows, k,
                                                                            scale, order);
  else
    split3_pack_kernel<<<grid_for(rows * k), kThreads, 0, STREAM>>>(src, static_cast<__half*>(dst), rows, k, scale,
                                                                    order);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_split3_unpack_f32(const void* src, float* dst, int64_t rows, int k, void* stream) {
  split3_unpack_kernel<<<grid_for(rows * k), kThreads, 0, STREAM>>>(static_cast<const __half*>(src), dst, rows, k);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_maxpool2x2_f16x3(const void* in, void* out, int n, int h, int w, int c, void* stream) {
  if (c % 8 != 0) return 1201;
  const int64_t total = static_cast<int64_t>(n) * (h / 2) * (w / 2 + 1) * (c / 8);
  maxpool2x2_x3_kernel<<<grid_for(total), kThreads, 0, STREAM>>>(static_cast<const __half*>(in),
                                                                static_cast<__half*>(out), n, h, w, c);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_conv1_u8_f16x3(const uint8_t* images, const int* hw_dev, int n, int hmax, int wmax,
                                     int64_t image_stride, const float* mean3_host, const float* std3_host,
                                     const float* w_f32, const float* bias, void* out_f16x3, void* stream) {
  const int64_t total = static_cast<int64_t>(n) * hmax * (wmax + 1) * 4;
  conv1_u8_x3_kernel<<<grid_for(total), kThreads, 0, STREAM>>>(
      images, hw_dev, n, hmax, wmax, image_stride, mean3_host[0], mean3_host[1], mean3_host[2],
      1.f / std3_host[0], 1.f / std3_host[1], 1.f / std3_host[2], w_f32, bias, static_cast<__half*>(out_f16x3));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_bias_act_split3_f16(const float* in, int slices, const float* bias, int relu, float alpha,
                                          int64_t rows, int n, int wp, int w_valid, void* out3, void* stream) {
  if (n % 4 != 0 || slices < 1) return 1203;
  bias_act_split3_kernel<<<grid_for(rows * n / 4), kThreads, 0, STREAM>>>(in, slices, bias, relu, alpha, rows, n, wp,
                                                                         w_valid, static_cast<__half*>(out3));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_pack_grad2_f16x3(const float* d0, int n0, const float* d1, int n1, const float* g0,
                                       const float* g1, float lscale, int64_t rows, int ld, void* out3,
                                       void* stream) {
  pack_grad2_x3_kernel<<<grid_for(rows * ld), kThreads, 0, STREAM>>>(d0, n0, d1, n1, g0, g1, lscale, rows, ld,
                                                                    static_cast<__half*>(out3));
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_maxpool2x2_relu_bwd_f16x3(const void* x3, const float* dpooled, void* dz3, int n, int h, int w,
                                                int c, void* stream) {
  if (c % 8 != 0) return 1201;
  const int64_t total = static_cast<int64_t>(n) * ((h + 1) / 2) * ((w + 2) / 2) * (c / 8);
  maxpool2x2_relu_bwd_x3_kernel<<<grid_for(total), kThreads, 0, STREAM>>>(static_cast<const __half*>(x3), dpooled,
                                                                         static_cast<__half*>(dz3), n, h, w, c);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_add_mask_f16x3(const float* a, const float* b, const void* aux3, void* out3, int64_t rows, int c,
                                     void* stream) {
  if (c % 4 != 0) return 1203;
  add_mask_x3_kernel<<<grid_for(rows * c / 4), kThreads, 0, STREAM>>>(a, b, static_cast<const __half*>(aux3),
                                                                     static_cast<__half*>(out3), rows, c);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_transpose_pack_f16x3(const float* src, void* dst3, int rows, int cols, int taps, int flip,
                                           float scale, void* stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, taps), block(32, 8);
  transpose_pack_x3_kernel<<<grid, block, 0, STREAM>>>(src, static_cast<__half*>(dst3), rows, cols, taps, flip, scale);
  return static_cast<int>(cudaGetLastError());
}
