// ROIAlign (aligned=True, adaptive sampling grid) over fp16 NHWC-flat features, forward and backward.
// Semantics follow torchvision.ops.roi_align as called by detectron2's ROIPooler / ROIAlignV2
// (reference call site: pt/modeling/roi_heads/roi_heads.py:68-73,126). The output is written as
// fp16 [roi][ph*7+pw][C], i.e. directly in the K-major layout the fc1 GEMM consumes; all 512
// channels of a sample are read with 16-byte vector loads (one thread = 8 channels).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

struct Tap {
  int o1, o2, o3, o4;       // pixel offsets (in pixels, within the image plane)
  float w1, w2, w3, w4;
};

__device__ __forceinline__ bool bilinear_taps(float y, float x, int H, int W, int Wp, Tap& t) {
  if (y < -1.0f || y > H || x < -1.0f || x > W) return false;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = static_cast<int>(y), xl = static_cast<int>(x);
  int yh, xh;
  if (yl >= H - 1) {
    yh = yl = H - 1;
    y = static_cast<float>(yl);
  } else {
    yh = yl + 1;
  }
  if (xl >= W - 1) {
    xh = xl = W - 1;
    x = static_cast<float>(xl);
  } else {
    xh = xl + 1;
  }
  const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  t.w1 = hy * hx;
  t.w2 = hy * lx;
  t.w3 = ly * hx;
  t.w4 = ly * lx;
  t.o1 = yl * Wp + xl;
  t.o2 = yl * Wp + xh;
  t.o3 = yh * Wp + xl;
  t.o4 = yh * Wp + xh;
  return true;
}

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
  float count;
};

__device__ __forceinline__ RoiGeom roi_geom(const float4 b, float scale, int P) {
  RoiGeom g;
  g.start_w = b.x * scale - 0.5f;
  g.start_h = b.y * scale - 0.5f;
  const float rw = (b.z * scale - 0.5f) - g.start_w, rh = (b.w * scale - 0.5f) - g.start_h;
  g.bin_w = rw / P;
  g.bin_h = rh / P;
  g.grid_h = static_cast<int>(ceilf(rh / P));
  g.grid_w = static_cast<int>(ceilf(rw / P));
  g.count = fmaxf(static_cast<float>(g.grid_h * g.grid_w), 1.f);
  return g;
}

// Bilinear weights are separable and the sampling grid of a bin is a product grid, so the bin value
// is sum_py sum_px Wy[py] * Wx[px] * feat[py][px] with 1-D weight vectors over the rows / columns the
// samples touch: (gh+1)*(gw+1) pixel visits instead of 4*gh*gw taps. kMaxSpan bounds the 1-D span
// (bins of rois up to ~kMaxSpan*7*16 px); larger bins take the generic tap loop.
constexpr int kMaxSpan = 32;

struct Axis {
  int lo;            // first pixel index touched
  int n;             // number of pixels touched (0 = nothing valid); -1 = span too large
  float w[kMaxSpan];
};

// 1-D weights of the `grid` samples of one bin along one axis (extent = H or W); run by ONE thread,
// result lives in shared memory and is broadcast to the CTA.
__device__ __forceinline__ void axis_weights(float start, float bin, int p, int grid, int extent, Axis* a) {
  int n = 0, lo0 = 0;
  bool first = true;
  for (int i = 0; i < kMaxSpan; ++i) a->w[i] = 0.f;
  for (int i = 0; i < grid; ++i) {
    float v = start + p * bin + (i + 0.5f) * bin / grid;
    if (v < -1.0f || v > extent) continue;
    if (v <= 0.f) v = 0.f;
    int lo = static_cast<int>(v), hi;
    if (lo >= extent - 1) {
      hi = lo = extent - 1;
      v = static_cast<float>(lo);
    } else {
      hi = lo + 1;
    }
    const float l = v - lo, h = 1.f - l;
    if (first) {
      lo0 = lo;
      first = false;
    }
    const int i0 = lo - lo0, i1 = hi - lo0;
    if (i1 >= kMaxSpan) {
      a->n = -1;
      return;
    }
    a->w[i0] += h;
    a->w[i1] += l;
    if (i1 + 1 > n) n = i1 + 1;
  }
  a->lo = lo0;
  a->n = n;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

constexpr int kMaxP = 7;

// One CTA per roi: threads = P (bin columns) x C/8 (channel groups). The 2P axis-weight vectors of
// the roi are computed once (threads 0..2P-1) and shared by all P*P bins; each thread then walks the
// P bin rows of its column. BWD = true scatters the output gradient instead (fp32 vector atomics).
// X3 (forward only): feature rows and output bins are f16x3 triples [hi | lo | hi] of width 3C.
// GF32 (backward only): the output gradient is fp32 (f16x3 training path) instead of fp16.
template <bool BWD, bool X3 = false, bool GF32 = false>
__global__ void __launch_bounds__(512)
roi_align_roi_kernel(const __half* __restrict__ feat, const void* __restrict__ dout_v, int H, int W, int C,
                     const float4* __restrict__ rois, const int* __restrict__ roi_count, int cap, float scale,
                     int P, __half* __restrict__ out, float* __restrict__ dfeat) {
  __shared__ Axis s_ax[2 * kMaxP];  // [0,P): y axes of bin rows, [P,2P): x axes of bin columns
  const int roi = blockIdx.x;
  const int n = roi / cap, j = roi - n * cap;
  const int c8 = C >> 3;
  const int pw = threadIdx.x / c8, cg = threadIdx.x - pw * c8;
  const int c0 = cg * 8;
  const int Wp = W + 1;
  const int LD = X3 ? 3 * C : C;  // row pitch of feat / out in elements
  const bool valid = roi_count == nullptr || j < roi_count[n];
  if (!valid) {
    if (!BWD)
      for (int ph = 0; ph < P; ++ph)
        for (int s3 = 0; s3 < (X3 ? 3 : 1); ++s3)
          *reinterpret_cast<uint4*>(out + (static_cast<int64_t>(roi) * P * P + ph * P + pw) * LD + s3 * C + c0) =
              make_uint4(0, 0, 0, 0);
    return;
  }
  const RoiGeom g = roi_geom(rois[roi], scale, P);
  if (threadIdx.x < 2 * P) {
    const int a = threadIdx.x;
    if (a < P)
      axis_weights(g.start_h, g.bin_h, a, g.grid_h, H, &s_ax[a]);
    else
      axis_weights(g.start_w, g.bin_w, a - P, g.grid_w, W, &s_ax[a]);
  }
  __syncthreads();
  const Axis& ax = s_ax[P + pw];
  const bool sep_x = ax.n >= 0;
  const int64_t img_off = static_cast<int64_t>(n) * H * Wp * LD + c0;
  for (int ph = 0; ph < P; ++ph) {
    const Axis& ay = s_ax[ph];
    const int bin = ph * P + pw;
    // X3: a roi row is the K-concatenation [hi (P*P*C) | lo (P*P*C) | hi (P*P*C)] of the plain [bin][C] row, so
    // that fc1 sees an ordinary activation triple (forward weights [Wh | Wh | Wl] over K = P*P*C, weight gradient
    // over column slices)
    const int64_t o_off = X3 ? (static_cast<int64_t>(roi) * 3 * P * P + bin) * C + c0
                             : (static_cast<int64_t>(roi) * P * P + bin) * LD + c0;
    float acc[8];
    float gr[8];
    if (BWD) {
      bool any = false;
      if (GF32) {
        const float* gp = static_cast<const float*>(dout_v) + o_off;
        const float4 g0 = *reinterpret_cast<const float4*>(gp), g1 = *reinterpret_cast<const float4*>(gp + 4);
        const float gf[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          gr[e] = gf[e] / g.count;
          any = any || gf[e] != 0.f;
        }
      } else {
        const uint4 gv = *reinterpret_cast<const uint4*>(static_cast<const __half*>(dout_v) + o_off);
        const __half2* gh = reinterpret_cast<const __half2*>(&gv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(gh[e]);
          gr[2 * e] = f.x / g.count;
          gr[2 * e + 1] = f.y / g.count;
          any = any || f.x != 0.f || f.y != 0.f;
        }
      }
      if (!any) continue;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    }
    if (sep_x && ay.n >= 0) {
      for (int iy = 0; iy < ay.n; ++iy) {
        const float wy = ay.w[iy];
        if (wy == 0.f) continue;
        const int64_t row_off = img_off + static_cast<int64_t>((ay.lo + iy) * Wp + ax.lo) * LD;
        for (int ix = 0; ix < ax.n; ++ix) {
          const float wgt = wy * ax.w[ix];
          if (wgt == 0.f) continue;
          if (BWD) {
            float* p = dfeat + row_off + static_cast<int64_t>(ix) * LD;
            red_add_v4(p, gr[0] * wgt, gr[1] * wgt, gr[2] * wgt, gr[3] * wgt);
            red_add_v4(p + 4, gr[4] * wgt, gr[5] * wgt, gr[6] * wgt, gr[7] * wgt);
          } else {
            const uint4 v = *reinterpret_cast<const uint4*>(feat + row_off + static_cast<int64_t>(ix) * LD);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
            uint4 vl = make_uint4(0, 0, 0, 0);
            if (X3) vl = *reinterpret_cast<const uint4*>(feat + row_off + static_cast<int64_t>(ix) * LD + C);
            const __half2* hl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __half22float2(h[e]);
              if (X3) {
                const float2 fl = __half22float2(hl[e]);
                f.x += fl.x;
                f.y += fl.y;
              }
              acc[2 * e] += wgt * f.x;
              acc[2 * e + 1] += wgt * f.y;
            }
          }
        }
      }
    } else {
      // very large bins: generic tap loop (torchvision order)
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / g.grid_h;
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / g.grid_w;
          Tap t;
          if (!bilinear_taps(y, x, H, W, Wp, t)) continue;
          const int offs[4] = {t.o1, t.o2, t.o3, t.o4};
          const float ws[4] = {t.w1, t.w2, t.w3, t.w4};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (BWD) {
              float* p = dfeat + img_off + static_cast<int64_t>(offs[k]) * LD;
              red_add_v4(p, gr[0] * ws[k], gr[1] * ws[k], gr[2] * ws[k], gr[3] * ws[k]);
              red_add_v4(p + 4, gr[4] * ws[k], gr[5] * ws[k], gr[6] * ws[k], gr[7] * ws[k]);
            } else {
              const uint4 v = *reinterpret_cast<const uint4*>(feat + img_off + static_cast<int64_t>(offs[k]) * LD);
              const __half2* h = reinterpret_cast<const __half2*>(&v);
              uint4 vl = make_uint4(0, 0, 0, 0);
              if (X3) vl = *reinterpret_cast<const uint4*>(feat + img_off + static_cast<int64_t>(offs[k]) * LD + C);
              const __half2* hl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(h[e]);
                if (X3) {
                  const float2 fl = __half22float2(hl[e]);
                  f.x += fl.x;
                  f.y += fl.y;
                }
                acc[2 * e] += ws[k] * f.x;
                acc[2 * e + 1] += ws[k] * f.y;
              }
            }
          }
        }
      }
    }
    if (!BWD) {
      __align__(16) __half r[8];
      __align__(16) __half rl[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float x = acc[e] / g.count;
        r[e] = __float2half_rn(x);
        if (X3) rl[e] = __float2half_rn(x - __half2float(r[e]));
      }
      *reinterpret_cast<uint4*>(out + o_off) = *reinterpret_cast<const uint4*>(r);
      if (X3) {
        const int64_t fin = static_cast<int64_t>(P) * P * C;
        *reinterpret_cast<uint4*>(out + o_off + fin) = *reinterpret_cast<const uint4*>(rl);
        *reinterpret_cast<uint4*>(out + o_off + 2 * fin) = *reinterpret_cast<const uint4*>(r);
      }
    }
  }
}

// out = half(mask ? a + scale * b : 0), mask = aux > 0 (aux optional)
__global__ void add_mask_kernel(const __half* __restrict__ a, const float* __restrict__ b, float scale,
                                const __half* __restrict__ aux, __half* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = (a != nullptr ? __half2float(a[i]) : 0.f) + scale * b[i];
    if (aux != nullptr && !(__half2float(aux[i]) > 0.f)) v = 0.f;
    out[i] = __float2half_rn(v);
  }
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ptb200_roi_align_fwd_f16(const void* feat, int n, int h, int w, int c, const float* rois,
                                        const int* roi_count, int cap, float spatial_scale, int pooled, void* out,
                                        void* stream) {
  if (c % 8 != 0 || pooled > kMaxP || pooled * (c / 8) > 512) return 1401;
  roi_align_roi_kernel<false><<<n * cap, pooled * (c / 8), 0, STREAM>>>(
      static_cast<const __half*>(feat), nullptr, h, w, c, reinterpret_cast<const float4*>(rois), roi_count, cap,
      spatial_scale, pooled, static_cast<__half*>(out), nullptr);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_roi_align_fwd_f16x3(const void* feat, int n, int h, int w, int c, const float* rois,
                                          const int* roi_count, int cap, float spatial_scale, int pooled, void* out,
                                          void* stream) {
  if (c % 8 != 0 || pooled > kMaxP || pooled * (c / 8) > 512) return 1401;
  roi_align_roi_kernel<false, true><<<n * cap, pooled * (c / 8), 0, STREAM>>>(
      static_cast<const __half*>(feat), nullptr, h, w, c, reinterpret_cast<const float4*>(rois), roi_count, cap,
      spatial_scale, pooled, static_cast<__half*>(out), nullptr);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_roi_align_bwd_f16(const void* dout, int n, int h, int w, int c, const float* rois,
                                        const int* roi_count, int cap, float spatial_scale, int pooled,
                                        float* dfeat, void* stream) {
  if (c % 8 != 0 || pooled > kMaxP || pooled * (c / 8) > 512) return 1401;
  roi_align_roi_kernel<true><<<n * cap, pooled * (c / 8), 0, STREAM>>>(
      nullptr, static_cast<const __half*>(dout), h, w, c, reinterpret_cast<const float4*>(rois), roi_count, cap,
      spatial_scale, pooled, nullptr, dfeat);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_roi_align_bwd_f32(const float* dout, int n, int h, int w, int c, const float* rois,
                                        const int* roi_count, int cap, float spatial_scale, int pooled,
                                        float* dfeat, void* stream) {
  if (c % 8 != 0 || pooled > kMaxP || pooled * (c / 8) > 512) return 1401;
  roi_align_roi_kernel<true, false, true><<<n * cap, pooled * (c / 8), 0, STREAM>>>(
      nullptr, dout, h, w, c, reinterpret_cast<const float4*>(rois), roi_count, cap, spatial_scale, pooled, nullptr,
      dfeat);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_add_mask_f16(const void* a, const float* b, float scale, const void* aux, void* out,
                                   int64_t n, void* stream) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  add_mask_kernel<<<static_cast<int>(g), 256, 0, STREAM>>>(static_cast<const __half*>(a), b, scale,
                                                          static_cast<const __half*>(aux),
                                                          static_cast<__half*>(out), n);
  return static_cast<int>(cudaGetLastError());
}
