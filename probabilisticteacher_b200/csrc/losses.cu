// Fused forward + analytic backward of the eight Probabilistic Teacher losses. Every kernel
// returns the normalised loss values (block partial sums -> atomicAdd) and the gradient of each
// loss w.r.t. the head outputs for a unit upstream gradient; ptb200_pack_grad2_f16 later scales
// them by the upstream gradients / loss scale and packs them as the fp16 operand of the data- and
// weight-gradient GEMMs.
//   RPN supervised     pt/modeling/proposal_generator/rpn.py:191-255 + box_regression.py:33-35,142-176
//   RPN unsupervised   pt/modeling/proposal_generator/rpn.py:257-361 (keeps the sigmoid([1-x, x]) quirk, :299)
//   ROI supervised     detectron2 FastRCNNOutputLayers.losses (mean CE) + fast_rcnn.py:265-336
//   ROI unsupervised   pt/modeling/roi_heads/fast_rcnn.py:179-263 + roi_heads.py:131-172
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

constexpr float kTwoPi = 6.283185307179586f;
constexpr float kHalfLog2PiE = 1.4189385332046727f;  // 0.5 * log(2*pi*e)
constexpr float kTwoPiE = 17.079468445347132f;       // 2*pi*e

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void block_add2(float a, float b, float* out) {
  __shared__ float sa[32], sb[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    sa[warp] = a;
    sb[warp] = b;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? sa[lane] : 0.f;
    b = lane < nw ? sb[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      if (a != 0.f) atomicAdd(out, a);
      if (b != 0.f) atomicAdd(out + 1, b);
    }
  }
}

// box_regression.py:66-99
__device__ __forceinline__ void get_deltas(const float4 s, const float4 t, float wx, float wy, float ww, float wh,
                                           float* d) {
  const float sw = s.z - s.x, sh = s.w - s.y;
  const float sx = s.x + 0.5f * sw, sy = s.y + 0.5f * sh;
  const float tw = t.z - t.x, th = t.w - t.y;
  const float tx = t.x + 0.5f * tw, ty = t.y + 0.5f * th;
  d[0] = wx * (tx - sx) / sw;
  d[1] = wy * (ty - sy) / sh;
  d[2] = ww * logf(tw / sw + 1e-9f);
  d[3] = wh * logf(th / sh + 1e-9f);
}

// -log(gaussian_dist_pdf(mu, t, v) + 1e-9), v = sigmoid(s); returns loss, writes d/dmu and d/ds
__device__ __forceinline__ float gauss_nll(float mu, float s, float t, float* gmu, float* gs) {
  const float v = sigm(s);
  const float diff = mu - t;
  const float e = expf(-(diff * diff) / (v + 1e-9f) / 2.0f);
  const float pdf = e / sqrtf(kTwoPi * (v + 0.3f));
  const float q = pdf / (pdf + 1e-9f);
  *gmu = q * diff / (v + 1e-9f);
  const float dv = -q * (diff * diff / (2.f * (v + 1e-9f) * (v + 1e-9f)) - 1.f / (2.f * (v + 0.3f)));
  *gs = dv * v * (1.f - v);
  return -logf(pdf + 1e-9f);
}

// KL(N_p || N_q) term of rpn.py:337-339 / fast_rcnn.py:247-249 with entropy weight wb.
__device__ __forceinline__ float gauss_kl(float mq, float sq, float mp, float vp, float wb, float* gmq, float* gsq,
                                          float* gmp) {
  const float vq = sigm(sq);
  const float diff = mq - mp;
  const float l = 0.5f * logf(vq / vp) - 0.5f + (vp + diff * diff) / (2.f * vq);
  *gmq = wb * diff / vq;
  *gmp = -wb * diff / vq;
  const float dvq = wb * (0.5f / vq - (vp + diff * diff) / (2.f * vq * vq));
  *gsq = dvq * vq * (1.f - vq);
  return wb * l;
}

// ------------------------------------------------------------------------------------------
__global__ void rpn_loss_sup_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ deltas,
                                    int ldd, const signed char* __restrict__ labels, const int* __restrict__ matched,
                                    const float4* __restrict__ gt, int gt_cap, const float4* __restrict__ anchors,
                                    int N, int H, int W, int A, float norm, float* __restrict__ loss,
                                    float* __restrict__ dlogits, float* __restrict__ ddeltas) {
  const int R = H * W * A, Wp = W + 1;
  const int64_t total = static_cast<int64_t>(N) * R;
  float lc = 0.f, ll = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int lab = labels[i];
    if (lab < 0) continue;
    const int n = static_cast<int>(i / R);
    const int r = static_cast<int>(i - static_cast<int64_t>(n) * R);
    const int a = r % A, loc = r / A;
    const int64_t row = static_cast<int64_t>(n) * H * Wp + (loc / W) * Wp + (loc % W);
    const float x = logits[row * ldl + a];
    const float y = static_cast<float>(lab);
    lc += (fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)))) * norm;
    dlogits[row * A + a] = (sigm(x) - y) * norm;
    if (lab == 1) {
      float t[4];
      get_deltas(anchors[r], gt[n * gt_cap + matched[i]], 1.f, 1.f, 1.f, 1.f, t);
      const float* d = deltas + row * ldd + a * 8;
      float* g = ddeltas + row * (A * 8) + a * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gm, gs;
        ll += gauss_nll(d[k], d[4 + k], t[k], &gm, &gs) * norm;
        g[k] = gm * norm;
        g[4 + k] = gs * norm;
      }
    }
  }
  block_add2(lc, ll, loss);
}

__global__ void rpn_loss_unsup_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ deltas,
                                      int ldd, const int* __restrict__ labels, const int* __restrict__ matched,
                                      const float4* __restrict__ pseudo, const float* __restrict__ pseudo_logits,
                                      const float* __restrict__ pseudo_sigma, int ps_cap,
                                      const float4* __restrict__ anchors, int N, int H, int W, int A, int K1, int efl,
                                      float lam0, float lam1, float tau0, float tau1, float norm,
                                      float* __restrict__ loss, float* __restrict__ dlogits,
                                      float* __restrict__ ddeltas, float* __restrict__ danchor_wh) {
  const int R = H * W * A, Wp = W + 1;
  const int64_t total = static_cast<int64_t>(N) * R;
  float lc = 0.f, ll = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (labels[i] != 1) continue;
    const int n = static_cast<int>(i / R);
    const int r = static_cast<int>(i - static_cast<int64_t>(n) * R);
    const int a = r % A, loc = r / A;
    const int64_t row = static_cast<int64_t>(n) * H * Wp + (loc / W) * Wp + (loc % W);
    const int m = n * ps_cap + matched[i];
    const float* zt = pseudo_logits + static_cast<int64_t>(m) * K1;
    // teacher distribution: entropy weight, fg flag, tempered 2-way target
    float mx = zt[0];
    int arg = 0;
    for (int q = 1; q < K1; ++q)
      if (zt[q] > mx) {
        mx = zt[q];
        arg = q;
      }
    float den = 0.f, dent = 0.f;
    for (int q = 0; q < K1; ++q) {
      den += expf(zt[q] - mx);
      dent += expf((zt[q] - mx) / tau0);
    }
    float w = 1.f;
    if (efl) {
      float ent = 0.f;
      for (int q = 0; q < K1; ++q) {
        const float pq = expf(zt[q] - mx) / den;
        ent -= pq * logf(pq);
      }
      w = powf(1.f - ent / logf(static_cast<float>(K1)), lam0);
    }
    const float t_bg = expf((zt[K1 - 1] - mx) / tau0) / dent;
    float t_fg = 0.f;
    for (int q = 0; q < K1 - 1; ++q) t_fg += expf((zt[q] - mx) / tau0) / dent;
    const float x = logits[row * ldl + a];
    const float p0 = sigm(1.f - x), p1 = sigm(x);
    lc += (t_bg * w * (-logf(p0 + 1e-9f)) + t_fg * w * (-logf(p1 + 1e-9f))) * norm;
    dlogits[row * A + a] =
        (t_bg * w * (p0 * (1.f - p0)) / (p0 + 1e-9f) - t_fg * w * (p1 * (1.f - p1)) / (p1 + 1e-9f)) * norm;
    if (arg != K1 - 1) {
      const float4 an = anchors[r];
      const float4 pb = pseudo[m];
      float mp[4];
      get_deltas(an, pb, 1.f, 1.f, 1.f, 1.f, mp);
      const float* d = deltas + row * ldd + a * 8;
      float* g = ddeltas + row * (A * 8) + a * 8;
      float gmp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float vp = sigm(pseudo_sigma[m * 4 + k]);
        float wb = 1.f;
        if (efl) wb = powf(1.f - (0.5f * logf(kTwoPiE * vp)) / kHalfLog2PiE, lam1);
        vp *= tau1;
        float gm, gs;
        ll += gauss_kl(d[k], d[4 + k], mp[k], vp, wb, &gm, &gs, &gmp[k]) * norm;
        g[k] = gm * norm;
        g[4 + k] = gs * norm;
        gmp[k] *= norm;
      }
      if (danchor_wh != nullptr) {
        // chain d(mean_p)/d(anchor box) -> d/d(anchor w, h): ax1 = sx - w/2, ax2 = sx + w/2
        const float sw = an.z - an.x, sh = an.w - an.y;
        const float tw = pb.z - pb.x, th = pb.w - pb.y;
        const float tx = pb.x + 0.5f * tw, ty = pb.y + 0.5f * th;
        const float sx = an.x + 0.5f * sw, sy = an.y + 0.5f * sh;
        // d(dx)/d(sw) with the centre fixed = -(tx - sx)/sw^2 ; d(dw)/d(sw) = -(tw/sw^2)/(tw/sw + 1e-9)
        const float ddx = -(tx - sx) / (sw * sw), ddw = -(tw / (sw * sw)) / (tw / sw + 1e-9f);
        const float ddy = -(ty - sy) / (sh * sh), ddh = -(th / (sh * sh)) / (th / sh + 1e-9f);
        atomicAdd(danchor_wh + 2 * a, gmp[0] * ddx + gmp[2] * ddw);
        atomicAdd(danchor_wh + 2 * a + 1, gmp[1] * ddy + gmp[3] * ddh);
      }
    }
  }
  block_add2(lc, ll, loss);
}

// ------------------------------------------------------------------------------------------
__global__ void roi_loss_sup_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                    const int* __restrict__ gt_cls, const float4* __restrict__ props,
                                    const float4* __restrict__ gt, const int* __restrict__ counts, int N, int cap,
                                    int K, float wx, float wy, float ww, float wh, float* __restrict__ loss,
                                    float* __restrict__ dscores, float* __restrict__ ddeltas) {
  __shared__ float s_norm;
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int n = 0; n < N; ++n) tot += min(counts[n], cap);
    s_norm = 1.f / fmaxf(static_cast<float>(tot), 1.f);
  }
  __syncthreads();
  const float norm = s_norm;
  const int K1 = K + 1;
  float lc = 0.f, lb = 0.f;
  const int total = N * cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / cap, j = i - n * cap;
    float* gz = dscores + static_cast<int64_t>(i) * K1;
    float* gd = ddeltas + static_cast<int64_t>(i) * (8 * K);
    if (j >= min(counts[n], cap)) continue;
    const int c = gt_cls[i];
    const float* z = scores + static_cast<int64_t>(i) * K1;
    float mx = z[0];
    for (int q = 1; q < K1; ++q) mx = fmaxf(mx, z[q]);
    float den = 0.f;
    for (int q = 0; q < K1; ++q) den += expf(z[q] - mx);
    const float lse = mx + logf(den);
    lc += (lse - z[c]) * norm;
    for (int q = 0; q < K1; ++q) gz[q] = (expf(z[q] - lse) - (q == c ? 1.f : 0.f)) * norm;
    if (c >= 0 && c < K) {
      float t[4];
      get_deltas(props[i], gt[i], wx, wy, ww, wh, t);
      const float* d = deltas + static_cast<int64_t>(i) * (8 * K) + c * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gm, gs;
        lb += gauss_nll(d[k], d[4 + k], t[k], &gm, &gs) * norm;
        gd[c * 8 + k] = gm * norm;
        gd[c * 8 + 4 + k] = gs * norm;
      }
    }
  }
  block_add2(lc, lb, loss);
}

// counts[0] = number of valid rows, counts[1] = rows whose teacher arg-max is a foreground class
__global__ void roi_unsup_count_kernel(const float* __restrict__ soft, const int* __restrict__ counts, int N,
                                       int cap, int K1, int* __restrict__ out) {
  int rows = 0, fg = 0;
  const int total = N * cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / cap, j = i - n * cap;
    if (j >= min(counts[n], cap)) continue;
    const float* zt = soft + static_cast<int64_t>(i) * K1;
    float mx = zt[0];
    int arg = 0;
    for (int q = 1; q < K1; ++q)
      if (zt[q] > mx) {
        mx = zt[q];
        arg = q;
      }
    rows += 1;
    fg += arg != K1 - 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rows += __shfl_xor_sync(0xffffffffu, rows, o);
    fg += __shfl_xor_sync(0xffffffffu, fg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rows) atomicAdd(out, rows);
    if (fg) atomicAdd(out + 1, fg);
  }
}

__global__ void roi_loss_unsup_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                      const float* __restrict__ soft, const float* __restrict__ sigma_t,
                                      const float4* __restrict__ props, const float4* __restrict__ pseudo,
                                      const int* __restrict__ counts, const int* __restrict__ totals, int N, int cap,
                                      int K, int efl, float lam0, float lam1, float tau0, float tau1, float wx,
                                      float wy, float ww, float wh, float* __restrict__ loss,
                                      float* __restrict__ dscores, float* __restrict__ ddeltas) {
  const int K1 = K + 1;
  // 0/0 -> NaN when no roi matched a pseudo box, as in the reference (fast_rcnn.py:208-209,260)
  const float norm_c = 1.f / static_cast<float>(totals[0]);
  const float norm_b = 1.f / (4.f * static_cast<float>(totals[1]));
  float lc = 0.f, lb = 0.f;
  const int total = N * cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / cap, j = i - n * cap;
    if (j >= min(counts[n], cap)) continue;
    const float* zs = scores + static_cast<int64_t>(i) * K1;
    const float* zt = soft + static_cast<int64_t>(i) * K1;
    float* gz = dscores + static_cast<int64_t>(i) * K1;
    float smx = zs[0], tmx = zt[0];
    int arg = 0;
    for (int q = 1; q < K1; ++q) {
      smx = fmaxf(smx, zs[q]);
      if (zt[q] > tmx) {
        tmx = zt[q];
        arg = q;
      }
    }
    float sden = 0.f, tden = 0.f, tden_tau = 0.f;
    for (int q = 0; q < K1; ++q) {
      sden += expf(zs[q] - smx);
      tden += expf(zt[q] - tmx);
      tden_tau += expf((zt[q] - tmx) / tau0);
    }
    const float lse = smx + logf(sden);
    float w = 1.f;
    if (efl) {
      float ent = 0.f;
      for (int q = 0; q < K1; ++q) {
        const float pq = expf(zt[q] - tmx) / tden;
        ent -= pq * logf(pq);
      }
      w = powf(1.f - ent / logf(static_cast<float>(K1)), lam0);
    }
    float tsum = 0.f;
    for (int q = 0; q < K1; ++q) {
      const float t = expf((zt[q] - tmx) / tau0) / tden_tau * w;
      tsum += t;
      lc += t * (lse - zs[q]) * norm_c;
    }
    for (int q = 0; q < K1; ++q) {
      const float t = expf((zt[q] - tmx) / tau0) / tden_tau * w;
      gz[q] = (tsum * expf(zs[q] - lse) - t) * norm_c;
    }
    if (arg != K) {
      float mp[4];
      get_deltas(props[i], pseudo[i], wx, wy, ww, wh, mp);
      const float* d = deltas + static_cast<int64_t>(i) * (8 * K) + arg * 8;
      float* gd = ddeltas + static_cast<int64_t>(i) * (8 * K) + arg * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float vp = sigm(sigma_t[static_cast<int64_t>(i) * 4 + k]);
        float wb = 1.f;
        if (efl) wb = powf(1.f - (0.5f * logf(kTwoPiE * vp)) / kHalfLog2PiE, lam1);
        vp *= tau1;
        float gm, gs, gmp;
        lb += gauss_kl(d[k], d[4 + k], mp[k], vp, wb, &gm, &gs, &gmp) * norm_b;
        gd[k] = gm * norm_b;
        gd[4 + k] = gs * norm_b;
      }
    }
  }
  block_add2(lc, lb, loss);
}

__global__ void axpy_dev_kernel(const float* __restrict__ alpha, float scale, const float* __restrict__ x,
                                float* __restrict__ y, int n) {
  const float a = (alpha != nullptr ? alpha[0] : 1.f) * scale;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += a * x[i];
}

inline int grid1d(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)
#define LAUNCH_OK() static_cast<int>(cudaGetLastError())

extern "C" int ptb200_rpn_loss_sup(const float* logits, int ld_logit, const float* deltas, int ld_delta,
                                   const signed char* labels, const int* matched, const float* gt_boxes, int gt_cap,
                                   const float* anchors, int n, int h, int w, int num_cell, float norm, float* loss2,
                                   float* dlogits, float* ddeltas, void* stream) {
  const int64_t rows = static_cast<int64_t>(n) * h * (w + 1);
  cudaMemsetAsync(loss2, 0, 2 * sizeof(float), STREAM);
  cudaMemsetAsync(dlogits, 0, rows * num_cell * sizeof(float), STREAM);
  cudaMemsetAsync(ddeltas, 0, rows * num_cell * 8 * sizeof(float), STREAM);
  rpn_loss_sup_kernel<<<grid1d(static_cast<int64_t>(n) * h * w * num_cell), 256, 0, STREAM>>>(
      logits, ld_logit, deltas, ld_delta, labels, matched, reinterpret_cast<const float4*>(gt_boxes), gt_cap,
      reinterpret_cast<const float4*>(anchors), n, h, w, num_cell, norm, loss2, dlogits, ddeltas);
  return LAUNCH_OK();
}

extern "C" int ptb200_rpn_loss_unsup(const float* logits, int ld_logit, const float* deltas, int ld_delta,
                                     const int* labels, const int* matched, const float* pseudo_boxes,
                                     const float* pseudo_logits, const float* pseudo_sigma, int pseudo_cap,
                                     const float* anchors, int n, int h, int w, int num_cell, int num_classes_plus1,
                                     int efl, float lam0, float lam1, float tau0, float tau1, float norm,
                                     float* loss2, float* dlogits, float* ddeltas, float* danchor_wh, void* stream) {
  const int64_t rows = static_cast<int64_t>(n) * h * (w + 1);
  cudaMemsetAsync(loss2, 0, 2 * sizeof(float), STREAM);
  cudaMemsetAsync(dlogits, 0, rows * num_cell * sizeof(float), STREAM);
  cudaMemsetAsync(ddeltas, 0, rows * num_cell * 8 * sizeof(float), STREAM);
  if (danchor_wh != nullptr) cudaMemsetAsync(danchor_wh, 0, num_cell * 2 * sizeof(float), STREAM);
  rpn_loss_unsup_kernel<<<grid1d(static_cast<int64_t>(n) * h * w * num_cell), 256, 0, STREAM>>>(
      logits, ld_logit, deltas, ld_delta, labels, matched, reinterpret_cast<const float4*>(pseudo_boxes),
      pseudo_logits, pseudo_sigma, pseudo_cap, reinterpret_cast<const float4*>(anchors), n, h, w, num_cell,
      num_classes_plus1, efl, lam0, lam1, tau0, tau1, norm, loss2, dlogits, ddeltas, danchor_wh);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_loss_sup(const float* scores, const float* deltas, const int* gt_classes,
                                   const float* props, const float* gt_boxes, const int* counts, int n, int cap,
                                   int num_classes, const float* weights4, float* loss2, float* dscores,
                                   float* ddeltas, void* stream) {
  const int64_t rows = static_cast<int64_t>(n) * cap;
  cudaMemsetAsync(loss2, 0, 2 * sizeof(float), STREAM);
  cudaMemsetAsync(dscores, 0, rows * (num_classes + 1) * sizeof(float), STREAM);
  cudaMemsetAsync(ddeltas, 0, rows * num_classes * 8 * sizeof(float), STREAM);
  roi_loss_sup_kernel<<<grid1d(rows), 256, 0, STREAM>>>(
      scores, deltas, gt_classes, reinterpret_cast<const float4*>(props), reinterpret_cast<const float4*>(gt_boxes),
      counts, n, cap, num_classes, weights4[0], weights4[1], weights4[2], weights4[3], loss2, dscores, ddeltas);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_loss_unsup(const float* scores, const float* deltas, const float* soft_logits,
                                     const float* sigma_t, const float* props, const float* pseudo_boxes,
                                     const int* counts, int n, int cap, int num_classes, int efl, float lam0,
                                     float lam1, float tau0, float tau1, const float* weights4, int* totals2,
                                     float* loss2, float* dscores, float* ddeltas, void* stream) {
  const int64_t rows = static_cast<int64_t>(n) * cap;
  cudaMemsetAsync(loss2, 0, 2 * sizeof(float), STREAM);
  cudaMemsetAsync(totals2, 0, 2 * sizeof(int), STREAM);
  cudaMemsetAsync(dscores, 0, rows * (num_classes + 1) * sizeof(float), STREAM);
  cudaMemsetAsync(ddeltas, 0, rows * num_classes * 8 * sizeof(float), STREAM);
  roi_unsup_count_kernel<<<grid1d(rows), 256, 0, STREAM>>>(soft_logits, counts, n, cap, num_classes + 1, totals2);
  roi_loss_unsup_kernel<<<grid1d(rows), 256, 0, STREAM>>>(
      scores, deltas, soft_logits, sigma_t, reinterpret_cast<const float4*>(props),
      reinterpret_cast<const float4*>(pseudo_boxes), counts, totals2, n, cap, num_classes, efl, lam0, lam1, tau0,
      tau1, weights4[0], weights4[1], weights4[2], weights4[3], loss2, dscores, ddeltas);
  return LAUNCH_OK();
}

extern "C" int ptb200_axpy_dev(const float* alpha_dev, float scale, const float* x, float* y, int n, void* stream) {
  axpy_dev_kernel<<<grid1d(n), 256, 0, STREAM>>>(alpha_dev, scale, x, y, n);
  return LAUNCH_OK();
}
