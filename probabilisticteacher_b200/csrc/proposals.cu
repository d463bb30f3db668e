// Proposal path of the Gaussian RPN and the teacher's pseudo-label filter, all on device with
// fixed-capacity buffers and device-side counts (no host synchronisation):
//   anchors            pt/modeling/anchor_generator.py:108-122,145-148
//   decode / clip      pt/modeling/box_regression.py:101-139, detectron2 Boxes.clip / nonempty
//   top-k + rescoring  pt/modeling/proposal_generator/proposal_utils.py:83-138 (incl. the :94 quirk)
//   NMS                detectron2 batched_nms -> torchvision nms (IoU > thr suppresses, stable order)
//   pseudo-label filter pt/modeling/roi_heads/fast_rcnn.py:34-120
// Head outputs are indexed by "row" = y * (W+1) + x of the flattened, right-padded feature map.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/ptb200.h"
#include "ptx.cuh"

namespace {

__device__ __forceinline__ uint32_t order_desc(float f) {
  // ascending radix order of the returned key == descending float order
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

constexpr float kScaleClamp = 4.135166556742356f;  // log(1000/16), box_regression.py:28

__device__ __forceinline__ void decode_box(const float* a, float dx, float dy, float dw, float dh,
                                           float wx, float wy, float ww, float wh, float* o) {
  const float w = a[2] - a[0], h = a[3] - a[1];
  const float cx = a[0] + 0.5f * w, cy = a[1] + 0.5f * h;
  dx = dx / wx;
  dy = dy / wy;
  dw = fminf(dw / ww, kScaleClamp);
  dh = fminf(dh / wh, kScaleClamp);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  o[0] = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  o[1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o[2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  o[3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
}

// ------------------------------------------------------------------------------------------
__global__ void cell_anchors_from_wh_kernel(const float* __restrict__ wh, int A, float* __restrict__ cell) {
  const int a = threadIdx.x;
  if (a < A) {
    cell[4 * a + 0] = -wh[2 * a] / 2.0f;
    cell[4 * a + 1] = -wh[2 * a + 1] / 2.0f;
    cell[4 * a + 2] = wh[2 * a] / 2.0f;
    cell[4 * a + 3] = wh[2 * a + 1] / 2.0f;
  }
}

__global__ void anchor_grid_kernel(const float* __restrict__ cell, int A, int H, int W, float stride,
                                   float offset, float4* __restrict__ out) {
  const int total = H * W * A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int a = i % A;
    const int loc = i / A;
    const int x = loc % W, y = loc / W;
    const float sx = offset * stride + x * stride, sy = offset * stride + y * stride;
    out[i] = make_float4(sx + cell[4 * a], sy + cell[4 * a + 1], sx + cell[4 * a + 2], sy + cell[4 * a + 3]);
  }
}

// keys/vals for the descending sort of objectness logits (one segment per image)
__global__ void rpn_make_keys_kernel(const float* __restrict__ logits, int ld, int N, int H, int W, int A,
                                     uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int R = H * W * A, Wp = W + 1;
  const int64_t total = static_cast<int64_t>(N) * R;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / R);
    const int r = static_cast<int>(i - static_cast<int64_t>(n) * R);
    const int a = r % A, loc = r / A;
    const int x = loc % W, y = loc / W;
    const int64_t row = static_cast<int64_t>(n) * H * Wp + y * Wp + x;
    keys[i] = order_desc(logits[row * ld + a]);
    vals[i] = static_cast<uint32_t>(r);
  }
}

// For the j-th best anchor of image n: decode, clip, validity, sigma-rescored score.
// sigma is read from anchor index j (NOT the sorted index) -- proposal_utils.py:94.
__global__ void rpn_topk_decode_kernel(const uint32_t* __restrict__ sorted_idx, int64_t idx_stride,
                                       const float* __restrict__ logits, int ld_logit,
                                       const float* __restrict__ deltas, int ld_delta,
                                       const float4* __restrict__ anchors, int N, int H, int W, int A, int k,
                                       const float* __restrict__ img_hw, float min_size,
                                       float4* __restrict__ boxes, float* __restrict__ scores,
                                       uint32_t* __restrict__ keys2, uint32_t* __restrict__ vals2,
                                       int* __restrict__ valid_count, int* __restrict__ nonfinite_flag) {
  const int Wp = W + 1;
  const int total = N * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / k, j = i - n * k;
    const int r = static_cast<int>(sorted_idx[n * idx_stride + j]);
    const int a = r % A, loc = r / A;
    const int64_t row = static_cast<int64_t>(n) * H * Wp + (loc / W) * Wp + (loc % W);
    const float logit = logits[row * ld_logit + a];
    const float* d = deltas + row * ld_delta + a * 8;
    const float4 an = anchors[r];
    const float av[4] = {an.x, an.y, an.z, an.w};
    float b[4];
    decode_box(av, d[0], d[1], d[2], d[3], 1.f, 1.f, 1.f, 1.f, b);
    bool finite = isfinite(b[0]) && isfinite(b[1]) && isfinite(b[2]) && isfinite(b[3]) && isfinite(logit);
    if (!finite) atomicOr(nonfinite_flag, 1);
    const float ih = img_hw[2 * n], iw = img_hw[2 * n + 1];
    b[0] = fminf(fmaxf(b[0], 0.f), iw);
    b[1] = fminf(fmaxf(b[1], 0.f), ih);
    b[2] = fminf(fmaxf(b[2], 0.f), iw);
    b[3] = fminf(fmaxf(b[3], 0.f), ih);
    const bool ok = finite && (b[2] - b[0] > min_size) && (b[3] - b[1] > min_size);
    // sigma of anchor index j (quirk)
    const int aj = j % A, locj = j / A;
    const int64_t rowj = static_cast<int64_t>(n) * H * Wp + (locj / W) * Wp + (locj % W);
    const float* sg = deltas + rowj * ld_delta + aj * 8 + 4;
    const float ssum = ((sigmoidf_(sg[0]) + sigmoidf_(sg[1])) + sigmoidf_(sg[2])) + sigmoidf_(sg[3]);
    const float sc = __fmul_rn(logit, 1.f - ssum / 4.0f);
    boxes[i] = make_float4(b[0], b[1], b[2], b[3]);
    scores[i] = sc;
    keys2[i] = ok ? order_desc(sc) : 0xFFFFFFFFu;
    vals2[i] = static_cast<uint32_t>(j);
    if (ok) atomicAdd(valid_count + n, 1);
  }
}

// ------------------------------------------------------------------------------------------
// NMS: 64x64 blocks of the (sorted) candidate list -> suppression bitmask, upper triangle only.
// The candidate rows are processed in BANDS (row blocks [row0_blk, row0_blk + gridDim.y)): the scan keeps at most
// max_keep boxes (2 000 of 12 000 RPN candidates, 100 of 16 000 detections), so the mask rows behind the point where
// the scan stops are never read. The host enqueues bitmask + scan per band; a band whose image already has max_keep
// survivors (keep_count, written by the previous band's scan) exits at once. Round 1 always built all cap^2 / 2
// tests (0.73 ms per step) and scanned one 18 MB mask per image.
__global__ void nms_bitmask_kernel(const float4* __restrict__ boxes, int64_t box_stride,
                                   const uint32_t* __restrict__ order, int64_t order_stride,
                                   const int* __restrict__ counts, int cap, int words, int wstride, float thr,
                                   int class_mod, unsigned long long* __restrict__ mask, int row0_blk,
                                   const int* __restrict__ keep_count, int max_keep) {
  const int cb = blockIdx.x, rb = row0_blk + blockIdx.y, n = blockIdx.z;
  if (cb < rb) return;
  if (row0_blk > 0 && keep_count[n] >= max_keep) return;  // this image is done
  int cnt = counts[n];
  if (cnt > cap) cnt = cap;
  if (rb * 64 >= cnt) return;
  __shared__ float4 cbox[64];
  __shared__ int ccls[64];
  const int t = threadIdx.x;
  const int cj = cb * 64 + t;
  if (cj < cnt) {
    const uint32_t oj = order[n * order_stride + cj];
    cbox[t] = boxes[n * box_stride + oj];
    ccls[t] = class_mod > 0 ? static_cast<int>(oj % class_mod) : 0;
  }
  __syncthreads();
  const int i = rb * 64 + t;
  if (i >= cnt) return;
  const uint32_t oi = order[n * order_stride + i];
  const float4 bi = boxes[n * box_stride + oi];
  const int ci = class_mod > 0 ? static_cast<int>(oi % class_mod) : 0;
  const float ai = __fmul_rn(bi.z - bi.x, bi.w - bi.y);
  unsigned long long bits = 0ull;
  const int jmax = min(64, cnt - cb * 64);
  const int j0 = (cb == rb) ? t + 1 : 0;
  for (int j = j0; j < jmax; ++j) {
    const float4 bj = cbox[j];
    const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
    const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
    const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
    const float inter = __fmul_rn(w, h);
    if (!(inter > 0.f) || ccls[j] != ci) continue;
    const float aj = __fmul_rn(bj.z - bj.x, bj.w - bj.y);
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    // decide without the division unless inter/uni is within ~1e-5 of the threshold (bit-exact result)
    const float tu = thr * uni;
    bool sup;
    if (inter > tu * 1.00001f) sup = true;
    else if (inter < tu * 0.99999f) sup = false;
    else sup = __fdiv_rn(inter, uni) > thr;
    if (sup) bits |= 1ull << j;
  }
  const int64_t band_rows = static_cast<int64_t>(gridDim.y) * 64;
  mask[(n * band_rows + (i - row0_blk * 64)) * wstride + cb] = bits;
}

// Sequential part of NMS, one CTA (1024 threads) per image. The mask rows of the NEXT 64-candidate
// block are prefetched into shared memory (one cp.async.bulk per row, mbarrier completion) while the
// current block is processed, so the per-block critical path only touches shared memory: one thread
// walks the survivors of the block (ffs over the not-yet-suppressed bits, so the cost is per kept
// box, bounded by max_keep overall), then all threads OR the surviving rows into the running
// `removed` bit vector. Stops after max_keep survivors. keep_idx = positions in the sorted order.
// PREFETCH = false (masks too wide for shared memory) reads the rows from global memory.
#ifdef NMS_DEBUG
__device__ long long g_nms_dbg[8];
#define DBG_T(k) if (tid == 0) { const long long t_ = clock64(); g_nms_dbg[k] += t_ - t_last; t_last = t_; }
#else
#define DBG_T(k)
#endif

template <bool PREFETCH>
__global__ void __launch_bounds__(1024, 1)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ counts, int cap, int words,
                int wpad, int max_keep, int* __restrict__ keep_idx, int* __restrict__ keep_count, int row0_blk,
                int band_blks, unsigned long long* __restrict__ removed_g) {
  extern __shared__ __align__(16) unsigned long long sm[];
  unsigned long long* removed = sm;                       // [wpad]
  unsigned long long* rowbuf = sm + wpad;                 // [2][64][wpad] when PREFETCH
  __shared__ uint64_t bar[2];
  __shared__ int s_rows[64];
  __shared__ int s_nk;
  __shared__ int s_total;
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int cnt = counts[n];
  if (cnt > cap) cnt = cap;
  // band-relative mask rows; the running state (suppressed bits, survivor count) lives in global memory between bands
  const unsigned long long* m = mask + (static_cast<int64_t>(n) * band_blks - row0_blk) * 64 * wpad;
  unsigned long long* rg = removed_g + static_cast<int64_t>(n) * wpad;
  const int total0 = row0_blk > 0 ? keep_count[n] : 0;
  if (total0 >= max_keep) return;  // (uniform) this image already has its max_keep survivors
  for (int w = tid; w < wpad; w += blockDim.x) removed[w] = row0_blk > 0 ? rg[w] : 0ull;
  if (tid == 0) {
    s_total = total0;
    ptb::mbar_init(&bar[0], 1);
    ptb::mbar_init(&bar[1], 1);
    ptb::fence_mbar_init();
  }
  const int nblk = min((cnt + 63) / 64, row0_blk + band_blks);
  __syncthreads();

  // warp 1: bulk-copies words [w0, wpad) (w0 even) of the 64 rows of block b into buffer `buf`
  auto prefetch = [&](int b, int buf) {
    if (b >= nblk) return;
    const int w0 = b & ~1;
    const uint32_t bytes = static_cast<uint32_t>(wpad - w0) * 8u;
    const int live = min(64, cnt - b * 64);
    if (lane == 0) ptb::mbar_arrive_expect_tx(&bar[buf], bytes * live);
    __syncwarp();
    for (int r = lane; r < live; r += 32)
      ptb::bulk_load_1d(rowbuf + (static_cast<int64_t>(buf) * 64 + r) * wpad + w0,
                        m + static_cast<int64_t>(b * 64 + r) * wpad + w0, bytes, &bar[buf]);
  };
  if (PREFETCH && warp == 1) prefetch(row0_blk, 0);
#ifdef NMS_DEBUG
  long long t_last = clock64();
#endif
  int b = row0_blk;
  for (; b < nblk; ++b) {
    if (s_total >= max_keep) break;
    const int cur = (b - row0_blk) & 1;
    if (PREFETCH) {
      ptb::mbar_wait(&bar[cur], ((b - row0_blk) >> 1) & 1);
      if (warp == 1) prefetch(b + 1, cur ^ 1);
    }
    DBG_T(0)
    const int live = min(64, cnt - b * 64);
    unsigned long long rem0 = removed[b];
    if (live < 64) rem0 |= ~0ull << live;
    if (rem0 != ~0ull) {
      if (tid == 0) {
        // one thread walks the survivors only: next = first zero bit of `rem` at or above i
        unsigned long long rem = rem0;
        const int total = s_total;
        int k = 0;
        unsigned long long avail = ~rem;
        while (avail) {
          const int i = __ffsll(static_cast<long long>(avail)) - 1;
          unsigned long long di;
          if (PREFETCH)
            di = rowbuf[(static_cast<int64_t>(cur) * 64 + i) * wpad + b];
          else
            di = m[static_cast<int64_t>(b * 64 + i) * wpad + b];
          s_rows[k] = i;
          if (total + k < max_keep) keep_idx[n * max_keep + total + k] = b * 64 + i;
          ++k;
          rem |= di | (1ull << i);
          avail = ~rem & ~((2ull << i) - 1ull);
        }
        s_nk = k;
        s_total = total + k;
      }
      DBG_T(3)
      __syncthreads();
      DBG_T(4)
      const int nk = s_nk;
      if (b + 1 < words) {
        const int g = tid >> 8;
        for (int w0 = b + 1 + (tid & 255); w0 < words; w0 += 256) {
          unsigned long long acc = 0ull;
          if (PREFETCH) {
            for (int j = g; j < nk; j += 4) acc |= rowbuf[(static_cast<int64_t>(cur) * 64 + s_rows[j]) * wpad + w0];
          } else {
            int j = g;
            for (; j + 12 < nk; j += 16) {
              const unsigned long long a0 = m[static_cast<int64_t>(b * 64 + s_rows[j]) * wpad + w0];
              const unsigned long long a1 = m[static_cast<int64_t>(b * 64 + s_rows[j + 4]) * wpad + w0];
              const unsigned long long a2 = m[static_cast<int64_t>(b * 64 + s_rows[j + 8]) * wpad + w0];
              const unsigned long long a3 = m[static_cast<int64_t>(b * 64 + s_rows[j + 12]) * wpad + w0];
              acc |= (a0 | a1) | (a2 | a3);
            }
            for (; j < nk; j += 4) acc |= m[static_cast<int64_t>(b * 64 + s_rows[j]) * wpad + w0];
          }
          if (acc) atomicOr(&removed[w0], acc);
        }
      }
      DBG_T(5)
    }
    __syncthreads();
    DBG_T(6)
#ifdef NMS_DEBUG
    if (tid == 0) g_nms_dbg[7] += 1;
#endif
  }
  // the last prefetch (if any) may still be in flight: drain it before the CTA exits
  if (PREFETCH && b < nblk) ptb::mbar_wait(&bar[(b - row0_blk) & 1], ((b - row0_blk) >> 1) & 1);
  __syncthreads();
  for (int w = tid; w < wpad; w += blockDim.x) rg[w] = removed[w];
  if (tid == 0) keep_count[n] = min(s_total, max_keep);
}

// proposals[n][i] = boxes[order[keep[i]]], logits likewise; rows >= keep_count zero-filled.
__global__ void rpn_gather_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, int k,
                                  const uint32_t* __restrict__ order, int64_t order_stride,
                                  const int* __restrict__ keep_idx, const int* __restrict__ keep_count,
                                  int max_keep, int N, float4* __restrict__ out_boxes, float* __restrict__ out_scores) {
  const int total = N * max_keep;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / max_keep, j = i - n * max_keep;
    if (j < keep_count[n]) {
      const uint32_t o = order[n * order_stride + keep_idx[i]];
      out_boxes[i] = boxes[n * k + o];
      out_scores[i] = scores[n * k + o];
    } else {
      out_boxes[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_scores[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Teacher pseudo-label filter, candidate stage: one thread per (image, roi, class).
__global__ void roi_infer_candidates_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                            const float4* __restrict__ props, const int* __restrict__ prop_count,
                                            int N, int cap, int K, const float* __restrict__ img_hw,
                                            float score_thresh, float wx, float wy, float ww, float wh,
                                            float4* __restrict__ cboxes, float* __restrict__ cscores,
                                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                            int* __restrict__ cand_count) {
  const int per_img = cap * K;
  const int total = N * per_img;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / per_img, rc = i - n * per_img;
    const int r = rc / K, c = rc - r * K;
    bool ok = r < prop_count[n];
    float sc = 0.f;
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
      const int64_t row = static_cast<int64_t>(n) * cap + r;
      const float* z = scores + row * (K + 1);
      float mx = z[0];
      for (int q = 1; q <= K; ++q) mx = fmaxf(mx, z[q]);
      float den = 0.f;
      for (int q = 0; q <= K; ++q) den += expf(z[q] - mx);
      const float prob = expf(z[c] - mx) / den;
      const float* d = deltas + row * (8 * K) + c * 8;
      const float4 p = props[row];
      const float pv[4] = {p.x, p.y, p.z, p.w};
      decode_box(pv, d[0], d[1], d[2], d[3], wx, wy, ww, wh, b);
      // validity of the roi row: every class box and every prob finite (fast_rcnn.py:66-71)
      bool fin = isfinite(prob);
      for (int q = 0; q < K && fin; ++q) {
        float t[4];
        const float* dq = deltas + row * (8 * K) + q * 8;
        decode_box(pv, dq[0], dq[1], dq[2], dq[3], wx, wy, ww, wh, t);
        fin = isfinite(t[0]) && isfinite(t[1]) && isfinite(t[2]) && isfinite(t[3]);
      }
      const float ih = img_hw[2 * n], iw = img_hw[2 * n + 1];
      b[0] = fminf(fmaxf(b[0], 0.f), iw);
      b[1] = fminf(fmaxf(b[1], 0.f), ih);
      b[2] = fminf(fmaxf(b[2], 0.f), iw);
      b[3] = fminf(fmaxf(b[3], 0.f), ih);
      ok = fin && (prob > score_thresh);
      const float ssum = ((sigmoidf_(d[4]) + sigmoidf_(d[5])) + sigmoidf_(d[6])) + sigmoidf_(d[7]);
      sc = __fmul_rn(prob, 1.f - ssum / 4.0f);
    }
    cboxes[i] = make_float4(b[0], b[1], b[2], b[3]);
    cscores[i] = sc;
    keys[i] = ok ? order_desc(sc) : 0xFFFFFFFFu;
    vals[i] = static_cast<uint32_t>(rc);
    if (ok) atomicAdd(cand_count + n, 1);
  }
}

__global__ void roi_infer_gather_kernel(const float4* __restrict__ cboxes, const float* __restrict__ cscores,
                                        const float* __restrict__ scores, const float* __restrict__ deltas,
                                        const uint32_t* __restrict__ order, const int* __restrict__ keep_idx,
                                        const int* __restrict__ keep_count, int N, int cap, int K, int topk,
                                        float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                                        int64_t* __restrict__ out_classes, float* __restrict__ out_logits,
                                        float* __restrict__ out_sigma, int* __restrict__ out_src_roi) {
  const int per_img = cap * K;
  const int total = N * topk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / topk, j = i - n * topk;
    if (j < keep_count[n]) {
      const uint32_t rc = order[static_cast<int64_t>(n) * per_img + keep_idx[i]];
      const int r = rc / K, c = rc - r * K;
      const int64_t row = static_cast<int64_t>(n) * cap + r;
      out_boxes[i] = cboxes[static_cast<int64_t>(n) * per_img + rc];
      out_scores[i] = cscores[static_cast<int64_t>(n) * per_img + rc];
      out_classes[i] = c;
      for (int q = 0; q <= K; ++q) out_logits[static_cast<int64_t>(i) * (K + 1) + q] = scores[row * (K + 1) + q];
      for (int q = 0; q < 4; ++q) out_sigma[i * 4 + q] = deltas[row * (8 * K) + c * 8 + 4 + q];
      out_src_roi[i] = r;
    } else {
      out_boxes[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_scores[i] = 0.f;
      out_classes[i] = 0;
      for (int q = 0; q <= K; ++q) out_logits[static_cast<int64_t>(i) * (K + 1) + q] = 0.f;
      for (int q = 0; q < 4; ++q) out_sigma[i * 4 + q] = 0.f;
      out_src_roi[i] = -1;
    }
  }
}

inline int grid1d(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)
#define LAUNCH_OK() static_cast<int>(cudaGetLastError())

extern "C" int ptb200_cell_anchors_from_wh(const float* wh, int num_cell, float* cell, void* stream) {
  cell_anchors_from_wh_kernel<<<1, 32, 0, STREAM>>>(wh, num_cell, cell);
  return LAUNCH_OK();
}

extern "C" int ptb200_anchor_grid(const float* cell, int num_cell, int h, int w, float stride, float offset,
                                  float* anchors, void* stream) {
  anchor_grid_kernel<<<grid1d(static_cast<int64_t>(h) * w * num_cell), 256, 0, STREAM>>>(
      cell, num_cell, h, w, stride, offset, reinterpret_cast<float4*>(anchors));
  return LAUNCH_OK();
}

extern "C" int ptb200_rpn_make_keys(const float* logits, int ld, int n, int h, int w, int num_cell, uint32_t* keys,
                                    uint32_t* vals, void* stream) {
  rpn_make_keys_kernel<<<grid1d(static_cast<int64_t>(n) * h * w * num_cell), 256, 0, STREAM>>>(logits, ld, n, h, w,
                                                                                              num_cell, keys, vals);
  return LAUNCH_OK();
}

extern "C" int ptb200_rpn_topk_decode(const uint32_t* sorted_idx, int64_t idx_stride, const float* logits,
                                      int ld_logit, const float* deltas, int ld_delta, const float* anchors, int n,
                                      int h, int w, int num_cell, int k, const float* img_hw, float min_size,
                                      float* boxes, float* scores, uint32_t* keys2, uint32_t* vals2,
                                      int* valid_count, int* nonfinite_flag, void* stream) {
  cudaMemsetAsync(valid_count, 0, sizeof(int) * n, STREAM);
  rpn_topk_decode_kernel<<<grid1d(static_cast<int64_t>(n) * k), 256, 0, STREAM>>>(
      sorted_idx, idx_stride, logits, ld_logit, deltas, ld_delta, reinterpret_cast<const float4*>(anchors), n, h, w,
      num_cell, k, img_hw, min_size, reinterpret_cast<float4*>(boxes), scores, keys2, vals2, valid_count,
      nonfinite_flag);
  return LAUNCH_OK();
}

// rows of the first band: enough for max_keep survivors when little is suppressed
static inline int nms_band0_blocks(int cap, int max_keep) {
  int rows = max_keep + max_keep / 4;
  if (rows < 1024) rows = 1024;
  if (rows > cap) rows = cap;
  return (rows + 63) / 64;
}

extern "C" int ptb200_nms(const float* boxes, int64_t box_stride, const uint32_t* order, int64_t order_stride,
                          const int* counts, int n, int cap, float thresh, int class_mod, int max_keep,
                          unsigned long long* mask_scratch, int* keep_idx, int* keep_count, void* stream) {
  if (n <= 0) return 0;
  const int words = (cap + 63) / 64;
  const int wpad = (words + 1) & ~1;  // row stride of the mask in 64-bit words (16-byte aligned rows)
  // bands: the first covers ~1.25 max_keep rows, the rest is cut into (at most) three equal bands
  const int b0 = nms_band0_blocks(cap, max_keep);
  int rest = (words - b0 + 2) / 3;
  if (rest < 1) rest = 1;
  const int band_max = b0 > rest ? b0 : rest;
  // scratch layout: [n][wpad] running `removed` bit vectors, then the band mask [n][band_max * 64][wpad]
  unsigned long long* removed_g = mask_scratch;
  unsigned long long* mask = mask_scratch + static_cast<size_t>(n) * wpad;
  const size_t smem_pf = (static_cast<size_t>(wpad) + 2ull * 64 * wpad) * sizeof(unsigned long long);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(nms_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    configured = true;
  }
  for (int row0 = 0; row0 < words;) {
    const int blks = row0 == 0 ? b0 : (rest < words - row0 ? rest : words - row0);
    dim3 grid(words, blks, n);
    nms_bitmask_kernel<<<grid, 64, 0, STREAM>>>(reinterpret_cast<const float4*>(boxes), box_stride, order,
                                               order_stride, counts, cap, words, wpad, thresh, class_mod, mask, row0,
                                               keep_count, max_keep);
    if (smem_pf <= 220 * 1024)
      nms_scan_kernel<true><<<n, 1024, smem_pf, STREAM>>>(mask, counts, cap, words, wpad, max_keep, keep_idx,
                                                         keep_count, row0, blks, removed_g);
    else
      nms_scan_kernel<false><<<n, 1024, wpad * sizeof(unsigned long long), STREAM>>>(
          mask, counts, cap, words, wpad, max_keep, keep_idx, keep_count, row0, blks, removed_g);
    row0 += blks;
  }
  (void)band_max;
  return LAUNCH_OK();
}

extern "C" int ptb200_rpn_gather(const float* boxes, const float* scores, int k, const uint32_t* order,
                                 int64_t order_stride, const int* keep_idx, const int* keep_count, int max_keep,
                                 int n, float* out_boxes, float* out_scores, void* stream) {
  rpn_gather_kernel<<<grid1d(static_cast<int64_t>(n) * max_keep), 256, 0, STREAM>>>(
      reinterpret_cast<const float4*>(boxes), scores, k, order, order_stride, keep_idx, keep_count, max_keep, n,
      reinterpret_cast<float4*>(out_boxes), out_scores);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_infer_candidates(const float* scores, const float* deltas, const float* props,
                                           const int* prop_count, int n, int cap, int num_classes,
                                           const float* img_hw, float score_thresh, const float* weights4,
                                           float* cboxes, float* cscores, uint32_t* keys, uint32_t* vals,
                                           int* cand_count, void* stream) {
  cudaMemsetAsync(cand_count, 0, sizeof(int) * n, STREAM);
  roi_infer_candidates_kernel<<<grid1d(static_cast<int64_t>(n) * cap * num_classes), 256, 0, STREAM>>>(
      scores, deltas, reinterpret_cast<const float4*>(props), prop_count, n, cap, num_classes, img_hw, score_thresh,
      weights4[0], weights4[1], weights4[2], weights4[3], reinterpret_cast<float4*>(cboxes), cscores, keys, vals,
      cand_count);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_infer_gather(const float* cboxes, const float* cscores, const float* scores,
                                       const float* deltas, const uint32_t* order, const int* keep_idx,
                                       const int* keep_count, int n, int cap, int num_classes, int topk,
                                       float* out_boxes, float* out_scores, int64_t* out_classes, float* out_logits,
                                       float* out_sigma, int* out_src_roi, void* stream) {
  roi_infer_gather_kernel<<<grid1d(static_cast<int64_t>(n) * topk), 256, 0, STREAM>>>(
      reinterpret_cast<const float4*>(cboxes), cscores, scores, deltas, order, keep_idx, keep_count, n, cap,
      num_classes, topk, reinterpret_cast<float4*>(out_boxes), out_scores, out_classes, out_logits, out_sigma,
      out_src_roi);
  return LAUNCH_OK();
}
