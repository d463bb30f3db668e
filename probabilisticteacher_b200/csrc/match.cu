// Box IoU matching, label assignment and (injected-randomness) sub-sampling, on device:
//   pairwise_iou + Matcher (+ low-quality matches)   detectron2 v0.5, called at
//       pt/modeling/proposal_generator/rpn.py:414-415 and pt/modeling/roi_heads/roi_heads.py:207-214
//   subsample_labels (positives first)               detectron2 v0.5, rpn.py:433 / roi_heads.py:223-225
//   add_ground_truth_to_proposals                    pt/modeling/proposal_generator/proposal_utils.py:157-224
//   _sample_proposals_unsup                          pt/modeling/roi_heads/roi_heads.py:257-291
// The M x R IoU matrix is never materialised. Sampling spec (shared with the oracle): the random
// permutation of the n candidates is the stable argsort of the first n entries of a priority vector.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

__device__ __forceinline__ float iou_pair(const float4 g, float ga, const float4 b) {
  // detectron2 pairwise_iou: inter / (area_gt + area_box - inter) if inter > 0 else 0
  const float w = fmaxf(__fsub_rn(fminf(g.z, b.z), fmaxf(g.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(g.w, b.w), fmaxf(g.y, b.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float ba = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(ga, ba), inter)) : 0.f;
}

__device__ __forceinline__ float box_area(const float4 g) {
  return __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
}

// pass 1: per box max IoU / first arg-max over the image's gt boxes; per gt best IoU, reduced
// warp -> shared memory -> one global atomicMax per (CTA, gt). grid = (chunks, N).
// boxes are shared by all images when box_img_stride == 0 (anchors).
constexpr int kMaxGt = 128;
__global__ void __launch_bounds__(256)
match_pass1_kernel(const float4* __restrict__ gt, const int* __restrict__ gt_count, int gt_cap,
                   const float4* __restrict__ boxes, int64_t box_img_stride,
                   const int* __restrict__ box_count, int R, int N, float* __restrict__ max_iou,
                   int* __restrict__ matched, int* __restrict__ best_per_gt) {
  __shared__ float4 sgt[kMaxGt];
  __shared__ float sarea[kMaxGt];
  __shared__ int sbest[kMaxGt];
  const int n = blockIdx.y;
  const int M = min(gt_count[n], gt_cap);
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const float4 g = gt[n * gt_cap + m];
    sgt[m] = g;
    sarea[m] = box_area(g);
    sbest[m] = 0;
  }
  __syncthreads();
  const int nbox = box_count == nullptr ? R : min(box_count[n], R);
  const int per_block = (R + gridDim.x - 1) / gridDim.x;
  const int r_begin = blockIdx.x * per_block;
  const int r_end = min(r_begin + per_block, R);
  const int span = (r_end - r_begin + 255) / 256 * 256;
  for (int rr = threadIdx.x; rr < span; rr += 256) {
    const int r = r_begin + rr;
    const bool live = r < r_end && r < nbox;
    float best = -1.f;
    int arg = 0;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) b = boxes[n * box_img_stride + r];
    for (int m = 0; m < M; ++m) {
      const float v = live ? iou_pair(sgt[m], sarea[m], b) : 0.f;
      if (live && v > best) {
        best = v;
        arg = m;
      }
      if (best_per_gt != nullptr) {
        const int wmax = __reduce_max_sync(0xffffffffu, __float_as_int(v));
        if ((threadIdx.x & 31) == 0 && wmax > 0) atomicMax(&sbest[m], wmax);
      }
    }
    if (r < r_end) {
      max_iou[static_cast<int64_t>(n) * R + r] = best;
      matched[static_cast<int64_t>(n) * R + r] = arg;
    }
  }
  __syncthreads();
  if (best_per_gt != nullptr)
    for (int m = threadIdx.x; m < M; m += blockDim.x)
      if (sbest[m] > 0) atomicMax(best_per_gt + n * gt_cap + m, sbest[m]);
}

// pass 2 (RPN): thresholds [lo, hi] -> {0, -1, 1}, then low-quality matches -> 1.
__global__ void rpn_label_kernel(const float4* __restrict__ gt, const int* __restrict__ gt_count, int gt_cap,
                                 const float4* __restrict__ anchors, int R, int N, const float* __restrict__ max_iou,
                                 const int* __restrict__ best_per_gt, float lo, float hi, int* __restrict__ labels) {
  const int64_t total = static_cast<int64_t>(N) * R;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / R);
    const int r = static_cast<int>(i - static_cast<int64_t>(n) * R);
    const int M = min(gt_count[n], gt_cap);
    int lab = 0;  // empty gt set: everything is background (Matcher, empty matrix)
    if (M > 0) {
      const float v = max_iou[i];
      lab = v >= hi ? 1 : (v >= lo ? -1 : 0);
      const float4 b = anchors[r];
      for (int m = 0; m < M; ++m) {
        const float4 g = gt[n * gt_cap + m];
        if (__float_as_int(iou_pair(g, box_area(g), b)) == best_per_gt[n * gt_cap + m]) {
          lab = 1;
          break;
        }
      }
    }
    labels[i] = lab;
  }
}

// ROI supervised: candidate i < prop_count is proposal i, then the image's gt boxes are appended.
// cls = gt class of the best gt if IoU >= thr, else K (background); no gt -> all background.
__global__ void roi_label_kernel(const float4* __restrict__ gt, const int* __restrict__ gt_classes,
                                 const int* __restrict__ gt_count, int gt_cap, const float4* __restrict__ props,
                                 const int* __restrict__ prop_count, int prop_cap, int N, int K, float thr,
                                 int* __restrict__ cls, int* __restrict__ matched, int* __restrict__ cand_count) {
  const int L = prop_cap + gt_cap;
  const int total = N * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / L, j = i - n * L;
    const int pc = min(prop_count[n], prop_cap), M = min(gt_count[n], gt_cap);
    if (j == 0) cand_count[n] = pc + M;
    int c = -2;  // beyond the candidate list
    int arg = 0;
    if (j < pc + M) {
      const float4 b = j < pc ? props[n * prop_cap + j] : gt[n * gt_cap + (j - pc)];
      float best = -1.f;
      for (int m = 0; m < M; ++m) {
        const float4 g = gt[n * gt_cap + m];
        const float v = iou_pair(g, box_area(g), b);
        if (v > best) {
          best = v;
          arg = m;
        }
      }
      c = (M > 0 && best >= thr) ? gt_classes[n * gt_cap + arg] : K;
    }
    cls[i] = c;
    matched[i] = arg;
  }
}

// Ordered compaction of the "positive" (v != -1 && v != bg && v != -2) and "negative" (v == bg)
// entries of each segment. One CTA per segment.
__global__ void __launch_bounds__(1024, 1)
compact_pos_neg_kernel(const int* __restrict__ labels, int64_t stride, const int* __restrict__ seg_len, int fixed_len,
                       int bg_label, int* __restrict__ pos_list, int* __restrict__ neg_list,
                       int* __restrict__ counts /* [seg][2] */) {
  __shared__ int wsum[2][32];
  __shared__ int base[2];
  const int seg = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = seg_len != nullptr ? min(seg_len[seg], fixed_len) : fixed_len;
  const int* lab = labels + seg * stride;
  if (tid < 2) base[tid] = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + tid;
    const int v = i < n ? lab[i] : -1;
    const bool isneg = (i < n) && (v == bg_label);
    const bool ispos = (i < n) && (v != -1) && (v != -2) && (v != bg_label);
    const unsigned bp = __ballot_sync(0xffffffffu, ispos), bn = __ballot_sync(0xffffffffu, isneg);
    const unsigned lt = (1u << lane) - 1u;
    if (lane == 0) {
      wsum[0][warp] = __popc(bp);
      wsum[1][warp] = __popc(bn);
    }
    __syncthreads();
    int offp = base[0], offn = base[1];
    for (int w = 0; w < warp; ++w) {
      offp += wsum[0][w];
      offn += wsum[1][w];
    }
    if (ispos) pos_list[seg * stride + offp + __popc(bp & lt)] = i;
    if (isneg) neg_list[seg * stride + offn + __popc(bn & lt)] = i;
    __syncthreads();
    if (tid == 0) {
      int tp = 0, tn = 0;
      for (int w = 0; w < 32; ++w) {
        tp += wsum[0][w];
        tn += wsum[1][w];
      }
      base[0] += tp;
      base[1] += tn;
    }
    __syncthreads();
  }
  if (tid == 0) {
    counts[2 * seg] = base[0];
    counts[2 * seg + 1] = base[1];
  }
}

// keys for the priority sort: segment 2*img = positives, 2*img+1 = negatives; length = counts.
__global__ void prio_keys_kernel(const float* __restrict__ prio_pos, const float* __restrict__ prio_neg,
                                 int64_t stride, const int* __restrict__ counts, int segs, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals) {
  const int64_t total = static_cast<int64_t>(segs) * stride;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int seg = static_cast<int>(i / stride);
    const int j = static_cast<int>(i - seg * stride);
    if (j < counts[seg]) {
      const float* p = (seg & 1) ? prio_neg : prio_pos;
      keys[i] = __float_as_uint(p[(seg >> 1) * stride + j]);
      vals[i] = static_cast<uint32_t>(j);
    }
  }
}

// RPN: write the sub-sampled label vector (-1 everywhere, 1 / 0 on the chosen anchors).
__global__ void rpn_sample_apply_kernel(const int* __restrict__ pos_list, const int* __restrict__ neg_list,
                                        const uint32_t* __restrict__ perm /* [2N][stride] */, int64_t stride,
                                        const int* __restrict__ counts, int N, int R, int batch_per_image,
                                        int max_pos, signed char* __restrict__ labels_out) {
  const int n = blockIdx.x;
  const int np = min(counts[2 * n], max_pos);
  const int nn = min(counts[2 * n + 1], batch_per_image - np);
  for (int i = threadIdx.x; i < R; i += blockDim.x) labels_out[static_cast<int64_t>(n) * R + i] = -1;
  __syncthreads();
  for (int i = threadIdx.x; i < np; i += blockDim.x)
    labels_out[static_cast<int64_t>(n) * R + pos_list[n * stride + perm[(2 * n) * stride + i]]] = 1;
  for (int i = threadIdx.x; i < nn; i += blockDim.x)
    labels_out[static_cast<int64_t>(n) * R + neg_list[n * stride + perm[(2 * n + 1) * stride + i]]] = 0;
}

// ROI supervised: emit the sampled rois (fg in permutation order, then bg), their classes and gt boxes.
__global__ void roi_sample_apply_kernel(const int* __restrict__ pos_list, const int* __restrict__ neg_list,
                                        const uint32_t* __restrict__ perm, int64_t stride,
                                        const int* __restrict__ counts, const int* __restrict__ cls,
                                        const int* __restrict__ matched, const float4* __restrict__ gt,
                                        const int* __restrict__ gt_count, int gt_cap, const float4* __restrict__ props,
                                        const int* __restrict__ prop_count, int prop_cap, int batch_per_image,
                                        int max_fg, int K, float4* __restrict__ out_rois, int* __restrict__ out_cls,
                                        float4* __restrict__ out_gt, int* __restrict__ out_count,
                                        int* __restrict__ out_src) {
  const int n = blockIdx.x;
  const int nf = min(counts[2 * n], max_fg);
  const int nb = min(counts[2 * n + 1], batch_per_image - nf);
  const int pc = min(prop_count[n], prop_cap), M = min(gt_count[n], gt_cap);
  if (threadIdx.x == 0) out_count[n] = nf + nb;
  for (int i = threadIdx.x; i < batch_per_image; i += blockDim.x) {
    const int o = n * batch_per_image + i;
    if (i < nf + nb) {
      const int j = i < nf ? pos_list[n * stride + perm[(2 * n) * stride + i]]
                           : neg_list[n * stride + perm[(2 * n + 1) * stride + (i - nf)]];
      out_rois[o] = j < pc ? props[n * prop_cap + j] : gt[n * gt_cap + (j - pc)];
      out_cls[o] = cls[n * stride + j];
      out_gt[o] = M > 0 ? gt[n * gt_cap + matched[n * stride + j]] : make_float4(0.f, 0.f, 0.f, 0.f);
      out_src[o] = j;
    } else {
      out_rois[o] = make_float4(0.f, 0.f, 1.f, 1.f);
      out_cls[o] = -1;
      out_gt[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_src[o] = -1;
    }
  }
}

// ROI unsupervised: keep, in order, every proposal whose best IoU with a pseudo box is >= thr and
// attach the matched pseudo box, teacher logits and teacher sigma logits. One CTA per image.
__global__ void __launch_bounds__(1024, 1)
roi_match_unsup_kernel(const float4* __restrict__ pseudo, const float* __restrict__ pseudo_logits,
                       const float* __restrict__ pseudo_sigma, const int* __restrict__ pseudo_count, int ps_cap,
                       const float4* __restrict__ props, const int* __restrict__ prop_count, int prop_cap, int K1,
                       float thr, float4* __restrict__ out_rois, float4* __restrict__ out_pseudo,
                       float* __restrict__ out_logits, float* __restrict__ out_sigma, int* __restrict__ out_count) {
  __shared__ int wsum[32];
  __shared__ int base;
  const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pc = min(prop_count[n], prop_cap), M = min(pseudo_count[n], ps_cap);
  if (tid == 0) base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < pc; b0 += 1024) {
    const int i = b0 + tid;
    bool keep = false;
    int arg = 0;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < pc) {
      b = props[n * prop_cap + i];
      float best = -1.f;
      for (int m = 0; m < M; ++m) {
        const float4 g = pseudo[n * ps_cap + m];
        const float v = iou_pair(g, box_area(g), b);
        if (v > best) {
          best = v;
          arg = m;
        }
      }
      keep = M > 0 && best >= thr;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (keep) {
      const int o = n * prop_cap + off + __popc(bal & ((1u << lane) - 1u));
      out_rois[o] = b;
      out_pseudo[o] = pseudo[n * ps_cap + arg];
      for (int q = 0; q < K1; ++q) out_logits[static_cast<int64_t>(o) * K1 + q] = pseudo_logits[(n * ps_cap + arg) * K1 + q];
      for (int q = 0; q < 4; ++q) out_sigma[o * 4 + q] = pseudo_sigma[(n * ps_cap + arg) * 4 + q];
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += wsum[w];
      base += t;
    }
    __syncthreads();
  }
  if (tid == 0) out_count[n] = base;
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

extern "C" int ptb200_rpn_match(const float* gt_boxes, const int* gt_count, int gt_cap, const float* anchors,
                                int num_anchors, int n, float iou_lo, float iou_hi, float* max_iou_scratch,
                                int* best_per_gt_scratch, int* matched_idx, int* labels, void* stream) {
  cudaMemsetAsync(best_per_gt_scratch, 0, sizeof(int) * n * gt_cap, STREAM);
  if (gt_cap > kMaxGt) return 1501;
  const int64_t total = static_cast<int64_t>(n) * num_anchors;
  int chunks = (num_anchors + 1023) / 1024;
  if (chunks < 1) chunks = 1;
  match_pass1_kernel<<<dim3(chunks, n), 256, 0, STREAM>>>(reinterpret_cast<const float4*>(gt_boxes), gt_count, gt_cap,
                                                       reinterpret_cast<const float4*>(anchors), 0, nullptr,
                                                       num_anchors, n, max_iou_scratch, matched_idx,
                                                       best_per_gt_scratch);
  rpn_label_kernel<<<grid1d(total), 256, 0, STREAM>>>(reinterpret_cast<const float4*>(gt_boxes), gt_count, gt_cap,
                                                     reinterpret_cast<const float4*>(anchors), num_anchors, n,
                                                     max_iou_scratch, best_per_gt_scratch, iou_lo, iou_hi, labels);
  return LAUNCH_OK();
}

extern "C" int ptb200_compact_pos_neg(const int* labels, int64_t stride, const int* seg_len, int fixed_len,
                                      int segments, int bg_label, int* pos_list, int* neg_list, int* counts,
                                      void* stream) {
  compact_pos_neg_kernel<<<segments, 1024, 0, STREAM>>>(labels, stride, seg_len, fixed_len, bg_label, pos_list,
                                                       neg_list, counts);
  return LAUNCH_OK();
}

extern "C" int ptb200_prio_keys(const float* prio_pos, const float* prio_neg, int64_t stride, const int* counts,
                                int n, uint32_t* keys, uint32_t* vals, void* stream) {
  prio_keys_kernel<<<grid1d(2 * n * stride), 256, 0, STREAM>>>(prio_pos, prio_neg, stride, counts, 2 * n, keys, vals);
  return LAUNCH_OK();
}

extern "C" int ptb200_rpn_sample_apply(const int* pos_list, const int* neg_list, const uint32_t* perm,
                                       int64_t stride, const int* counts, int n, int num_anchors,
                                       int batch_per_image, int max_pos, signed char* labels_out, void* stream) {
  rpn_sample_apply_kernel<<<n, 1024, 0, STREAM>>>(pos_list, neg_list, perm, stride, counts, n, num_anchors,
                                                 batch_per_image, max_pos, labels_out);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_label(const float* gt_boxes, const int* gt_classes, const int* gt_count, int gt_cap,
                                const float* props, const int* prop_count, int prop_cap, int n, int num_classes,
                                float iou_thr, int* cls, int* matched, int* cand_count, void* stream) {
  roi_label_kernel<<<grid1d(static_cast<int64_t>(n) * (prop_cap + gt_cap)), 256, 0, STREAM>>>(
      reinterpret_cast<const float4*>(gt_boxes), gt_classes, gt_count, gt_cap, reinterpret_cast<const float4*>(props),
      prop_count, prop_cap, n, num_classes, iou_thr, cls, matched, cand_count);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_sample_apply(const int* pos_list, const int* neg_list, const uint32_t* perm,
                                       int64_t stride, const int* counts, const int* cls, const int* matched,
                                       const float* gt_boxes, const int* gt_count, int gt_cap, const float* props,
                                       const int* prop_count, int prop_cap, int n, int batch_per_image, int max_fg,
                                       int num_classes, float* out_rois, int* out_cls, float* out_gt, int* out_count,
                                       int* out_src, void* stream) {
  roi_sample_apply_kernel<<<n, 512, 0, STREAM>>>(
      pos_list, neg_list, perm, stride, counts, cls, matched, reinterpret_cast<const float4*>(gt_boxes), gt_count,
      gt_cap, reinterpret_cast<const float4*>(props), prop_count, prop_cap, batch_per_image, max_fg, num_classes,
      reinterpret_cast<float4*>(out_rois), out_cls, reinterpret_cast<float4*>(out_gt), out_count, out_src);
  return LAUNCH_OK();
}

extern "C" int ptb200_roi_match_unsup(const float* pseudo_boxes, const float* pseudo_logits,
                                      const float* pseudo_sigma, const int* pseudo_count, int pseudo_cap,
                                      const float* props, const int* prop_count, int prop_cap, int n,
                                      int num_classes_plus1, float iou_thr, float* out_rois, float* out_pseudo,
                                      float* out_logits, float* out_sigma, int* out_count, void* stream) {
  roi_match_unsup_kernel<<<n, 1024, 0, STREAM>>>(
      reinterpret_cast<const float4*>(pseudo_boxes), pseudo_logits, pseudo_sigma, pseudo_count, pseudo_cap,
      reinterpret_cast<const float4*>(props), prop_count, prop_cap, num_classes_plus1, iou_thr,
      reinterpret_cast<float4*>(out_rois), reinterpret_cast<float4*>(out_pseudo), out_logits, out_sigma, out_count);
  return LAUNCH_OK();
}
