// Flat-arena parameter updates (HBM-bound streaming kernels, float4 vectorised):
//   EMA teacher update              pt/engine/trainer.py:431-449
//   global grad-norm clip + SGD     pt/engine/trainer.py:592-603 + torch.optim.SGD(momentum, weight_decay)
//                                   (detectron2 build_optimizer), pt/engine/trainer.py:383-386
// The clip coefficient is computed on device from the squared-norm scalar, so the step has no host sync.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

// keep_dev != nullptr: the keep rate is read from device memory (CUDA-graph replays of the step: the host writes
// the EMA keep rate on update iterations and 1.0 on the others, pt/engine/trainer.py:296-298; keep == 1 leaves the
// teacher untouched, bit for bit)
__global__ void ema_kernel(float* __restrict__ teacher, const float* __restrict__ student, int64_t n, float keep,
                           const float* __restrict__ keep_dev) {
  if (keep_dev != nullptr) {
    keep = keep_dev[0];
    if (keep == 1.f) return;
  }
  const float ks = 1.f - keep;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 t = reinterpret_cast<float4*>(teacher)[i];
    const float4 s = reinterpret_cast<const float4*>(student)[i];
    t.x = s.x * ks + t.x * keep;
    t.y = s.y * ks + t.y * keep;
    t.z = s.z * ks + t.z * keep;
    t.w = s.w * ks + t.w * keep;
    reinterpret_cast<float4*>(teacher)[i] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    teacher[i] = student[i] * ks + teacher[i] * keep;
  }
}

// Deterministic squared norm: every replica must derive the SAME clip coefficient from the (bit-identical)
// all-reduced gradients, or the data-parallel replicas drift apart in the last bits (an atomicAdd of the block
// partials in arrival order did exactly that: tools/check_ddp.py). Block partials go to a scratch array; the
// last block to finish adds them in index order with a fixed-shape tree.
__device__ float g_sumsq_partial[148 * 16];
__device__ unsigned int g_sumsq_done = 0;

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float pre_scale, float* __restrict__ out) {
  __shared__ float sm[32];
  __shared__ bool is_last;
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  acc *= pre_scale * pre_scale;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) {
      g_sumsq_partial[blockIdx.x] = acc;
      __threadfence();
      is_last = atomicAdd(&g_sumsq_done, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float tot = 0.f;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) tot += __ldcg(&g_sumsq_partial[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x < 32) {
    tot = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (threadIdx.x == 0) {
      out[0] = tot;
      g_sumsq_done = 0;
    }
  }
}

// g' = pre_scale * g * clip/max(norm, clip) ; g' += wd * p ; m = mu * m + g' ; p -= lr * m
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, int64_t n,
                           float lr, float mu, float wd, float clip, float pre_scale,
                           const float* __restrict__ sumsq) {
  float coef = pre_scale;
  if (sumsq != nullptr) {
    const float nrm = sqrtf(sumsq[0]);
    coef *= clip / fmaxf(nrm, clip);
  }
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    mv.x = mu * mv.x + (gv.x * coef + wd * pv.x);
    mv.y = mu * mv.y + (gv.y * coef + wd * pv.y);
    mv.z = mu * mv.z + (gv.z * coef + wd * pv.z);
    mv.w = mu * mv.w + (gv.w * coef + wd * pv.w);
    pv.x -= lr * mv.x;
    pv.y -= lr * mv.y;
    pv.z -= lr * mv.z;
    pv.w -= lr * mv.w;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float mm = mu * m[i] + (g[i] * coef + wd * p[i]);
    m[i] = mm;
    p[i] -= lr * mm;
  }
}

inline int grid_stream(int64_t n) {
  int64_t g = (n / 4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int ptb200_ema_update(float* teacher, const float* student, int64_t n, float keep_rate, void* stream) {
  ema_kernel<<<grid_stream(n), 256, 0, STREAM>>>(teacher, student, n, keep_rate, nullptr);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_ema_update_dev(float* teacher, const float* student, int64_t n, const float* keep_rate_dev,
                                     void* stream) {
  ema_kernel<<<grid_stream(n), 256, 0, STREAM>>>(teacher, student, n, 1.f, keep_rate_dev);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_grad_sumsq(const float* grads, int64_t n, float pre_scale, float* sumsq_out, void* stream) {
  sumsq_kernel<<<grid_stream(n), 256, 0, STREAM>>>(grads, n, pre_scale, sumsq_out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ptb200_clip_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr,
                                    float momentum, float weight_decay, float clip_norm, float pre_scale,
                                    const float* sumsq_dev, void* stream) {
  sgd_kernel<<<grid_stream(n), 256, 0, STREAM>>>(params, grads, momentum_buf, n, lr, momentum, weight_decay,
                                                clip_norm, pre_scale, sumsq_dev);
  return static_cast<int>(cudaGetLastError());
}
