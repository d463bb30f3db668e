// First VGG convolution fused with image pre-processing:
//   uint8 CHW image -> (x - mean) / std -> 3x3 conv (3 -> 64) + bias + ReLU -> fp16 NHWC-flat.
// Replaces detectron2 `preprocess_image` (pt/modeling/meta_arch/rcnn.py:38-43) followed by
// vgg_block1.conv1 (pt/modeling/backbone/vgg.py:45-53,66-69). K = 27 is far too small to feed
// tcgen05 from an im2col matrix in HBM (that round trip cost 274 MB per image), so the im2col
// fragment is built in registers straight from the image (L1-resident neighbourhood) and the
// 128x64x32 product runs on the legacy mma.sync path; the kernel is bound by the 137 MB/image
// output write, which goes out as full 128-byte rows staged through shared memory.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ptb200.h"

namespace {

__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One warp = a 32-pixel segment of one image row x 64 output channels; 8 warps per CTA.
// The warp first stages the normalised 3 x 34 x 3 input patch in shared memory (coalesced byte loads,
// one conversion per element), then every lane assembles its mma.sync A fragments from the patch.
__global__ void __launch_bounds__(256, 2)
conv1_u8_kernel(const uint8_t* __restrict__ img, const int* __restrict__ hw, int N, int Hmax, int Wmax,
                int64_t img_stride, float m0, float m1, float m2, float is0, float is1, float is2,
                const __half* __restrict__ wpack /* [64][32] */, const float* __restrict__ bias,
                __half* __restrict__ out) {
  __shared__ __align__(16) __half stage[8][32][64 + 8];  // +8 halfs: conflict-free fragment stores
  __shared__ __half patch[8][3][34 * 3 + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = Wmax + 1;
  const int segs = (Wp + 31) / 32;
  const int64_t total_tiles = static_cast<int64_t>(N) * Hmax * segs;
  const int gq = lane >> 2, tq = lane & 3;

  uint32_t bf[8][2][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __half* wr = wpack + (nt * 8 + gq) * 32 + ks * 16 + tq * 2;
      bf[nt][ks][0] = *reinterpret_cast<const uint32_t*>(wr);
      bf[nt][ks][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  float bcol[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bcol[nt][0] = bias[nt * 8 + tq * 2];
    bcol[nt][1] = bias[nt * 8 + tq * 2 + 1];
  }
  // patch offsets of the 8 k indices this thread supplies: k = ks*16 + tq*2 + {0,1,8,9}
  int koff[8];  // dy * 104 + dx * 3 + c  (dx in 0..2 relative to pixel xi), -1 for the K padding
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = (j >> 2) * 16 + tq * 2 + (j & 1) + ((j >> 1) & 1) * 8;
    const int t = k / 3;
    koff[j] = k < 27 ? (t / 3) * (34 * 3 + 2) + (t % 3) * 3 + (k - t * 3) : -1;
  }
  const __half hzero = __float2half(0.f);

  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * 8 + warp; tile < total_tiles;
       tile += static_cast<int64_t>(gridDim.x) * 8) {
    const int seg = static_cast<int>(tile % segs);
    const int64_t ny = tile / segs;
    const int y = static_cast<int>(ny % Hmax);
    const int n = static_cast<int>(ny / Hmax);
    const int x0 = seg * 32;
    const int h = hw[2 * n], w = hw[2 * n + 1];
    const uint8_t* ib = img + n * img_stride;
    // ---- stage the patch: rows y-1..y+1, columns x0-1..x0+32, 3 channels interleaved
    __half* pw = &patch[warp][0][0];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      const bool yok = yy >= 0 && yy < h;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
        const float istd = c == 0 ? is0 : (c == 1 ? is1 : is2);
        const uint8_t* rowp = ib + (static_cast<int64_t>(c) * h + yy) * w;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int xi = lane + rep * 32;
          if (xi < 34) {
            const int xx = x0 + xi - 1;
            float v = 0.f;
            if (yok && xx >= 0 && xx < w) v = (static_cast<float>(__ldg(rowp + xx)) - mean) * istd;
            pw[dy * (34 * 3 + 2) + xi * 3 + c] = __float2half_rn(v);
          }
        }
      }
    }
    __syncwarp();
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t af[2][4];
#pragma unroll
      for (int half_ = 0; half_ < 2; ++half_) {
        const int xi = mt * 16 + gq + half_ * 8;  // pixel within the segment
        const __half* pp = pw + xi * 3;
        __half v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = koff[j] >= 0 ? pp[koff[j]] : hzero;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          __half2 lo = __halves2half2(v[ks * 4 + 0], v[ks * 4 + 1]);
          __half2 hi = __halves2half2(v[ks * 4 + 2], v[ks * 4 + 3]);
          af[ks][half_] = *reinterpret_cast<uint32_t*>(&lo);
          af[ks][2 + half_] = *reinterpret_cast<uint32_t*>(&hi);
        }
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma_16816(acc[mt][nt], af[ks], bf[nt][ks]);
    }
    // ---- epilogue: bias + ReLU; pixels at x >= Wmax (pad column) are written as zero
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half_ = 0; half_ < 2; ++half_) {
        const int r = mt * 16 + gq + half_ * 8;
        const bool live = (x0 + r) < Wmax;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          float a = fmaxf(acc[mt][nt][half_ * 2] + bcol[nt][0], 0.f);
          float b = fmaxf(acc[mt][nt][half_ * 2 + 1] + bcol[nt][1], 0.f);
          if (!live) a = b = 0.f;
          *reinterpret_cast<__half2*>(&stage[warp][r][nt * 8 + tq * 2]) = __floats2half2_rn(a, b);
        }
      }
    __syncwarp();
    const int64_t row_base = (static_cast<int64_t>(n) * Hmax + y) * Wp + x0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = i * 32 + lane;  // 256 chunks of 16 B: 32 rows x 8 chunks
      const int r = idx >> 3, c = idx & 7;
      if (x0 + r < Wp)
        *reinterpret_cast<uint4*>(out + (row_base + r) * 64 + c * 8) =
            *reinterpret_cast<const uint4*>(&stage[warp][r][c * 8]);
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" int ptb200_conv1_u8_f16(const uint8_t* images, const int* hw_dev, int n, int hmax, int wmax,
                                   int64_t image_stride, const float* mean3_host, const float* std3_host,
                                   const void* wpack_f16, const float* bias, void* out_f16, void* stream) {
  const int64_t tiles = static_cast<int64_t>(n) * hmax * ((wmax + 1 + 31) / 32);
  int64_t blocks = (tiles + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  conv1_u8_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, hw_dev, n, hmax, wmax, image_stride, mean3_host[0], mean3_host[1], mean3_host[2],
      1.f / std3_host[0], 1.f / std3_host[1], 1.f / std3_host[2], static_cast<const __half*>(wpack_f16), bias,
      static_cast<__half*>(out_f16));
  return static_cast<int>(cudaGetLastError());
}
