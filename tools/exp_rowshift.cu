// Hardware-behaviour probe (not on the product path): does a UMMA SWIZZLE_128B K-major operand
// descriptor whose start address is offset by s*128 B (s rows, not 1024-byte aligned) read rows
// s..s+127 of a TMA-written tile correctly? Decides whether the 3 horizontal taps of a 3x3 conv can
// share one staged row window. out[s][128][64] fp32 = A[s:s+128] * B^T for s = 0..7.
#include "ptx.cuh"

namespace ptb {
int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

__global__ void __launch_bounds__(128, 1)
exp_rowshift_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    float* out, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;                 // 136 rows x 128 B = 17408
  uint8_t* sb = smem + 17408;         // 64 rows x 128 B
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full, 17408 + 8192);
    tma_load_2d(sa, &map_a, &full, 0, 0);
    tma_load_2d(sb, &map_b, &full, 0, 0);
  }
  for (int s = 0; s < 8; ++s) {
    if (warp == 0) {
      mbar_wait(&full, 0);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_f16(128, 64, 0, 0);
        uint64_t da = umma_desc_sw128(smem_u32(sa) + s * 128, 16, 1024);
        if (use_base_offset) da |= static_cast<uint64_t>(s & 7) << 49;
        const uint64_t db = umma_desc_sw128(smem_u32(sb), 16, 1024);
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0);
        umma_commit(&done);
      }
      __syncwarp();
    }
    mbar_wait(&done, s & 1);
    tc_fence_after();
    const int r = warp * 32 + lane;
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(s * 128 + r) * 64 + j] = __uint_as_float(v[j]);
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(s * 128 + r) * 64 + 32 + j] = __uint_as_float(v[j]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}
}  // namespace ptb

extern "C" int ptb200_exp_rowshift(const void* A /* [136][64] fp16 */, const void* B /* [64][64] fp16 */,
                                   float* out /* [8][128][64] */, int use_base_offset, void* stream) {
  using namespace ptb;
  CUtensorMap ma, mb;
  uint64_t da[2] = {64, 136};
  uint64_t sa_[1] = {128};
  uint32_t ba[2] = {64, 136};
  if (make_tmap_f16(&ma, A, 2, da, sa_, ba)) return 1;
  uint64_t db[2] = {64, 64};
  uint32_t bb[2] = {64, 64};
  if (make_tmap_f16(&mb, B, 2, db, sa_, bb)) return 2;
  cudaFuncSetAttribute(exp_rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  exp_rowshift_kernel<<<1, 128, 17408 + 8192 + 1024, static_cast<cudaStream_t>(stream)>>>(ma, mb, out, use_base_offset);
  return static_cast<int>(cudaGetLastError());
}
