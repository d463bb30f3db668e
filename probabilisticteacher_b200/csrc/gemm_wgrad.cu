// Weight-gradient contraction on tcgen05 (sm_100a):
//   dW[m][t*n_total + n] += sum_b sum_p  G[b][p][m] * X[b][p + shift_t][n]
// G = gradient w.r.t. the layer output (fp16, [batch][rows][m_total]), X = layer input (fp16,
// [batch][rows][n_total]); both are "MN-major" operands (the reduction index p is the slow axis),
// so the UMMA descriptors use the MN-major SWIZZLE_128B canonical layout and TMA loads
// [64 k-rows][64 channels] boxes. Covers
//   * 3x3 conv weight gradients of VGG blocks 3-5 and the RPN conv (9 taps = 9 row shifts of X;
//     reference: autograd of pt/modeling/backbone/vgg.py:65-72 through cuDNN wgrad),
//   * fc1 / fc2 / predictor weight gradients (taps = 1).
// The reduction over (batch, rows) is split across CTAs (split-K); partial tiles are accumulated
// into the fp32 gradient arena with vectorised red.global.add (gradients of the two student
// forward passes of one step accumulate into the same arena, as autograd does for the reference).
#include "ptx.cuh"
#include <stdlib.h>
#include "gemm_tn.h"

namespace ptb {

static constexpr int WG_BM = 128;   // output-channel tile (UMMA M)
static constexpr int WG_THREADS = 192;

struct WgCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};

struct WgParams {
  int batch, rows, m_total, n_total, bn, taps;
  int shifts[9];
  int ksplit, stages;
  float* out;
  int64_t ld_out;
  float scale;
  float* bias_out;  // optional: bias_out[m] += scale * sum_{b,p} G[b][p][m] (fused bias gradient)
  const int* seg_counts;  // optional segment mode (batch == 1): 64-row chunks without a live row are skipped
  int seg_cap;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ bool chunk_live(const int* __restrict__ seg_counts, int seg_cap, int r0, int rows, int bk) {
  if (seg_counts == nullptr) return true;
  int r = r0;
  const int rend = min(r0 + bk, rows);
  while (r < rend) {
    const int n = r / seg_cap;
    if (r - n * seg_cap < min(seg_counts[n], seg_cap)) return true;
    r = (n + 1) * seg_cap;
  }
  return false;
}

// X3 = f16x3 (fp32-equivalent) precision: G and X are [hi | lo | hi] triples; a stage holds the hi AND lo tiles of both
// operands and the accumulator receives Gh'Xh + Gl'Xh + Gh'Xl (12 MMAs per 64-row chunk instead of 4). One fused
// launch instead of three passes of the plain kernel over column slices: a stage moves 2x the bytes for 3x the
// MMAs, which lifts the kernel off its L2 -> shared-memory bound (48 KB per 512 tensor cycles at BN = 256).
// WG_BK = reduction rows per stage (one [WG_BK rows][64 ch] box per 64-channel block): 64 for the fp16 kernel; the X3
// kernel runs 32-row stages -- four 48 KB stages instead of two of 96 KB, same bytes in flight but every load is
// issued a stage-time earlier (measured per step: 7.83 -> 7.24 ms, 1 152 -> 1 243 TFLOP/s; PTB200_WG_BK=64 restores).
// MT = 2 (X3 only, m_total % 256 == 0): a CTA owns TWO 128-channel output tiles (two TMEM accumulators of bn columns)
// that share every X tile -- a 32-row stage moves 64 KB for 24 MMAs instead of 48 KB for 12: a third fewer bytes
// through L2 per MMA (ncu: this kernel sat at the ~12 TB/s LTS cap, 12.8 TB/s time-weighted).
template <bool X3, int WG_BK, int MT>
__global__ void __launch_bounds__(WG_THREADS, 2)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x,
                  const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int WG_BLK_BYTES = WG_BK * 128;  // one [WG_BK rows][64 ch] box
  const int bn = p.bn;
  constexpr int half_bytes = (WG_BM / 64) * WG_BLK_BYTES;  // one 128-channel G tile: 16 KB at WG_BK = 64
  const int a_bytes = MT * half_bytes;
  const int b_bytes = (bn / 64) * WG_BLK_BYTES;
  // stage layout: [Gh][Gl (X3)][Xh][Xl (X3)]
  const int stage_bytes = (X3 ? 2 : 1) * (a_bytes + b_bytes);
  WgCtl* ctl = reinterpret_cast<WgCtl*>(smem + p.stages * stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode: blockIdx.x = ((tap * m_tiles + mt) * n_tiles + nt) * ksplit + ks
  const int m_tiles = p.m_total / (WG_BM * MT);
  const int n_tiles = p.n_total / bn;
  int id = blockIdx.x;
  const int ks = id % p.ksplit;
  id /= p.ksplit;
  const int nt = id % n_tiles;
  id /= n_tiles;
  const int mt = id % m_tiles;
  const int tap = id / m_tiles;

  // the CTAs of tap 0 / N-tile 0 also reduce the bias gradient: their (otherwise idle) epilogue
  // warps sum the staged G tile column-wise while the MMA warp consumes it
  const bool do_bias = p.bias_out != nullptr && tap == 0 && nt == 0;
  const int chunks_per_img = (p.rows + WG_BK - 1) / WG_BK;
  const int total_chunks = chunks_per_img * p.batch;
  const int c_begin = static_cast<int>((static_cast<int64_t>(total_chunks) * ks) / p.ksplit);
  const int c_end = static_cast<int>((static_cast<int64_t>(total_chunks) * (ks + 1)) / p.ksplit);
  int k_iters = c_end - c_begin;
  if (p.seg_counts != nullptr) {  // count the live chunks of this CTA's range (every role does the same)
    k_iters = 0;
    for (int c = c_begin; c < c_end; ++c) k_iters += chunk_live(p.seg_counts, p.seg_cap, c * WG_BK, p.rows, WG_BK) ? 1 : 0;
  }

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(MT * bn)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_g);
    tma_prefetch_desc(&map_x);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], do_bias ? 5 : 1);  // MMA commit (+ one arrival per epilogue warp)
    }
    mbar_init(&ctl->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&ctl->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (k_iters > 0) {
    if (warp == 0) {
      // the whole warp walks the loop (converged control flow keeps the TMA operands in uniform registers:
      // no R2UR.BROADCAST waterfall per load); lane 0 issues
      int s = 0;
      uint32_t ph = 0;
      const int shift = p.shifts[tap];
      for (int c = c_begin; c < c_end; ++c) {
        const int b = c / chunks_per_img;
        const int r0 = (c - b * chunks_per_img) * WG_BK;
        if (!chunk_live(p.seg_counts, p.seg_cap, r0, p.rows, WG_BK)) continue;
        mbar_wait(&ctl->empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + (X3 ? 2 : 1) * a_bytes;
          mbar_arrive_expect_tx(&ctl->full[s], stage_bytes);
          tma_load_4d(sa, &map_g, &ctl->full[s], 0, r0, mt * (MT * WG_BM / 64), b);
          tma_load_4d(sb, &map_x, &ctl->full[s], 0, r0 + shift, nt * (bn / 64), b);
          if (X3) {  // the lo halves: channel blocks [m_total/64, 2 m_total/64) and [n_total/64, 2 n_total/64)
            tma_load_4d(sa + a_bytes, &map_g, &ctl->full[s], 0, r0, p.m_total / 64 + mt * (MT * WG_BM / 64), b);
            tma_load_4d(sb + b_bytes, &map_x, &ctl->full[s], 0, r0 + shift, p.n_total / 64 + nt * (bn / 64), b);
          }
        }
        __syncwarp();
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = umma_idesc_f16(WG_BM, bn, 1, 1);
      int s = 0;
      uint32_t ph = 0;
      // provably warp-uniform TMEM address (see gemm_tn.cu): avoids a R2UR.BROADCAST waterfall per MMA
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      for (int ki = 0; ki < k_iters; ++ki) {
        mbar_wait(&ctl->full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = __shfl_sync(0xffffffffu, smem_u32(smem + s * stage_bytes), 0);
        const uint32_t b_addr = __shfl_sync(0xffffffffu, a_addr + (X3 ? 2 : 1) * a_bytes, 0);
        if (elect_one()) {
          // MN-major SW128: LBO = stride between 64-channel blocks, SBO = stride between 8-row groups
          const uint64_t db = umma_desc_sw128(b_addr, WG_BLK_BYTES, 1024);
          const uint64_t dbl = umma_desc_sw128(b_addr + b_bytes, WG_BLK_BYTES, 1024);  // Xl (X3)
#pragma unroll
          for (int h = 0; h < MT; ++h) {  // output tile h: G channel blocks [2h, 2h + 2) of the stage, accumulator h
            const uint64_t da = umma_desc_sw128(a_addr + h * half_bytes, WG_BLK_BYTES, 1024);
            const uint32_t d_tmem = tmem_base_u + h * bn;
#pragma unroll
            for (int k = 0; k < WG_BK / 16; ++k) {
              // 16 reduction rows = 2048 B further into each block
              umma_f16_ss(d_tmem, da + (2048 >> 4) * k, db + (2048 >> 4) * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
            }
            if (X3) {
              const uint64_t dal = umma_desc_sw128(a_addr + a_bytes + h * half_bytes, WG_BLK_BYTES, 1024);  // Gl
#pragma unroll
              for (int k = 0; k < WG_BK / 16; ++k) umma_f16_ss(d_tmem, dal + (2048 >> 4) * k, db + (2048 >> 4) * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < WG_BK / 16; ++k) umma_f16_ss(d_tmem, da + (2048 >> 4) * k, dbl + (2048 >> 4) * k, idesc, 1u);
            }
          }
          umma_commit(&ctl->empty[s]);
          if (ki == k_iters - 1) umma_commit(&ctl->acc_full);
        }
        __syncwarp();
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      if (do_bias) {
        // thread t of the 128 epilogue threads: column pair (2*m2, 2*m2+1) of the 128-wide G tile
        // ([blk = col / 64][k row][64 ch, SW128]) over one half of the 64 staged rows; 4 independent
        // fp32 partial sums keep the loop off the critical path of the MMA pipeline
        const int t = (warp - 2) * 32 + lane;
        const int m2 = t & 63, half_ = t >> 6;
        const int colp = 2 * m2;
        const int blk = colp >> 6, col = colp & 63;
        float s0[MT], s1[MT], s2[MT], s3[MT];
#pragma unroll
        for (int h = 0; h < MT; ++h) s0[h] = s1[h] = s2[h] = s3[h] = 0.f;
        int s = 0;
        uint32_t ph = 0;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&ctl->full[s], ph);
#pragma unroll
          for (int h = 0; h < MT; ++h) {  // output tile h: channel blocks [2h, 2h + 2) of the staged G tile
            const uint8_t* g = smem + s * stage_bytes + h * half_bytes + blk * WG_BLK_BYTES + (col & 7) * 2;
#pragma unroll
            for (int part = 0; part < (X3 ? 2 : 1); ++part) {  // X3: Gh then Gl (sum of both = the fp32 gradient)
              const uint8_t* gp = g + part * a_bytes;
#pragma unroll
              for (int i = 0; i < WG_BK / 2; i += 2) {
                const int kr0 = half_ * (WG_BK / 2) + i, kr1 = kr0 + 1;
                const __half2 a = *reinterpret_cast<const __half2*>(gp + kr0 * 128 + (((col >> 3) ^ (kr0 & 7)) << 4));
                const __half2 b = *reinterpret_cast<const __half2*>(gp + kr1 * 128 + (((col >> 3) ^ (kr1 & 7)) << 4));
                const float2 fa = __half22float2(a), fb = __half22float2(b);
                s0[h] += fa.x;
                s1[h] += fa.y;
                s2[h] += fb.x;
                s3[h] += fb.y;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->empty[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
#pragma unroll
        for (int h = 0; h < MT; ++h) {
          atomicAdd(p.bias_out + (mt * MT + h) * WG_BM + colp, (s0[h] + s2[h]) * p.scale);
          atomicAdd(p.bias_out + (mt * MT + h) * WG_BM + colp + 1, (s1[h] + s3[h]) * p.scale);
        }
      }
      mbar_wait(&ctl->acc_full, 0);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < MT; ++h) {
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * bn;
        float* orow = p.out + static_cast<int64_t>((mt * MT + h) * WG_BM + r) * p.ld_out +
                      static_cast<int64_t>(tap) * p.n_total + nt * bn;
        for (int c0 = 0; c0 < bn; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(orow + c0 + 4 * j, __uint_as_float(v[4 * j]) * p.scale,
                       __uint_as_float(v[4 * j + 1]) * p.scale, __uint_as_float(v[4 * j + 2]) * p.scale,
                       __uint_as_float(v[4 * j + 3]) * p.scale);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

static int g_wg_sms = 0;

// G: [batch][rows][ldg] fp16 (m_total channels used), X: [batch][rows][ldx] fp16 (n_total used)
int gemm_wgrad_launch(const void* G, int64_t ldg, int64_t g_batch_stride, const void* X, int64_t ldx,
                      int64_t x_batch_stride, int batch, int rows, int m_total, int n_total, int taps,
                      const int* shifts, float* out, int64_t ld_out, float scale, int ksplit, float* bias_out,
                      const int* seg_counts, int seg_cap, cudaStream_t stream, bool x3) {
  if (seg_counts != nullptr && (batch != 1 || seg_cap <= 0)) return 1103;
  if (m_total % WG_BM != 0 || n_total % 64 != 0 || taps < 1 || taps > 9) return 1101;
  int bn = 256;  // (x3 stages hold hi + lo tiles of both operands: 48 KB per 32-row stage at BN = 256, four stages)
  while (n_total % bn != 0) bn >>= 1;
  if (bn < 64) return 1102;
  if (x3 && (ldg < 3 * m_total || ldx < 3 * n_total)) return 1104;
  if (g_wg_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static int bk_opt = -1;
  if (bk_opt < 0) {
    const char* e = getenv("PTB200_WG_BK");
    bk_opt = (e && atoi(e) == 64) ? 64 : 32;
  }
  const int WG_BK = x3 ? bk_opt : 64;
  const int WG_BLK_BYTES = WG_BK * 128;
  // two output tiles per CTA (see the kernel): f16x3 with 32-row stages, 256 output channels per tile pair and an X
  // tile of at most 256 columns (2 x 256 TMEM columns). PTB200_WG_MT=1 keeps one tile per CTA.
  static int mt_opt = -1;
  if (mt_opt < 0) {
    const char* e = getenv("PTB200_WG_MT");
    mt_opt = (e && atoi(e) == 1) ? 1 : 2;
  }
  const int MT = (x3 && WG_BK == 32 && mt_opt == 2 && m_total % (2 * WG_BM) == 0) ? 2 : 1;
  CUtensorMap mg, mx;
  {
    uint64_t dims[4] = {64, (uint64_t)rows, (uint64_t)((x3 ? 3 : 1) * m_total / 64), (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)ldg * 2, 128, (uint64_t)g_batch_stride * 2};
    uint32_t box[4] = {64, (uint32_t)WG_BK, (uint32_t)(MT * WG_BM / 64), 1};
    if (make_tmap_f16(&mg, G, 4, dims, str, box)) return 1110;
  }
  {
    uint64_t dims[4] = {64, (uint64_t)rows, (uint64_t)((x3 ? 3 : 1) * n_total / 64), (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)ldx * 2, 128, (uint64_t)x_batch_stride * 2};
    uint32_t box[4] = {64, (uint32_t)WG_BK, (uint32_t)(bn / 64), 1};
    if (make_tmap_f16(&mx, X, 4, dims, str, box)) return 1111;
  }
  WgParams p;
  p.batch = batch;
  p.rows = rows;
  p.m_total = m_total;
  p.n_total = n_total;
  p.bn = bn;
  p.taps = taps;
  for (int i = 0; i < 9; ++i) p.shifts[i] = (shifts != nullptr && i < taps) ? shifts[i] : 0;
  p.out = out;
  p.ld_out = ld_out;
  p.scale = scale;
  p.bias_out = bias_out;
  p.seg_counts = seg_counts;
  p.seg_cap = seg_cap;
  const int tiles = taps * (m_total / (WG_BM * MT)) * (n_total / bn);
  const int total_chunks = ((rows + WG_BK - 1) / WG_BK) * batch;
  // a caller's split count is meant for 128-channel tiles: twice the splits over half as many (double) tiles keeps the
  // CTA count and the work per CTA
  if (MT == 2 && ksplit > 0) ksplit = ksplit * 2 <= total_chunks ? ksplit * 2 : ksplit;
  if (ksplit <= 0) {
    ksplit = g_wg_sms / tiles;                       // aim at one full wave of CTAs
    const int per_cta = 2048 / WG_BK;                          // at least 2048 reduction rows per CTA
    const int max_split = (total_chunks + per_cta - 1) / per_cta;
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
  }
  p.ksplit = ksplit;
  const int stage_bytes = (x3 ? 2 : 1) * (MT * WG_BM / 64 + bn / 64) * WG_BLK_BYTES;
  int stages = (232448 - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  static int stage_cap = -1;  // experiment knob: PTB200_WG_STAGES=2 lets two CTAs share an SM
  if (stage_cap < 0) {
    const char* e = getenv("PTB200_WG_STAGES");
    stage_cap = e ? atoi(e) : 0;
  }
  if (stage_cap > 0 && stages > stage_cap) stages = stage_cap;
  // many tiles with a short reduction (fc1: 784 tiles x 63 chunks): two 2-stage CTAs per SM, so that one CTA's
  // prologue / red.add epilogue overlaps the other's main loop (fc1 wgrad 229 -> 160 us); long reductions
  // (conv layers) keep the deep single-CTA pipeline
  if (!x3 && stage_cap == 0 && tiles * ksplit >= 2 * g_wg_sms && total_chunks / ksplit <= 128 && stages > 2) stages = 2;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + (int)sizeof(WgCtl) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel<false, 64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_wgrad_kernel<true, 64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_wgrad_kernel<true, 32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_wgrad_kernel<true, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (x3 && WG_BK == 32 && MT == 2)
    gemm_wgrad_kernel<true, 32, 2><<<tiles * ksplit, WG_THREADS, smem_bytes, stream>>>(mg, mx, p);
  else if (x3 && WG_BK == 32)
    gemm_wgrad_kernel<true, 32, 1><<<tiles * ksplit, WG_THREADS, smem_bytes, stream>>>(mg, mx, p);
  else if (x3)
    gemm_wgrad_kernel<true, 64, 1><<<tiles * ksplit, WG_THREADS, smem_bytes, stream>>>(mg, mx, p);
  else
    gemm_wgrad_kernel<false, 64, 1><<<tiles * ksplit, WG_THREADS, smem_bytes, stream>>>(mg, mx, p);
  return (int)cudaGetLastError();
}

}  // namespace ptb
