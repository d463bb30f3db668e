// Implicit-GEMM "TN" kernel for sm_100a: D[b][p][n] = sum_t sum_k A[b][p + shift_t][k] * B[n][t*K + k]
//
// One kernel covers every dense contraction on the Probabilistic Teacher hot path whose
// operands are K-major:
//   * 3x3 conv forward over NHWC activations stored as flattened, right-padded rows
//     ([N][H*Wp][C], Wp = W+1, pad column kept at zero) -- the 9 taps are 9 row shifts of
//     the same TMA box, top/bottom padding comes from TMA out-of-bounds zero fill
//     (reference: pt/modeling/backbone/vgg.py:45-53,65-72 via detectron2 Conv2d -> cuDNN);
//   * 3x3 conv data-gradient (same kernel, weights pre-flipped/transposed);
//   * RPN 3x3 conv + 1x1 objectness/(mu,sigma) heads (pt/modeling/proposal_generator/rpn.py:44-55,96);
//   * box head fc1/fc2/predictor and their data-gradients (taps = 1)
//     (pt/modeling/roi_heads/roi_heads.py:127-128, fast_rcnn.py:157-169).
//
// Structure: persistent CTAs (one per SM), warp-specialised:
//   warp 0    : TMA producer (A tile 128 rows x 64 k, B tile BN rows x 64 k, SWIZZLE_128B)
//   warp 1    : tcgen05.mma issuer (M=128, N=BN, K=16 per instruction, fp32 accumulators in TMEM,
//               two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2-9 : epilogue (tcgen05.ld -> bias/ReLU/pad-mask -> fp16 -> swizzled smem -> TMA store,
//               or fp32 direct stores for the narrow head outputs); two warps per TMEM lane quarter
#include "ptx.cuh"
#include "gemm_tn.h"
#include <stdlib.h>

namespace ptb {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
static constexpr int STAGING_BYTES = BM * 128;     // one 128 x 64 fp16 output chunk
static constexpr int NUM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
static constexpr int EPI_THREADS = 256;

struct SmemCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t aux_full[2];
  uint64_t bres_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// ROWWIN (3x3 convs with small Cout, where the tile is L2-bandwidth bound): a pipeline stage holds one
// 136-row window of A (rows p0 + (ky-1)*Wp - 1 ...) shared by the three horizontal taps kx = 0..2 --
// the taps are UMMA descriptors whose start address is offset by kx*128 B (SWIZZLE_128B is applied on
// absolute shared-memory address bits, verified on B200 with tools/exp_rowshift.cu) -- plus the three
// B tiles of that filter row: 3x fewer A bytes through L2 than one box per tap.
static constexpr int WIN_ROWS = 136;
static constexpr int WIN_BYTES = WIN_ROWS * 128;  // 17408 = 17 * 1024

// true when the 128-row tile starting at row0 contains at least one live row (segment mode)
__device__ __forceinline__ bool tile_live(const int* __restrict__ seg_counts, int seg_cap, int row0, int rows) {
  if (seg_counts == nullptr) return true;
  int r = row0;
  const int rend = min(row0 + BM, rows);
  while (r < rend) {
    const int n = r / seg_cap;
    if (r - n * seg_cap < min(seg_counts[n], seg_cap)) return true;
    r = (n + 1) * seg_cap;
  }
  return false;
}

template <bool ROWWIN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_aux,
               const GemmTnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] [2 x staging] [bias BN floats] [ctl]
  const int bn = p.bn;
  const int stages = p.stages;
  const bool bres = ROWWIN && p.b_resident;
  const int a_stage_bytes = ROWWIN ? WIN_BYTES : A_STAGE_BYTES;
  const int b_stage_bytes = bres ? 0 : (ROWWIN ? 3 : 1) * bn * BK * 2;
  const int stage_bytes = a_stage_bytes + b_stage_bytes;
  const int kStagingBufs = p.staging_bufs;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* b_res = smem + stages * stage_bytes;                       // 9 * bn * 128 B when resident
  uint8_t* staging = b_res + (bres ? 9 * bn * BK * 2 : 0);
  float* bias_s = reinterpret_cast<float*>(staging + kStagingBufs * STAGING_BYTES);
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(bias_s + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.rows + BM - 1) / BM;
  const int n_tiles = p.n_total / bn;
  const int tiles_per_batch = m_tiles * n_tiles;
  const int ksplit = p.ksplit;
  const int num_tiles = tiles_per_batch * p.batch * ksplit;   // work items (tile x k-split)
  const int k_chunks = p.k_per_tap / BK;
  const int k_iters_total = k_chunks * (ROWWIN ? 3 : p.taps);

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * bn)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.epi != EPI_F32_SPLIT && p.epi != EPI_ATOMIC_F32) tma_prefetch_desc(&map_d);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->tmem_full[i], 1);
      mbar_init(&ctl->tmem_empty[i], EPI_THREADS / 32);
      mbar_init(&ctl->aux_full[i], 1);
    }
    mbar_init(&ctl->bres_full, 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      if (bres) {
        mbar_arrive_expect_tx(&ctl->bres_full, 9 * bn * BK * 2);
        for (int ky = 0; ky < 3; ++ky) tma_load_3d(b_res + ky * 3 * bn * BK * 2, &map_b, &ctl->bres_full, 0, 0, ky * 3);
      }
      for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
        const int ks = work % ksplit;
        const int tile = work / ksplit;
        const int b = tile / tiles_per_batch;
        const int rem = tile - b * tiles_per_batch;
        const int mt = rem / n_tiles;
        const int nt = rem - mt * n_tiles;
        const int row0 = mt * BM;
        const int n0 = nt * bn;
        if (!tile_live(p.seg_counts, p.seg_cap, row0, p.rows)) continue;
        const int ki0 = (k_iters_total * ks) / ksplit, ki1 = (k_iters_total * (ks + 1)) / ksplit;
        for (int ki = ki0; ki < ki1; ++ki) {
          const int t = ki / k_chunks, kc = ki - t * k_chunks;
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + a_stage_bytes;
          mbar_arrive_expect_tx(&ctl->full[s], stage_bytes);
          if (ROWWIN) {
            // t = filter row ky: window starts one pixel left of the kx = 0 tap
            tma_load_3d(sa, &map_a, &ctl->full[s], kc * BK, row0 + (t - 1) * p.wp - 1, b);
            if (!bres) tma_load_3d(sb, &map_b, &ctl->full[s], kc * BK, n0, t * 3);
          } else {
            tma_load_3d(sa, &map_a, &ctl->full[s], kc * BK, row0 + p.shifts[t], b);
            tma_load_2d(sb, &map_b, &ctl->full[s], t * p.k_per_tap + kc * BK, n0);
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_f16(BM, bn, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    // warp-uniform copy of the TMEM base: a value loaded from shared memory is not provably uniform, and every
    // tcgen05.mma then paid an ELECT / R2UR.BROADCAST waterfall (~100 cycles per MMA, the bound of the N = 64 tiles)
    const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (bres) mbar_wait(&ctl->bres_full, 0);
    for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
      if (p.seg_counts != nullptr) {
        const int tile_ = work / ksplit;
        const int mt_ = (tile_ % tiles_per_batch) / n_tiles;
        if (!tile_live(p.seg_counts, p.seg_cap, mt_ * BM, p.rows)) continue;
      }
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      ++it;
      const int ks = work % ksplit;
      const int k_iters = (k_iters_total * (ks + 1)) / ksplit - (k_iters_total * ks) / ksplit;
      mbar_wait(&ctl->tmem_empty[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base_u + as * bn;
      for (int ki = 0; ki < k_iters; ++ki) {
        mbar_wait(&ctl->full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          // resident filter: stage index ki = filter row ky (single K chunk)
          const uint32_t b_addr = bres ? smem_u32(b_res) + ki * 3 * bn * BK * 2 : a_addr + a_stage_bytes;
          const uint64_t da = umma_desc_sw128(a_addr, 16, 1024);
          const uint64_t db = umma_desc_sw128(b_addr, 16, 1024);
          if (ROWWIN) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint64_t dax = da + (128 >> 4) * kx;               // one pixel row further
              const uint64_t dbx = db + ((bn * 128) >> 4) * kx;        // next tap's weight tile
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_f16_ss(d_tmem, dax + 2 * k, dbx + 2 * k, idesc, (ki > 0 || kx > 0 || k > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing 16 fp16 (32 B) along K inside the 128 B swizzle row = +2 in the >>4 field
              umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&ctl->empty[s]);
          if (ki == k_iters - 1) umma_commit(&ctl->tmem_full[as]);
        }
        __syncwarp();
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter, each converting one 32-column half of a 64-column chunk: with a single
    // warp per scheduler the ~450 dependent instructions per chunk ran at ~6 cycles each and the epilogue,
    // not the tensor pipe, bounded the small-K layers (conv1_2 .. conv3_1; ncu: tensor pipe 30 % active).
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int hf = (warp - 2) >> 2;    // column half handled by this warp
    const int r = q * 32 + lane;       // row of the 128-row tile owned by this thread
    const int et = threadIdx.x - 64;   // 0..255 index among epilogue threads
    const int nstg = p.staging_bufs;
    int it = 0;
    int st_buf = 0;
    int staged_n0 = -1;
    uint32_t aux_ph = 0;  // phase bit per staging buffer
    for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
      const int tile = work / ksplit;
      const int b = tile / tiles_per_batch;
      const int rem = tile - b * tiles_per_batch;
      const int mt = rem / n_tiles;
      const int nt = rem - mt * n_tiles;
      const int row0 = mt * BM;
      const int n0 = nt * bn;
      if (!tile_live(p.seg_counts, p.seg_cap, row0, p.rows)) continue;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      ++it;

      if (n0 != staged_n0) {
        // stage the bias slice (the previous slice's readers are past the first barrier)
        named_bar_sync(1, EPI_THREADS);
        for (int i = et; i < bn; i += EPI_THREADS)
          bias_s[i] = (p.bias != nullptr && n0 + i < p.n_bias) ? p.bias[n0 + i] : 0.f;
        named_bar_sync(1, EPI_THREADS);
        staged_n0 = n0;
      }

      mbar_wait(&ctl->tmem_full[as], aph);
      tc_fence_after();

      const int row = row0 + r;
      bool row_live = row < p.rows;
      if (p.wp > 0) row_live = row_live && ((row % p.wp) < p.w_valid);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * bn;

      if (p.epi == EPI_ATOMIC_F32) {
        // split == 1 ("slices"): every K split owns a slice [batch][rows][ld0] of d0 and stores its partial sum; the
        // finishing kernel adds the slices in a fixed order (a bit-reproducible forward). Otherwise fp32 red.add.
        const bool slices = p.split != 0;
        const int ks = work % ksplit;
        float* orow = p.d0 + ((static_cast<size_t>(slices ? ks : 0) * p.batch + b) * p.rows + row) * p.ld0 + n0;
        for (int c0 = 32 * hf; c0 < bn; c0 += 64) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c0, v);
          tmem_ld_wait();
          if (row < p.rows && slices) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(orow + c0 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else if (row < p.rows) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c0 + 4 * j),
                           "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                           "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                           : "memory");
          }
        }
      } else if (p.epi == EPI_F32_SPLIT) {
        for (int c0 = 16 * hf; c0 < bn; c0 += 32) {
          uint32_t v[16];
          tmem_ld_32x16(t_addr + c0, v);
          tmem_ld_wait();
          if (row < p.rows) {
            const size_t grow = static_cast<size_t>(b) * p.rows + row;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + c0 + j;
              if (n < p.n_valid) {
                const float x = __uint_as_float(v[j]) * p.alpha + bias_s[c0 + j];
                if (n < p.split)
                  p.d0[grow * p.ld0 + n] = x;
                else
                  p.d1[grow * p.ld1 + (n - p.split)] = x;
              }
            }
          }
        }
      } else if (p.epi == EPI_SPLIT3_RELU_F16 || p.epi == EPI_SPLIT3_F16) {
        // f16x3 output: hi chunk in staging buffer 0, lo chunk in buffer 1, three TMA stores
        // (columns [n], [n_total + n], [2 n_total + n] of the 3*n_total wide D). Needs both staging buffers.
        for (int c0 = 0; c0 < bn; c0 += 64) {
          if (warp == 2 && elect_one()) tma_store_wait_read<0>();
          named_bar_sync(1, EPI_THREADS);
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c0 + 32 * hf, v);
          tmem_ld_wait();
          uint8_t* rowh = staging + r * 128;
          uint8_t* rowl = staging + STAGING_BYTES + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = 4 * hf + j;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f[e] = __uint_as_float(v[j * 8 + e]) * p.alpha + bias_s[c0 + jj * 8 + e];
              if (p.epi == EPI_SPLIT3_RELU_F16) f[e] = fmaxf(f[e], 0.f);
              if (!row_live) f[e] = 0.f;
            }
            uint4 oh, ol;
            uint32_t* ohp = reinterpret_cast<uint32_t*>(&oh);
            uint32_t* olp = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 h = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
              const float2 hfl = __half22float2(h);
              const __half2 l = __floats2half2_rn(f[2 * e] - hfl.x, f[2 * e + 1] - hfl.y);
              ohp[e] = *reinterpret_cast<const uint32_t*>(&h);
              olp[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
            *reinterpret_cast<uint4*>(rowh + ((jj ^ (r & 7)) << 4)) = oh;
            *reinterpret_cast<uint4*>(rowl + ((jj ^ (r & 7)) << 4)) = ol;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EPI_THREADS);
          if (warp == 2 && elect_one()) {
            tma_store_3d(&map_d, staging, n0 + c0, row0, b);
            tma_store_3d(&map_d, staging + STAGING_BYTES, p.n_total + n0 + c0, row0, b);
            tma_store_3d(&map_d, staging, 2 * p.n_total + n0 + c0, row0, b);
            tma_store_commit();
          }
        }
      } else {
        const __half2 zero2 = __float2half2_rn(0.f);
        for (int c0 = 0; c0 < bn; c0 += 64) {
          uint8_t* stg = staging + st_buf * STAGING_BYTES;
          // the TMA store that last read this staging buffer must have drained
          if (warp == 2 && elect_one()) {
            if (nstg == 1) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
          }
          named_bar_sync(1, EPI_THREADS);
          if (p.epi == EPI_MASK_F16) {
            if (warp == 2 && elect_one()) {  // the aux tile is loaded into the staging buffer itself
              mbar_arrive_expect_tx(&ctl->aux_full[st_buf], STAGING_BYTES);
              tma_load_3d(stg, &map_aux, &ctl->aux_full[st_buf], n0 + c0, row0, b);
            }
          }
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c0 + 32 * hf, v);
          tmem_ld_wait();
          if (p.epi == EPI_MASK_F16) {
            mbar_wait(&ctl->aux_full[st_buf], (aux_ph >> st_buf) & 1u);
            aux_ph ^= 1u << st_buf;
          }
          uint8_t* rowp = stg + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = 4 * hf + j;
            uint4* dst = reinterpret_cast<uint4*>(rowp + ((jj ^ (r & 7)) << 4));
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0 + jj * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + jj * 8 + 4);
            float f[8];
            f[0] = __uint_as_float(v[j * 8 + 0]) * p.alpha + b0.x;
            f[1] = __uint_as_float(v[j * 8 + 1]) * p.alpha + b0.y;
            f[2] = __uint_as_float(v[j * 8 + 2]) * p.alpha + b0.z;
            f[3] = __uint_as_float(v[j * 8 + 3]) * p.alpha + b0.w;
            f[4] = __uint_as_float(v[j * 8 + 4]) * p.alpha + b1.x;
            f[5] = __uint_as_float(v[j * 8 + 5]) * p.alpha + b1.y;
            f[6] = __uint_as_float(v[j * 8 + 6]) * p.alpha + b1.z;
            f[7] = __uint_as_float(v[j * 8 + 7]) * p.alpha + b1.w;
            if (p.epi == EPI_MASK_F16) {
              // keep the gradient only where the forward activation (aux tile) was positive
              const uint4 a = *dst;
              const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 af = __half22float2(ah[e]);
                if (!(af.x > 0.f)) f[2 * e] = 0.f;
                if (!(af.y > 0.f)) f[2 * e + 1] = 0.f;
              }
            }
            __half2 h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
            if (p.epi == EPI_BIAS_RELU_F16) {
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __hmax2(h[e], zero2);  // relu(round(x)) == round(relu(x))
            }
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&h[0]);
            o.y = *reinterpret_cast<uint32_t*>(&h[1]);
            o.z = *reinterpret_cast<uint32_t*>(&h[2]);
            o.w = *reinterpret_cast<uint32_t*>(&h[3]);
            if (!row_live) o = make_uint4(0, 0, 0, 0);
            *dst = o;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EPI_THREADS);
          if (warp == 2 && elect_one()) {
            tma_store_3d(&map_d, stg, n0 + c0, row0, b);
            tma_store_commit();
          }
          if (nstg == 2) st_buf ^= 1;
        }
      }
      // this accumulator stage may now be overwritten by the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tmem_empty[as]);
    }
    if (warp == 2 && elect_one()) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// fp16 tensor map with up to 3 dims, innermost contiguous; box inner dim = 64 elements (128 B).
int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

static int g_num_sms = 0;

int gemm_tn_launch(const GemmTnArgs& a, cudaStream_t stream) {
  if (a.k_per_tap % BK != 0 || a.bn % 16 != 0 || a.bn > 256 || a.bn < 16) return 1001;
  if (a.n_total % a.bn != 0) return 1002;
  const bool f32_out = a.epi == EPI_F32_SPLIT || a.epi == EPI_ATOMIC_F32;
  if (!f32_out && a.bn % 64 != 0) return 1003;
  if (a.epi == EPI_ATOMIC_F32 && a.bn % 32 != 0) return 1006;
  if (a.ksplit > 1 && a.epi != EPI_ATOMIC_F32) return 1007;
  if (a.taps < 1 || a.taps > 9) return 1004;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // row-window mode: 3x3 conv tap pattern over the flattened rows and a narrow N tile (L2-bound case)
  static int rowwin_opt = -1;
  if (rowwin_opt < 0) {
    const char* e = getenv("PTB200_ROWWIN");
    rowwin_opt = (e == nullptr) ? 1 : atoi(e);
  }
  const bool split3 = a.epi == EPI_SPLIT3_RELU_F16 || a.epi == EPI_SPLIT3_F16;
  bool rowwin = rowwin_opt != 0 && a.taps == 9 && a.wp > 0 && a.bn <= 128 && !f32_out && a.ksplit <= 1 && !split3;
  if (rowwin)
    for (int t = 0; t < 9; ++t) rowwin = rowwin && a.shifts[t] == (t / 3 - 1) * a.wp + (t % 3 - 1);
  CUtensorMap ma, mb, md, mx;
  {
    uint64_t dims[3] = {(uint64_t)a.k_per_tap, (uint64_t)a.rows, (uint64_t)a.batch};
    uint64_t str[2] = {(uint64_t)a.lda * 2, (uint64_t)a.a_batch_stride * 2};
    uint32_t box[3] = {BK, (uint32_t)(rowwin ? WIN_ROWS : BM), 1};
    if (make_tmap_f16(&ma, a.A, 3, dims, str, box)) return 1010;
  }
  if (rowwin) {
    // B viewed as [tap][n][k] so that one box brings the three taps of a filter row
    uint64_t dims[3] = {(uint64_t)a.k_per_tap, (uint64_t)a.n_total, 9};
    uint64_t str[2] = {(uint64_t)a.k_per_tap * 9 * 2, (uint64_t)a.k_per_tap * 2};
    uint32_t box[3] = {BK, (uint32_t)a.bn, 3};
    if (make_tmap_f16(&mb, a.B, 3, dims, str, box)) return 1011;
  } else {
    uint64_t dims[2] = {(uint64_t)a.k_per_tap * a.taps, (uint64_t)a.n_total};
    uint64_t str[1] = {(uint64_t)a.k_per_tap * a.taps * 2};
    uint32_t box[2] = {BK, (uint32_t)a.bn};
    if (make_tmap_f16(&mb, a.B, 2, dims, str, box)) return 1011;
  }
  if (!f32_out) {
    uint64_t dims[3] = {(uint64_t)a.n_total * (split3 ? 3 : 1), (uint64_t)a.rows, (uint64_t)a.batch};
    uint64_t str[2] = {(uint64_t)a.ldd * 2, (uint64_t)a.d_batch_stride * 2};
    uint32_t box[3] = {64, BM, 1};
    if (make_tmap_f16(&md, a.D, 3, dims, str, box)) return 1012;
    if (a.epi == EPI_MASK_F16) {
      if (make_tmap_f16(&mx, a.aux, 3, dims, str, box)) return 1013;
    } else {
      mx = md;
    }
  } else {
    md = ma;
    mx = ma;
  }
  GemmTnParams p;
  p.batch = a.batch;
  p.rows = a.rows;
  p.k_per_tap = a.k_per_tap;
  p.taps = a.taps;
  for (int i = 0; i < 9; ++i) p.shifts[i] = i < a.taps ? a.shifts[i] : 0;
  p.n_total = a.n_total;
  p.bn = a.bn;
  p.w_valid = a.w_valid;
  p.wp = a.wp;
  p.epi = a.epi;
  p.ksplit = a.ksplit > 1 ? a.ksplit : 1;
  p.bias = a.bias;
  p.n_bias = a.n_bias;
  p.d0 = a.d0;
  p.ld0 = a.ld0;
  p.d1 = a.d1;
  p.ld1 = a.ld1;
  p.split = a.split;
  p.n_valid = a.n_valid;
  p.seg_counts = a.seg_counts;
  p.seg_cap = a.seg_cap;
  p.alpha = a.alpha;
  if (a.seg_counts != nullptr && (a.batch != 1 || a.seg_cap <= 0)) return 1008;
  // keep the whole filter resident when every CTA uses the same one (single N tile, single K chunk)
  const bool bres = rowwin && a.n_total == a.bn && a.k_per_tap == BK;
  p.b_resident = bres ? 1 : 0;
  const int stage_bytes = rowwin ? WIN_BYTES + (bres ? 0 : 3 * a.bn * BK * 2) : A_STAGE_BYTES + a.bn * BK * 2;
  // one staging buffer only where a second one would cost a pipeline stage that matters (row-window tiles
  // with streamed filters: 65 KB per stage); the f16x3 epilogue needs both (hi and lo chunk)
  p.staging_bufs = (rowwin && !bres) ? 1 : 2;
  const int fixed = p.staging_bufs * STAGING_BYTES + 256 * 4 + (int)sizeof(SmemCtl) + 1024 +
                    (bres ? 9 * a.bn * BK * 2 : 0);
  int stages = (232448 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return 1005;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + fixed;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         232448);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int m_tiles = (a.rows + BM - 1) / BM;
  const int num_tiles = m_tiles * (a.n_total / a.bn) * a.batch * p.ksplit;
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
  if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
  if (grid < 1) return 0;
  if (rowwin)
    gemm_tn_kernel<true><<<grid, NUM_THREADS, smem_bytes, stream>>>(ma, mb, md, mx, p);
  else
    gemm_tn_kernel<false><<<grid, NUM_THREADS, smem_bytes, stream>>>(ma, mb, md, mx, p);
  return (int)cudaGetLastError();
}

}  // namespace ptb
