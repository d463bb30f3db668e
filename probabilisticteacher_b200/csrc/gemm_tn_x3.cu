// f16x3 ("split-fp16", fp32-equivalent) implicit GEMM for sm_100a with IN-KERNEL fp32 PROMOTION.
//
// Same contraction as gemm_tn.cu (D[b][p][n] = sum_t sum_k A[b][p + shift_t][k] * B[n][t*K + k], K-major fp16
// operands staged by TMA, tcgen05.mma kind::f16 into TMEM), used where the reference computes in fp32
// (pt/engine/trainer.py:271-277 runs without autocast): activations are [hi | lo | hi] triples, weights
// [Wh | Wh | Wl] triples, so the main loop evaluates hi*Wh + lo*Wh + hi*Wl.
//
// What is different from gemm_tn.cu: the tensor core adds into its fp32 accumulator with TRUNCATION (measured on
// B200: one chain over K = 3 x 25088 is biased by -2.4e-4 relative, ~0.5 ulp per MMA), which compounds through 16
// stacked layers. Round 1 cut the reduction into split-K chunks that were summed with red.global.add (round to
// nearest) -- 50-100 fp32 atomic passes over every output tensor. Here the MMA warp still works in chunks of
// `chunk` k-iterations (4 MMAs each) alternating between the two TMEM accumulator stages, but the epilogue warps
// PROMOTE each finished chunk into fp32 REGISTER accumulators (tcgen05.ld + FADD, round to nearest) while the
// tensor core runs the next chunk; the output tile is written once. The tensor pipe stays the bound: a chunk of 4
// k-iterations at N = 256 is 2048 tensor cycles against ~128 FADD + 4 tcgen05.ld per epilogue thread.
//
// Epilogues (all from the register accumulators):
//   EPI_SPLIT3_RELU_F16 / EPI_SPLIT3_F16 : act(alpha*acc + bias) -> [hi | lo | hi] triple (forward)
//   EPI_SPLIT3_MASK_F16                  : (aux > 0 ? alpha*acc : 0) -> triple; aux = forward activation triple
//                                          (ReLU backward fused into the data-gradient GEMM)
//   EPI_F32_STORE                        : alpha*acc (+ bias) -> fp32 d0[row][n] (data gradients consumed by the
//                                          max-pool / ROIAlign backward kernels)
//   EPI_ATOMIC_F32                       : split-K partial (skinny problems): d0[row][n] += acc
#include "ptx.cuh"
#include "gemm_tn.h"
#include <stdio.h>
#include <stdlib.h>

namespace ptb {

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int STAGING_BYTES = BM * 128;
constexpr int NUM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int EPI_THREADS = 256;

// ROWWIN (3x3 convs with N <= 128, where streaming one A box per tap makes the tile L2-BANDWIDTH bound: conv1_2 in
// f16x3 pulled 432 KB of A per 128-pixel tile through L2, 10 TB/s against a ~12 TB/s LTS cap, tensor pipe 31 %): the
// A ring holds 136-row windows (rows p0 + (ky-1)*Wp - 1 ...) shared by the three horizontal taps of a filter row --
// the taps are UMMA descriptors offset by kx*128 B, SWIZZLE_128B being a function of absolute shared-memory address
// bits (tools/exp_rowshift.cu) -- and the B tiles travel through a ring of their own (one tile per tap and K chunk),
// so a slot is recycled as soon as its 4 MMAs retire: 2.8x fewer A bytes through L2. With hi_share the k-iterations of
// a filter row are (hi_j, lo_j) pairs: the `hi` window of channel chunk j is loaded ONCE and multiplied with the Wh
// tiles (K chunk j) and the Wl tiles (K chunk 2C/64 + j) -- the [hi | lo | hi] triple holds it twice --, the `lo`
// window with the second Wh copy: a third fewer A bytes again. hi_share == 2 ("pair mode") goes one step further: a
// k-iteration is the PAIR (hi_j, lo_j) of windows, both resident, and every Wh tile is loaded once for both of them
// (the [Wh | Wh | Wl] triple holds it twice): 6 instead of 9 B tiles per pair, a third fewer B bytes -- B is the
// larger part of the L2 traffic at N = 256.
constexpr int WIN_ROWS = 136;
constexpr int WIN_BYTES = WIN_ROWS * 128;  // 17 * 1024
constexpr int RW_MAX_A = 4;
constexpr int RW_MAX_B = 16;

struct Ctl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t full_a[RW_MAX_A];
  uint64_t empty_a[RW_MAX_A];
  uint64_t full_b[RW_MAX_B];
  uint64_t empty_b[RW_MAX_B];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t aux_full;
  uint32_t tmem_base;
  uint32_t pad;
};

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

__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// MH = 2 (row-window mode, N <= 128): a CTA tile is TWO 128-row halves that share every B tile (the filter slice of
// a tap is read from L2 once per 256 pixels; B was 60-75 % of the L2 bytes of these layers), each half with its own
// pair of TMEM accumulator stages and its own fp32 register accumulators.
template <int NCH, bool ROWWIN, int MH>  // N tile = 64 * NCH columns, M tile = 128 * MH rows
// 10 warps = 3 on two of the four SM sub-partitions (16 K registers each): at most 168 registers per thread
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_promote_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_aux,
                       const GemmTnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int bn = 64 * NCH;
  constexpr int stage_bytes = A_STAGE_BYTES + bn * BK * 2;
  constexpr int b_tile_bytes = bn * BK * 2;
  static_assert(MH == 1 || (ROWWIN && NCH <= 2), "two M halves: row-window mode, N <= 128");
  constexpr int a_slot_bytes = MH * WIN_BYTES;  // ROWWIN
  const int stages = p.stages;      // ROWWIN: slots of the A-window ring
  const int b_slots = p.b_resident;  // ROWWIN: slots of the B-tile ring (the field is otherwise unused here)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ring_b = smem + stages * a_slot_bytes;  // ROWWIN only
  // two 16 KB buffers: hi chunk, lo chunk
  uint8_t* staging = ROWWIN ? ring_b + b_slots * b_tile_bytes : smem + stages * stage_bytes;
  float* bias_s = reinterpret_cast<float*>(staging + 2 * STAGING_BYTES);
  Ctl* ctl = reinterpret_cast<Ctl*>(bias_s + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  constexpr int BMT = BM * MH;
  const int m_tiles = (p.rows + BMT - 1) / BMT;
  const int n_tiles = p.n_total / bn;
  const int tiles_per_batch = m_tiles * n_tiles;
  const int ksplit = p.ksplit;
  const int num_tiles = tiles_per_batch * p.batch * ksplit;
  const int k_chunks = p.k_per_tap / BK;
  // ROWWIN: one k-iteration = a filter row's 3 taps over one A window; hi_share: 2 windows (hi, lo) per channel chunk
  const bool hi_share = ROWWIN && p.hi_share == 1;
  const bool pair_mode = ROWWIN && p.hi_share == 2;
  const int kc3 = k_chunks / 3;
  // pair mode also runs plain GEMMs (rw_ny = rw_nx = 1: the fc layers) through the A / B rings: the same sharing of the
  // hi window and of the Wh tile applies to any [hi | lo | hi] x [Wh | Wh | Wl] contraction
  const int rw_ny = pair_mode ? p.rw_ny : 3, rw_nx = pair_mode ? p.rw_nx : 3;
  const int rw_row_shift = rw_nx == 3 ? -1 : 0;       // the window starts one pixel left of the kx = 0 tap
  const int rw_dy = rw_ny == 3 ? p.wp : 0, rw_y0 = rw_ny == 3 ? 1 : 0;
  const int k_iters_total = ROWWIN ? rw_ny * (pair_mode ? kc3 : hi_share ? 2 * kc3 : k_chunks) : k_chunks * p.taps;
  const int chunk = p.chunk;

  constexpr uint32_t tmem_cols = 2 * MH * bn < 32 ? 32 : 2 * MH * bn;  // 128 / 256 / 512: powers of two

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.epi != EPI_F32_STORE && p.epi != EPI_ATOMIC_F32) tma_prefetch_desc(&map_d);
    if (ROWWIN) {
      for (int i = 0; i < stages; ++i) {
        mbar_init(&ctl->full_a[i], 1);
        mbar_init(&ctl->empty_a[i], 1);
      }
      for (int i = 0; i < b_slots; ++i) {
        mbar_init(&ctl->full_b[i], 1);
        mbar_init(&ctl->empty_b[i], 1);
      }
    } else {
      for (int i = 0; i < stages; ++i) {
        mbar_init(&ctl->full[i], 1);
        mbar_init(&ctl->empty[i], 1);
      }
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->tmem_full[i], 1);
      mbar_init(&ctl->tmem_empty[i], EPI_THREADS / 32);
    }
    mbar_init(&ctl->aux_full, 1);
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
      int sb = 0;
      uint32_t phb = 0;
      for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
        const int ks = work % ksplit;
        const int tile = work / ksplit;
        const int b = tile / tiles_per_batch;
        const int rem = tile - b * tiles_per_batch;
        const int mt = rem / n_tiles;
        const int nt = rem - mt * n_tiles;
        const int row0 = mt * BMT;
        const int n0 = nt * bn;
        if (!tile_live(p.seg_counts, p.seg_cap, row0, p.rows)) continue;
        const int ki0 = (k_iters_total * ks) / ksplit, ki1 = (k_iters_total * (ks + 1)) / ksplit;
        for (int ki = ki0; ki < ki1; ++ki) {
          int t = ki / k_chunks, kc = ki - t * k_chunks;
          if (ROWWIN && pair_mode) {
            t = ki / kc3;
            const int j = ki - t * kc3;
            for (int part = 0; part < 2; ++part) {  // the hi window(s), then the lo window(s): two consecutive A slots
              mbar_wait(&ctl->empty_a[s], ph ^ 1);
              mbar_arrive_expect_tx(&ctl->full_a[s], a_slot_bytes);
#pragma unroll
              for (int h = 0; h < MH; ++h)
                tma_load_3d(smem + s * a_slot_bytes + h * WIN_BYTES, &map_a, &ctl->full_a[s], (part ? kc3 + j : j) * BK,
                            row0 + h * BM + (t - rw_y0) * rw_dy + rw_row_shift, b);
              if (++s == stages) {
                s = 0;
                ph ^= 1;
              }
            }
            for (int bt = 0; bt < 2 * rw_nx; ++bt) {  // per tap: the Wh tile (for hi and lo), then the Wl tile (for hi)
              mbar_wait(&ctl->empty_b[sb], phb ^ 1);
              mbar_arrive_expect_tx(&ctl->full_b[sb], b_tile_bytes);
              tma_load_2d(ring_b + sb * b_tile_bytes, &map_b, &ctl->full_b[sb],
                          (t * rw_nx + (bt >> 1)) * p.k_per_tap + ((bt & 1) ? 2 * kc3 + j : j) * BK, n0);
              if (++sb == b_slots) {
                sb = 0;
                phb ^= 1;
              }
            }
            continue;
          }
          if (ROWWIN) {
            // t = filter row ky: the window starts one pixel left of the kx = 0 tap
            int kc_b2 = -1;  // hi_share, hi window: second B chunk (the Wl tiles)
            if (hi_share) {
              t = ki / (2 * kc3);
              const int rr = ki - t * 2 * kc3, j = rr >> 1;
              kc = (rr & 1) ? kc3 + j : j;
              if (!(rr & 1)) kc_b2 = 2 * kc3 + j;
            }
            mbar_wait(&ctl->empty_a[s], ph ^ 1);
            mbar_arrive_expect_tx(&ctl->full_a[s], a_slot_bytes);
#pragma unroll
            for (int h = 0; h < MH; ++h)
              tma_load_3d(smem + s * a_slot_bytes + h * WIN_BYTES, &map_a, &ctl->full_a[s], kc * BK,
                          row0 + h * BM + (t - 1) * p.wp - 1, b);
            if (++s == stages) {
              s = 0;
              ph ^= 1;
            }
            for (int kx = 0; kx < 3; ++kx) {
              for (int g = 0; g < (kc_b2 >= 0 ? 2 : 1); ++g) {
                mbar_wait(&ctl->empty_b[sb], phb ^ 1);
                mbar_arrive_expect_tx(&ctl->full_b[sb], b_tile_bytes);
                tma_load_2d(ring_b + sb * b_tile_bytes, &map_b, &ctl->full_b[sb],
                            (t * 3 + kx) * p.k_per_tap + (g == 0 ? kc : kc_b2) * BK, n0);
                if (++sb == b_slots) {
                  sb = 0;
                  phb ^= 1;
                }
              }
            }
            continue;
          }
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&ctl->full[s], stage_bytes);
          tma_load_3d(sa, &map_a, &ctl->full[s], kc * BK, row0 + p.shifts[t], b);
          tma_load_2d(sb, &map_b, &ctl->full[s], t * p.k_per_tap + kc * BK, n0);
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (chunked)
    const uint32_t idesc = umma_idesc_f16(BM, bn, 0, 0);
    int s = 0;
    uint32_t ph = 0;
    int sb = 0;
    uint32_t phb = 0;
    int it = 0;  // accumulator-stage use counter: one per CHUNK
    const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
      if (p.seg_counts != nullptr) {
        const int tile_ = work / ksplit;
        const int mt_ = (tile_ % tiles_per_batch) / n_tiles;
        if (!tile_live(p.seg_counts, p.seg_cap, mt_ * BMT, p.rows)) continue;
      }
      const int ks = work % ksplit;
      const int k_iters = (k_iters_total * (ks + 1)) / ksplit - (k_iters_total * ks) / ksplit;
      for (int kb = 0; kb < k_iters; kb += chunk) {
        const int kn = min(chunk, k_iters - kb);
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        ++it;
        mbar_wait(&ctl->tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base_u + as * (MH * bn);
        for (int ki = 0; ki < kn; ++ki) {
          if (ROWWIN && pair_mode) {
            const int s_hi = s;
            const uint32_t ph_hi = ph;
            if (++s == stages) {
              s = 0;
              ph ^= 1;
            }
            const int s_lo = s;
            const uint32_t ph_lo = ph;
            if (++s == stages) {
              s = 0;
              ph ^= 1;
            }
            mbar_wait(&ctl->full_a[s_hi], ph_hi);
            mbar_wait(&ctl->full_a[s_lo], ph_lo);
            const uint64_t da_hi = umma_desc_sw128(smem_u32(smem + s_hi * a_slot_bytes), 16, 1024);
            const uint64_t da_lo = umma_desc_sw128(smem_u32(smem + s_lo * a_slot_bytes), 16, 1024);
            const int nbt = 2 * rw_nx;
            for (int bt = 0; bt < nbt; ++bt) {
              const int kx = bt >> 1;
              mbar_wait(&ctl->full_b[sb], phb);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t db = umma_desc_sw128(smem_u32(ring_b + sb * b_tile_bytes), 16, 1024);
#pragma unroll
                for (int h = 0; h < MH; ++h) {
                  const uint64_t off = (h * WIN_BYTES >> 4) + (128 >> 4) * kx;  // half h, one pixel row further
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_f16_ss(d_tmem + h * bn, da_hi + off + 2 * k, db + 2 * k, idesc, (ki > 0 || bt > 0 || k > 0) ? 1u : 0u);
                  if (!(bt & 1)) {  // a Wh tile also meets the lo window
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) umma_f16_ss(d_tmem + h * bn, da_lo + off + 2 * k, db + 2 * k, idesc, 1u);
                  }
                }
                umma_commit(&ctl->empty_b[sb]);
                if (bt == nbt - 1) {
                  umma_commit(&ctl->empty_a[s_hi]);
                  umma_commit(&ctl->empty_a[s_lo]);
                  if (ki == kn - 1) umma_commit(&ctl->tmem_full[as]);
                }
              }
              __syncwarp();
              if (++sb == b_slots) {
                sb = 0;
                phb ^= 1;
              }
            }
            continue;
          }
          if (ROWWIN) {
            mbar_wait(&ctl->full_a[s], ph);
            const uint64_t da = umma_desc_sw128(smem_u32(smem + s * a_slot_bytes), 16, 1024);
            // hi_share: even k-iterations of a tile hold a `hi` window that meets two B tiles per tap (Wh, Wl)
            const int nb = (hi_share && !((kb + ki) & 1)) ? 6 : 3;
            for (int bt = 0; bt < nb; ++bt) {
              const int kx = nb == 6 ? bt >> 1 : bt;
              mbar_wait(&ctl->full_b[sb], phb);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t db = umma_desc_sw128(smem_u32(ring_b + sb * b_tile_bytes), 16, 1024);
#pragma unroll
                for (int h = 0; h < MH; ++h) {
                  const uint64_t dax = da + (h * WIN_BYTES >> 4) + (128 >> 4) * kx;  // half h, one pixel row further
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k)
                    umma_f16_ss(d_tmem + h * bn, dax + 2 * k, db + 2 * k, idesc, (ki > 0 || bt > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&ctl->empty_b[sb]);
                if (bt == nb - 1) {
                  umma_commit(&ctl->empty_a[s]);
                  if (ki == kn - 1) umma_commit(&ctl->tmem_full[as]);
                }
              }
              __syncwarp();
              if (++sb == b_slots) {
                sb = 0;
                phb ^= 1;
              }
            }
            if (++s == stages) {
              s = 0;
              ph ^= 1;
            }
            continue;
          }
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
            const uint64_t da = umma_desc_sw128(a_addr, 16, 1024);
            const uint64_t db = umma_desc_sw128(a_addr + A_STAGE_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
            umma_commit(&ctl->empty[s]);
            if (ki == kn - 1) umma_commit(&ctl->tmem_full[as]);
          }
          __syncwarp();
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ promotion + epilogue (warps 2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int hf = (warp - 2) >> 2;    // which 32-column half of every 64-column group this warp owns
    const int r = q * 32 + lane;       // row of the 128-row tile owned by this thread
    const int et = threadIdx.x - 64;
    int it = 0;
    int staged_n0 = -1;
    uint32_t aux_ph = 0;
    for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
      const int tile = work / ksplit;
      const int ks = work % ksplit;
      const int b = tile / tiles_per_batch;
      const int rem = tile - b * tiles_per_batch;
      const int mt = rem / n_tiles;
      const int nt = rem - mt * n_tiles;
      const int row0 = mt * BMT;
      const int n0 = nt * bn;
      if (!tile_live(p.seg_counts, p.seg_cap, row0, p.rows)) continue;
      const int k_iters = (k_iters_total * (ks + 1)) / ksplit - (k_iters_total * ks) / ksplit;

      if (n0 != staged_n0) {
        named_bar_sync(1, EPI_THREADS);
        for (int i = et; i < bn; i += EPI_THREADS)
          bias_s[i] = (p.bias != nullptr && n0 + i < p.n_bias) ? p.bias[n0 + i] : 0.f;
        named_bar_sync(1, EPI_THREADS);
        staged_n0 = n0;
      }

      // accumulator group g = (M half h, 64-column chunk ch), g = h * NCH + ch: TMEM columns g * 64 of a stage
      constexpr int G = MH * NCH;
      float acc[G][32];
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[g][j] = 0.f;

      for (int kb = 0; kb < k_iters; kb += chunk) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        ++it;
        mbar_wait(&ctl->tmem_full[as], aph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * (MH * bn) + 32 * hf;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          // 16 columns at a time: with 128 accumulators live, a 32-register landing buffer would spill
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[16];
            tmem_ld_32x16(t_addr + g * 64 + hh * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[g][hh * 16 + j] += __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->tmem_empty[as]);
      }

      if (p.epi == EPI_ATOMIC_F32) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int row = row0 + (g / NCH) * BM + r;
          if (row < p.rows) {
            float* orow = p.d0 + (static_cast<size_t>(b) * p.rows + row) * p.ld0 + n0 + (g % NCH) * 64 + 32 * hf;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + 4 * j), "f"(acc[g][4 * j]),
                           "f"(acc[g][4 * j + 1]), "f"(acc[g][4 * j + 2]), "f"(acc[g][4 * j + 3])
                           : "memory");
          }
        }
      } else if (p.epi == EPI_F32_STORE) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int row = row0 + (g / NCH) * BM + r;
          bool row_live = row < p.rows;
          if (p.wp > 0) row_live = row_live && ((row % p.wp) < p.w_valid);
          if (row < p.rows) {
            // ksplit > 1: one slice [batch][rows][ld0] per K split (summed in a fixed order by the finishing kernel)
            float* orow = p.d0 + ((static_cast<size_t>(ks) * p.batch + b) * p.rows + row) * p.ld0 + n0 +
                          (g % NCH) * 64 + 32 * hf;
            const float* bs = bias_s + (g % NCH) * 64 + 32 * hf;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o;
              o.x = row_live ? acc[g][4 * j] * p.alpha + bs[4 * j] : 0.f;
              o.y = row_live ? acc[g][4 * j + 1] * p.alpha + bs[4 * j + 1] : 0.f;
              o.z = row_live ? acc[g][4 * j + 2] * p.alpha + bs[4 * j + 2] : 0.f;
              o.w = row_live ? acc[g][4 * j + 3] * p.alpha + bs[4 * j + 3] : 0.f;
              *reinterpret_cast<float4*>(orow + 4 * j) = o;
            }
          }
        }
      } else {
        // triple output: hi chunk in staging buffer 0, lo chunk in buffer 1, three TMA stores per 64 columns
        uint8_t* rowh = staging + r * 128;
        uint8_t* rowl = staging + STAGING_BYTES + r * 128;
        const bool masked = p.epi == EPI_SPLIT3_MASK_F16;
        const bool relu = p.epi == EPI_SPLIT3_RELU_F16;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int c0 = (g % NCH) * 64;
          const int rowg0 = row0 + (g / NCH) * BM;
          const int row = rowg0 + r;
          bool row_live = row < p.rows;
          if (p.wp > 0) row_live = row_live && ((row % p.wp) < p.w_valid);
          if (warp == 2 && elect_one()) tma_store_wait_read<0>();
          named_bar_sync(1, EPI_THREADS);
          if (masked) {
            if (warp == 2 && elect_one()) {  // forward activation (hi, lo) tiles land in the staging buffers
              mbar_arrive_expect_tx(&ctl->aux_full, 2 * STAGING_BYTES);
              tma_load_3d(staging, &map_aux, &ctl->aux_full, n0 + c0, rowg0, b);
              tma_load_3d(staging + STAGING_BYTES, &map_aux, &ctl->aux_full, p.n_total + n0 + c0, rowg0, b);
            }
            mbar_wait(&ctl->aux_full, aux_ph);
            aux_ph ^= 1u;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = 4 * hf + j;
            uint4* dh = reinterpret_cast<uint4*>(rowh + ((jj ^ (r & 7)) << 4));
            uint4* dl = reinterpret_cast<uint4*>(rowl + ((jj ^ (r & 7)) << 4));
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f[e] = acc[g][j * 8 + e] * p.alpha + bias_s[c0 + jj * 8 + e];
              if (relu) f[e] = fmaxf(f[e], 0.f);
              if (!row_live) f[e] = 0.f;
            }
            if (masked) {
              const uint4 ah = *dh, al = *dl;
              const __half2* ahp = reinterpret_cast<const __half2*>(&ah);
              const __half2* alp = reinterpret_cast<const __half2*>(&al);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xh = __half22float2(ahp[e]), xl = __half22float2(alp[e]);
                if (!(xh.x + xl.x > 0.f)) f[2 * e] = 0.f;
                if (!(xh.y + xl.y > 0.f)) f[2 * e + 1] = 0.f;
              }
            }
            uint4 oh, ol;
            split_pair(f[0], f[1], oh.x, ol.x);
            split_pair(f[2], f[3], oh.y, ol.y);
            split_pair(f[4], f[5], oh.z, ol.z);
            split_pair(f[6], f[7], oh.w, ol.w);
            *dh = oh;
            *dl = ol;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, EPI_THREADS);
          // (a half that starts beyond the last row has nothing to store: the box would lie outside the tensor)
          if (warp == 2 && rowg0 < p.rows && elect_one()) {
            tma_store_3d(&map_d, staging, n0 + c0, rowg0, b);
            tma_store_3d(&map_d, staging + STAGING_BYTES, p.n_total + n0 + c0, rowg0, b);
            tma_store_3d(&map_d, staging, 2 * p.n_total + n0 + c0, rowg0, b);
            tma_store_commit();
          }
        }
      }
    }
    if (warp == 2 && elect_one()) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

int g_num_sms = 0;

template <int NCH, bool ROWWIN, int MH>
int launch_nch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const CUtensorMap& mx,
               GemmTnParams& p, int max_ctas, cudaStream_t stream) {
  constexpr int bn = 64 * NCH;
  constexpr int stage_bytes = A_STAGE_BYTES + bn * BK * 2;
  const int fixed = 2 * STAGING_BYTES + 256 * 4 + (int)sizeof(Ctl) + 1024;
  int smem_bytes;
  if (ROWWIN) {
    // A-window ring + B-tile ring: the B ring takes what 3 (N = 128) / 4 (N = 64) windows leave
    // pair mode needs the hi and the lo slot of a pair resident together: four slots = two pairs
    const int a_slots = ((NCH == 1 && MH == 1) || p.hi_share == 2) ? 4 : 3;
    int b_slots = (232448 - fixed - a_slots * MH * WIN_BYTES) / (bn * BK * 2);
    if (b_slots > RW_MAX_B) b_slots = RW_MAX_B;
    if (b_slots < 3) return 1005;
    p.stages = a_slots;
    p.b_resident = b_slots;
    smem_bytes = a_slots * MH * WIN_BYTES + b_slots * bn * BK * 2 + fixed;
  } else {
    int stages = (232448 - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return 1005;
    p.stages = stages;
    smem_bytes = stages * stage_bytes + fixed;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_promote_kernel<NCH, ROWWIN, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         232448);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int m_tiles = (p.rows + BM * MH - 1) / (BM * MH);
  const int num_tiles = m_tiles * (p.n_total / bn) * p.batch * p.ksplit;
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) return 0;
  gemm_tn_promote_kernel<NCH, ROWWIN, MH><<<grid, NUM_THREADS, smem_bytes, stream>>>(ma, mb, md, mx, p);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, gemm_tn_promote_kernel<NCH, ROWWIN, MH>) == cudaSuccess)
      fprintf(stderr, "gemm_tn_promote_kernel<%d>: launch failed (%s): regs %d, maxThreadsPerBlock %d, static smem %zu, "
              "max dynamic smem %d, requested %d threads / %d B\n", NCH, cudaGetErrorString(err), fa.numRegs,
              fa.maxThreadsPerBlock, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, NUM_THREADS, smem_bytes);
  }
  return (int)err;
}

}  // namespace

// a.D / a.aux: triples [batch][rows][3*n_total] (ldd = their row pitch); chunk = k-iterations of 64 per promotion
int gemm_tn_promote_launch(const GemmTnArgs& a, int chunk, cudaStream_t stream) {
  if (a.k_per_tap % BK != 0) return 1001;
  if (a.bn != 64 && a.bn != 128 && a.bn != 256) return 1003;
  if (a.n_total % a.bn != 0) return 1002;
  if (a.taps < 1 || a.taps > 9) return 1004;
  const bool f32_out = a.epi == EPI_F32_STORE || a.epi == EPI_ATOMIC_F32;
  const bool triple = a.epi == EPI_SPLIT3_RELU_F16 || a.epi == EPI_SPLIT3_F16 || a.epi == EPI_SPLIT3_MASK_F16;
  if (!f32_out && !triple) return 1021;
  if (a.ksplit > 1 && a.epi != EPI_ATOMIC_F32 && !(a.epi == EPI_F32_STORE && a.bias == nullptr)) return 1007;
  if (a.seg_counts != nullptr && (a.batch != 1 || a.seg_cap <= 0)) return 1008;
  if (a.epi == EPI_SPLIT3_MASK_F16 && a.aux == nullptr) return 1009;
  if (f32_out && a.d0 == nullptr) return 1009;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // row-window mode: 3x3 tap pattern over the flattened rows, narrow N tile (the L2-bound case), whole-K tiles
  static int rowwin_opt = -1, rw_chunk = 0;
  if (rowwin_opt < 0) {
    const char* e = getenv("PTB200_X3_ROWWIN");
    rowwin_opt = (e == nullptr) ? 1 : atoi(e);
    const char* c = getenv("PTB200_X3_RW_CHUNK");
    rw_chunk = c ? atoi(c) : 0;
  }
  // (PTB200_X3_ROWWIN = 2 restricts it to the N <= 128 layers)
  bool rowwin = rowwin_opt != 0 && a.taps == 9 && a.wp > 0 && (a.bn <= 128 || rowwin_opt == 1) && a.ksplit <= 1 &&
                a.seg_counts == nullptr;
  if (rowwin)
    for (int t = 0; t < 9; ++t) rowwin = rowwin && a.shifts[t] == (t / 3 - 1) * a.wp + (t % 3 - 1);
  // plain GEMMs at N = 256 tiles (fc1 / fc2 and their data gradients; split-K and segment mode included) run through
  // the same A / B rings in pair mode
  static int hs_env = -2, plain_opt = 1;
  if (hs_env == -2) {
    const char* e = getenv("PTB200_X3_HISHARE");
    hs_env = e ? atoi(e) : -1;
    const char* q = getenv("PTB200_X3_PLAIN");  // 0: plain GEMMs keep the per-tap pipeline (A/B)
    plain_opt = q ? atoi(q) : 1;
  }
  const bool plain_pair = plain_opt != 0 && rowwin_opt == 1 && (hs_env < 0 || hs_env == 2) && a.taps == 1 && a.shifts[0] == 0 && a.bn == 256;
  if (plain_pair) rowwin = true;
  CUtensorMap ma, mb, md, mx;
  {
    uint64_t dims[3] = {(uint64_t)a.k_per_tap, (uint64_t)a.rows, (uint64_t)a.batch};
    uint64_t str[2] = {(uint64_t)a.lda * 2, (uint64_t)a.a_batch_stride * 2};
    uint32_t box[3] = {BK, (uint32_t)(rowwin ? WIN_ROWS : BM), 1};
    if (make_tmap_f16(&ma, a.A, 3, dims, str, box)) return 1010;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.k_per_tap * a.taps, (uint64_t)a.n_total};
    uint64_t str[1] = {(uint64_t)a.k_per_tap * a.taps * 2};
    uint32_t box[2] = {BK, (uint32_t)a.bn};
    if (make_tmap_f16(&mb, a.B, 2, dims, str, box)) return 1011;
  }
  if (triple) {
    uint64_t dims[3] = {(uint64_t)a.n_total * 3, (uint64_t)a.rows, (uint64_t)a.batch};
    uint64_t str[2] = {(uint64_t)a.ldd * 2, (uint64_t)a.d_batch_stride * 2};
    uint32_t box[3] = {64, BM, 1};
    if (make_tmap_f16(&md, a.D, 3, dims, str, box)) return 1012;
    if (a.epi == EPI_SPLIT3_MASK_F16) {
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
  p.b_resident = 0;
  p.staging_bufs = 2;
  p.bias = a.bias;
  p.n_bias = a.n_bias;
  p.d0 = a.d0;
  p.ld0 = a.ld0;
  p.d1 = nullptr;
  p.ld1 = 0;
  p.split = 0;
  p.n_valid = a.n_total;
  p.seg_counts = a.seg_counts;
  p.seg_cap = a.seg_cap;
  p.alpha = a.alpha;
  static int chunk_env = -1;
  if (chunk_env < 0) {
    const char* e = getenv("PTB200_X3_CHUNK");
    chunk_env = e ? atoi(e) : 0;
  }
  p.chunk = chunk_env > 0 ? chunk_env : (chunk > 0 ? chunk : 4);
  if (rowwin) {
    // a row-window k-iteration issues 12 MMAs (3 taps x 64 columns of K). One promotion per 3 k-iterations: chains of
    // 36 truncating tensor-core accumulations (16 in the per-tap kernel at chunk 4) over K <= 3456 -- measured on
    // B200 at full size, conv1_2: chunk 1 / 3 / 9 = 0.681 / 0.609 / 0.592 ms per 2 images (per-tap kernel 0.725) at
    // 2.5e-7 / 8.6e-7 / 2.1e-6 relative error against fp64
    // N = 256: a k-iteration is 1536 tensor cycles, one promotion per k-iteration costs nothing (chains of 12 MMAs)
    static int hs_opt = -2;
    if (hs_opt == -2) {
      const char* e = getenv("PTB200_X3_HISHARE");
      hs_opt = e ? atoi(e) : -1;
    }
    // default: pair mode (measured against hi sharing, ms per 2 images: 512->512 conv 0.350 -> 0.329, conv1_2 0.528 ->
    // 0.504 with four A slots -- with three, the B ring drained at every pair boundary and conv1_2 ran at 0.571).
    // PTB200_X3_HISHARE = 0 / 1 / 2: off / hi sharing / pair mode
    p.hi_share = hs_opt < 0 ? 2 : hs_opt;
    p.rw_ny = p.rw_nx = plain_pair ? 1 : 3;
    // hi_share: a (hi, lo) pair of k-iterations is 24 + 12 MMAs: promote per pair at N <= 128, per window at N = 256;
    // pair mode (2): one k-iteration IS the pair (36 MMAs per accumulator), promoted every time
    if (p.hi_share == 2)
      p.chunk = rw_chunk > 0 ? rw_chunk : 1;
    else if (p.hi_share)
      p.chunk = rw_chunk > 0 ? rw_chunk : (a.bn == 256 ? 1 : 2);
    else
      p.chunk = rw_chunk > 0 ? rw_chunk : (a.bn == 256 ? 1 : 3);
    // a caller that asks for very long chains (chunk >= 64: the test that demonstrates the truncation bias of a single
    // tensor-core accumulation chain) gets them in this mode's own k-iteration unit
    if (chunk >= 64) p.chunk = chunk;
    static int mh_opt = -1;
    if (mh_opt < 0) {
      const char* e = getenv("PTB200_X3_MH");
      mh_opt = e ? atoi(e) : 2;
    }
    if (a.bn == 256) return launch_nch<4, true, 1>(ma, mb, md, mx, p, a.max_ctas, stream);
    if (mh_opt == 2 && a.rows > 2 * BM)
      return a.bn == 64 ? launch_nch<1, true, 2>(ma, mb, md, mx, p, a.max_ctas, stream)
                        : launch_nch<2, true, 2>(ma, mb, md, mx, p, a.max_ctas, stream);
    return a.bn == 64 ? launch_nch<1, true, 1>(ma, mb, md, mx, p, a.max_ctas, stream)
                      : launch_nch<2, true, 1>(ma, mb, md, mx, p, a.max_ctas, stream);
  }
  switch (a.bn) {
    case 64: return launch_nch<1, false, 1>(ma, mb, md, mx, p, a.max_ctas, stream);
    case 128: return launch_nch<2, false, 1>(ma, mb, md, mx, p, a.max_ctas, stream);
    default: return launch_nch<4, false, 1>(ma, mb, md, mx, p, a.max_ctas, stream);
  }
}

}  // namespace ptb
