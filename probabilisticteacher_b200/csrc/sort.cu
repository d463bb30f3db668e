// Segmented stable LSD radix sort (uint32 keys ascending, uint32 payload), one thread-block CLUSTER per segment.
// Replaces torch.sort / the sort inside torchvision nms / torch.randperm-based sampling on the
// proposal path (pt/modeling/proposal_generator/proposal_utils.py:87, fast_rcnn.py:104,
// detectron2 subsample_labels). Segments are a few 10^4 elements and there are only 2-4 of them per launch, so
// one CTA per segment (round 1) left 144 of the 148 SMs idle for 1.5 ms per step. Now the CTAs of a cluster (up
// to 8) each own a contiguous slice of the segment: per pass every CTA histograms its slice, the histograms are
// exchanged through distributed shared memory (one barrier.cluster), every CTA derives the global start of each
// digit for ITS slice (digits before it + same digit in lower-ranked CTAs: stability), and scatters its tiles
// exactly as the single-CTA version did. A second barrier.cluster closes the pass (all scatters visible).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/ptb200.h"

namespace {

constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int ITEMS = 8;                       // elements per thread per tile
constexpr int TILE = SORT_THREADS * ITEMS;     // 4096

// lanes of the warp holding the same 8-bit digit (invalid lanes match nobody): 8 ballots. (match.any.sync
// resolves one distinct value per iteration and stalled ~25 % of the kernel's samples.)
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return valid ? peers : 0u;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// value of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t ld_dsmem_u32(const uint32_t* local, uint32_t rank) {
  uint32_t laddr = static_cast<uint32_t>(__cvta_generic_to_shared(local)), raddr, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory");
  return v;
}

__global__ void __launch_bounds__(SORT_THREADS, 1)
segmented_radix_sort_kernel(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                            int64_t seg_stride, const int* __restrict__ seg_len, int fixed_len,
                            int begin_bit, int end_bit, int csize) {
  __shared__ uint32_t hist[256];  // this CTA's digit histogram of the pass (read by the whole cluster)
  __shared__ uint32_t bin[256];
  __shared__ uint32_t tile_base[256];
  __shared__ uint16_t warp_cnt[SORT_WARPS][256];
  __shared__ uint32_t scan_tmp[8];
  __shared__ uint32_t local_start[256];
  // tile-local reorder buffers: scattering straight to global memory issued one 32-byte sector per key (two
  // uncoalesced 4-byte stores per element were half of the pass time); keys are first placed in digit order in
  // shared memory and then copied out as runs of consecutive addresses
  extern __shared__ uint32_t tile_buf[];
  uint32_t* tile_keys = tile_buf;
  uint32_t* tile_vals = tile_buf + TILE;
  const int seg = blockIdx.x / csize;
  const uint32_t crank = blockIdx.x % csize;  // == %cluster_ctarank (1-D clusters of csize CTAs)
  int n_seg = seg_len != nullptr ? seg_len[seg] : fixed_len;
  if (n_seg > fixed_len) n_seg = fixed_len;
  // this CTA's slice [lo, lo + n) of the segment (slices in rank order: a stable sort of the whole segment)
  const int per = (n_seg + csize - 1) / csize;
  const int lo = min(n_seg, static_cast<int>(crank) * per);
  const int n = min(n_seg, lo + per) - lo;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* kin = keys_a + seg * seg_stride;   // (inputs are read at kin + lo, outputs are segment-relative)
  uint32_t* vin = vals_a + seg * seg_stride;
  uint32_t* kout = keys_b + seg * seg_stride;
  uint32_t* vout = vals_b + seg * seg_stride;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    // warp-aggregated histogram: the high digits of score keys fall into a handful of bins, and one shared-memory
    // atomic per key serialised ~37 000 updates on the same word (the sort took 220 us per launch)
    for (int i0 = warp * 32; i0 < n; i0 += SORT_THREADS) {
      const int i = i0 + lane;
      const bool valid = i < n;
      const uint32_t d = valid ? ((kin[lo + i] >> shift) & 255u) : 0u;
      const uint32_t peers = digit_peers(d, valid);
      if (valid && (peers & lt_mask) == 0u) atomicAdd(&hist[d], static_cast<uint32_t>(__popc(peers)));
    }
    __syncthreads();
    if (csize > 1) cluster_sync_all();  // every CTA's histogram is complete
    // digit totals over the cluster, and the keys with the same digit in lower-ranked CTAs (they go first)
    uint32_t cnt = 0, incl = 0, before = 0;
    if (tid < 256) {
      if (csize > 1) {
        for (uint32_t c = 0; c < static_cast<uint32_t>(csize); ++c) {
          const uint32_t h = ld_dsmem_u32(&hist[tid], c);
          cnt += h;
          if (c < crank) before += h;
        }
      } else {
        cnt = hist[tid];
      }
      // exclusive scan of the 256 digit totals (8 warps)
      incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) scan_tmp[warp] = incl;
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t off = 0;
      for (int w = 0; w < warp; ++w) off += scan_tmp[w];
      bin[tid] = off + incl - cnt + before;
    }
    __syncthreads();

    for (int base = 0; base < n; base += TILE) {
      for (int i = tid; i < SORT_WARPS * 256 / 2; i += SORT_THREADS)
        reinterpret_cast<uint32_t*>(&warp_cnt[0][0])[i] = 0u;
      __syncthreads();
      uint32_t key[ITEMS], val[ITEMS];
      uint16_t rank[ITEMS];
#pragma unroll
      for (int r = 0; r < ITEMS; ++r) {
        const int idx = base + warp * (32 * ITEMS) + r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? kin[lo + idx] : 0u;
        val[r] = valid ? vin[lo + idx] : 0u;
        const uint32_t d = (key[r] >> shift) & 255u;
        const uint32_t peers = digit_peers(d, valid);
        const uint32_t before = warp_cnt[warp][d];
        rank[r] = static_cast<uint16_t>(before + __popc(peers & lt_mask));
        __syncwarp();
        if (valid && (peers & lt_mask) == 0u) warp_cnt[warp][d] = static_cast<uint16_t>(before + __popc(peers));
        __syncwarp();
      }
      __syncthreads();
      uint32_t run = 0, lincl = 0;
      if (tid < 256) {
#pragma unroll 8
        for (int w = 0; w < SORT_WARPS; ++w) {
          const uint32_t c = warp_cnt[w][tid];
          warp_cnt[w][tid] = static_cast<uint16_t>(run);
          run += c;
        }
        // exclusive scan of the tile's digit counts -> start of each digit's run inside the tile
        lincl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, lincl, o);
          if (lane >= o) lincl += t;
        }
        if (lane == 31) scan_tmp[warp] = lincl;
      }
      __syncthreads();
      if (tid < 256) {
        uint32_t off = 0;
        for (int w = 0; w < warp; ++w) off += scan_tmp[w];
        const uint32_t b0 = bin[tid];
        local_start[tid] = off + lincl - run;
        tile_base[tid] = b0;
        bin[tid] = b0 + run;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < ITEMS; ++r) {
        const int idx = base + warp * (32 * ITEMS) + r * 32 + lane;
        if (idx < n) {
          const uint32_t d = (key[r] >> shift) & 255u;
          const uint32_t lp = local_start[d] + warp_cnt[warp][d] + rank[r];
          tile_keys[lp] = key[r];
          tile_vals[lp] = val[r];
        }
      }
      __syncthreads();
      const int tile_n = min(TILE, n - base);
      for (int i = tid; i < tile_n; i += SORT_THREADS) {
        const uint32_t k = tile_keys[i];
        const uint32_t d = (k >> shift) & 255u;
        const uint32_t pos = tile_base[d] + (static_cast<uint32_t>(i) - local_start[d]);
        kout[pos] = k;
        vout[pos] = tile_vals[i];
      }
      __syncthreads();
    }
    uint32_t* t = kin;
    kin = kout;
    kout = t;
    t = vin;
    vin = vout;
    vout = t;
    __syncthreads();
    if (csize > 1) {  // the next pass reads what the other CTAs of the cluster scattered, and rewrites `hist`
      __threadfence();
      cluster_sync_all();
    }
  }
}

}  // namespace

// Sorts `segments` segments in place (result in keys/vals when the pass count is even, which the
// entry point enforces). keys_tmp / vals_tmp are scratch of the same size.
extern "C" int ptb200_segmented_sort_u32(uint32_t* keys, uint32_t* vals, uint32_t* keys_tmp, uint32_t* vals_tmp,
                                         int segments, int64_t seg_stride, const int* seg_len_dev, int max_len,
                                         int begin_bit, int end_bit, void* stream) {
  if (segments <= 0) return 0;
  if ((end_bit - begin_bit) % 16 != 0 || begin_bit < 0 || end_bit > 32) return 1301;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(segmented_radix_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         2 * TILE * 4);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  // CTAs per segment: one per ~2 tiles of 4096 keys, at most the portable cluster size of 8
  static int csize_env = -1;
  if (csize_env < 0) {
    const char* e = getenv("PTB200_SORT_CLUSTER");
    csize_env = e ? atoi(e) : 0;
  }
  int csize = csize_env > 0 ? csize_env : (max_len + TILE - 1) / TILE;
  if (csize > 8) csize = 8;
  if (csize < 1) csize = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(segments * csize, 1, 1);
  cfg.blockDim = dim3(SORT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = 2 * TILE * 4;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int64_t stride = seg_stride;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, segmented_radix_sort_kernel, keys, vals, keys_tmp, vals_tmp, stride,
                                             seg_len_dev, max_len, begin_bit, end_bit, csize));
}
