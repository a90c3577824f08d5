// vmap_sort.cuh -- VMAP_DEDUPE_SORT: the key sets of one scan built the way BASELINE.json's
// north_star spells it out -- every ray-voxel traversal is emitted as a packed key, the keys are
// radix-sorted, and "first of each run" is the unique voxel, with occupied winning over free
// because the key is (morton48 << 1) | is_free (an occupied key sorts directly before the free
// key of the same voxel).  The unique voxels are then marked in the per-block masks and the
// ordinary K4 update applies them, so both dedupe modes share one update path.
//
// All kernels here are hand-written (no CUB): emit (two-pass: count, then write), an 8-bit LSD
// radix sort (histogram / scan / stable scatter per pass), and the run-head ("unique") kernel.
#pragma once

#include "vmap_kernels.cuh"

namespace vmapdev {

constexpr uint32_t kFlagOcc = 8u;        // set by k_resolve: this point inserts an occupied key
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;           // keys per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortKeyBits = 49;         // 48 Morton bits + the free/occupied bit
constexpr int kSortPasses = (kSortKeyBits + kRadixBits - 1) / kRadixBits;   // 7

struct SortBuffers {
  unsigned long long* keys_a;   // [cap]
  unsigned long long* keys_b;   // [cap]
  uint32_t* ray_off;            // [scan_cap] first key slot of each cast ray
  uint32_t* hist;               // [kRadix * n_tiles] digit-major tile histograms
  uint32_t* bin_total;          // [kRadix]
  size_t cap;                   // keys
  size_t ray_cap;
  size_t hist_cap;              // entries
};

// 16 -> 48 bit spreading (bit i -> bit 3i)
__host__ __device__ __forceinline__ unsigned long long spread3(uint32_t v) {
  unsigned long long x = v & 0xFFFFull;
  x = (x | (x << 16)) & 0x0000FF0000FFull;
  x = (x | (x << 8)) & 0x00F00F00F00Full;
  x = (x | (x << 4)) & 0x0C30C30C30C3ull;
  x = (x | (x << 2)) & 0x249249249249ull;
  return x;
}
__host__ __device__ __forceinline__ uint32_t compact3(unsigned long long x) {
  x &= 0x249249249249ull;
  x = (x | (x >> 2)) & 0x0C30C30C30C3ull;
  x = (x | (x >> 4)) & 0x00F00F00F00Full;
  x = (x | (x >> 8)) & 0x0000FF0000FFull;
  x = (x | (x >> 16)) & 0xFFFFull;
  return (uint32_t)x;
}
// Morton order == octree depth-first order (x is the lowest bit of each triple)
__host__ __device__ __forceinline__ unsigned long long morton48(uint32_t kx, uint32_t ky, uint32_t kz) {
  return spread3(kx) | (spread3(ky) << 1) | (spread3(kz) << 2);
}

// ---- emit -------------------------------------------------------------------------------------
// pass 1: number of keys of every cast ray and its slot range in the key buffer
template <bool kFilter>
__global__ void __launch_bounds__(128)
k_walk_count(DevParams P, ScanBuffers sb, Counters* cnt, uint32_t* __restrict__ ray_off, float ox,
             float oy, float oz) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  RayWalk w;
  w.live = false;
  if (r < cnt->n_rays) {
    const uint32_t i = sb.rays[r];
    w.init(P, ox, oy, oz, sb.ray_end[3 * i], sb.ray_end[3 * i + 1], sb.ray_end[3 * i + 2]);
  }
  uint32_t count = 0, kept = 0;
  while (w.live) {
    ++count;
    if (!kFilter || free_filter_pass(P, ox, oy, oz, w.c0, w.c1, w.c2)) ++kept;
    if (!w.advance()) break;
    if (count >= kKeyRayMax - 1u) break;
  }
  // slot allocation: order is irrelevant (the keys are sorted next), one atomic per warp
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t incl = kept;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += v;
  }
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  uint32_t base = 0;
  if (lane == 0 && total) base = atomicAdd(&cnt->n_keys, total);
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (r < cnt->n_rays) ray_off[r] = base + incl - kept;
  uint32_t sum = count, mx = count;
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  if (lane == 0 && sum) {
    atomicAdd(&cnt->traversals, (unsigned long long)sum);
    atomicMax(&cnt->longest_ray, mx);
  }
}

// pass 2: write (morton48 << 1) | 1 for every free key of every cast ray
template <bool kFilter>
__global__ void __launch_bounds__(128)
k_walk_emit(DevParams P, ScanBuffers sb, Counters* cnt, const uint32_t* __restrict__ ray_off,
            unsigned long long* __restrict__ keys, uint32_t cap, float ox, float oy, float oz) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cnt->n_rays) return;
  const uint32_t i = sb.rays[r];
  RayWalk w;
  w.init(P, ox, oy, oz, sb.ray_end[3 * i], sb.ray_end[3 * i + 1], sb.ray_end[3 * i + 2]);
  uint32_t count = 0, pos = ray_off[r];
  while (w.live) {
    ++count;
    if (!kFilter || free_filter_pass(P, ox, oy, oz, w.c0, w.c1, w.c2)) {
      if (pos < cap) keys[pos] = (morton48(w.c0, w.c1, w.c2) << 1) | 1ull;
      ++pos;
    }
    if (!w.advance()) break;
    if (count >= kKeyRayMax - 1u) break;
  }
}

// occupied endpoints (flagged by k_resolve) -> (morton48 << 1) | 0, appended after the free keys
__global__ void __launch_bounds__(256)
k_emit_occupied(ScanBuffers sb, Counters* cnt, uint32_t n, unsigned long long* __restrict__ keys,
                uint32_t n_free, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(sb.flags[i] & kFlagOcc)) return;
  const unsigned long long k = sb.key[i];
  const uint32_t pos = n_free + atomicAdd(&cnt->occ_cursor, 1u);
  if (pos < cap)
    keys[pos] = morton48((uint32_t)(k & 0xFFFFu), (uint32_t)((k >> 16) & 0xFFFFu), (uint32_t)((k >> 32) & 0xFFFFu)) << 1;
}

// ---- 8-bit LSD radix sort ---------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_radix_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist,
             uint32_t n_tiles) {
  __shared__ uint32_t sh[kRadix];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t tile = blockIdx.x;
  const uint32_t base = tile * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortItems; j++) {
    const uint32_t idx = base + j * kSortThreads + threadIdx.x;
    if (idx < n) atomicAdd(&sh[(uint32_t)(keys[idx] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * n_tiles + tile] = sh[threadIdx.x];
}

// one block per digit: exclusive scan of that digit's tile counts, digit total
__global__ void __launch_bounds__(256)
k_radix_scan_tiles(uint32_t* __restrict__ hist, uint32_t n_tiles, uint32_t* __restrict__ bin_total) {
  __shared__ uint32_t warp_sum[8];
  __shared__ uint32_t carry_s;
  uint32_t* row = hist + (size_t)blockIdx.x * n_tiles;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += 256) {
    const uint32_t idx = base + threadIdx.x;
    const uint32_t v = idx < n_tiles ? row[idx] : 0u;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w2 = 0; w2 < warp; w2++) woff += warp_sum[w2];
    const uint32_t carry = carry_s;
    if (idx < n_tiles) row[idx] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) bin_total[blockIdx.x] = carry_s;
}

// stable scatter of one tile
__global__ void __launch_bounds__(kSortThreads)
k_radix_scatter(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, uint32_t n,
                int shift, const uint32_t* __restrict__ hist, uint32_t n_tiles,
                const uint32_t* __restrict__ bin_total) {
  __shared__ uint32_t digit_base[kRadix];          // global position of the tile's next key per digit
  __shared__ uint32_t warp_cnt[kSortThreads / 32][kRadix];
  const uint32_t tile = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  {
    // exclusive scan of the digit totals (256 values, one per thread)
    __shared__ uint32_t ws[8];
    const uint32_t v = bin_total[threadIdx.x];
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w2 = 0; w2 < warp; w2++) woff += ws[w2];
    digit_base[threadIdx.x] = woff + incl - v + hist[threadIdx.x * n_tiles + tile];
  }
  const uint32_t base = tile * kSortTile;
  for (int j = 0; j < kSortItems; j++) {
    for (int w2 = 0; w2 < kSortThreads / 32; w2++) warp_cnt[w2][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t idx = base + j * kSortThreads + threadIdx.x;
    const bool valid = idx < n;
    unsigned long long key = 0;
    uint32_t d = 0, rank_in_warp = 0;
    if (valid) {
      key = in[idx];
      d = (uint32_t)(key >> shift) & (kRadix - 1);
    }
    const unsigned act = __ballot_sync(0xFFFFFFFFu, valid);
    if (valid) {
      const unsigned peers = __match_any_sync(act, d);
      rank_in_warp = __popc(peers & ((1u << lane) - 1u));
      if (rank_in_warp == 0) warp_cnt[warp][d] = __popc(peers);
    }
    __syncthreads();
    if (valid) {
      uint32_t before = 0;
      for (uint32_t w2 = 0; w2 < warp; w2++) before += warp_cnt[w2][d];
      out[digit_base[d] + before + rank_in_warp] = key;
    }
    __syncthreads();
    {
      uint32_t tot = 0;
      for (int w2 = 0; w2 < kSortThreads / 32; w2++) tot += warp_cnt[w2][threadIdx.x];
      digit_base[threadIdx.x] += tot;
    }
    __syncthreads();
  }
}

// ---- unique with occupied-wins ------------------------------------------------------------------
// Head of every run of equal voxels: the head's low bit tells occupied (0) or free (1).
__global__ void __launch_bounds__(256)
k_unique_mark(MapView m, Counters* cnt, const unsigned long long* __restrict__ keys, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  if (i > 0 && (keys[i - 1] >> 1) == (k >> 1)) return;
  const unsigned long long mo = k >> 1;
  const uint32_t kx = compact3(mo), ky = compact3(mo >> 1), kz = compact3(mo >> 2);
  mark_voxel(m, cnt, (k & 1ull) ? m.free_mask : m.occ_mask, kx, ky, kz);
}

}  // namespace vmapdev
