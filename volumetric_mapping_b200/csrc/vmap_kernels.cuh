// vmap_kernels.cuh -- the sm_100a kernels of the scan-insertion path (bitmask key-set mode) and
// of the batched queries.  Launch wrappers live in vmap.cu.
//
//   K1   k_front_cloud / k_front_projected / k_front_disparity
//          sensor->world transform (+ disparity projection), endpoint key, range classification,
//          clip to max range, and registration in the per-scan "first index" table.
//   K1b  k_resolve     the order-dependent skip rule (octomap_world.cc:131-136) in closed form,
//          cast-ray work list, occupied endpoint marks.
//   K2   k_walk_mask   one thread per cast ray: Amanatides-Woo DDA, each visited voxel is one
//          RED.OR into its block's 512-bit free mask (the KeySet insert of the reference).
//   K4   k_update      one warp per touched block: free &= ~occupied, one clamped float
//          log-odds update per set bit, masks cleared for the next scan.
//   K5   k_query_points / k_query_lines
#pragma once

#include "vmap_device.cuh"
#include "vmap_seq.cuh"
#include "vmap_walk.cuh"
#include "vmap_query.cuh"
#include "vmap_front.cuh"

namespace vmapdev {

constexpr uint32_t kFlagValid = 1u;   // point takes part in the scan
constexpr uint32_t kFlagRange = 2u;   // castRay's in-range branch
constexpr uint32_t kFlagBound = 4u;   // coordToKeyChecked(point) succeeds
//                  kFlagOcc = 8u       (vmap_sort.cuh) the point inserts an occupied key
constexpr uint32_t kFlagCast = 16u;   // castRay is invoked for this point (set by k_resolve)
struct ScanBuffers {
  float* world;             // [n*3] world-frame points (what the reference leaves in the cloud)
  float* ray_end;           // [n*3] ray end: the point, or the clipped end when out of range
  unsigned long long* key;  // [n]   unchecked endpoint key k0 | k1<<16 | k2<<32
  uint8_t* flags;           // [n]
  uint32_t* rays;           // [n]   indices of cast rays, longest rays first
  uint32_t* len;            // [n]   DDA steps of the ray (key-space Manhattan distance)
  uint32_t* rank;           // [n]   position of a cast ray inside its length bin
  uint32_t* len_hist;       // [kFineBins] cast rays per fine bin (vmap_walk.cuh: fine_bin)
  uint32_t* seg_base;       // [kLenBins + 1] first work slot of every segment level
  uint32_t* bin_off;        // [kFineBins] first list position of every fine bin (k_resolve_fold)
  RayRec* ray_rec;          // [n]
  Checkpoint* ckpt;         // [ckpt_cap] one per (segment level, ray)
  uint32_t ckpt_cap;
  unsigned long long* first_keys;  // [first_cap] per-scan table: endpoint key -> first index
  uint32_t* first_idx;             // [first_cap]
  uint32_t first_mask;
};

// ---------------------------------------------------------------------------------------------
// per-scan first-index table (endpoint key -> smallest point index that occupies it)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_key48(unsigned long long k) {
  k ^= k >> 29; k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 32;
  return (uint32_t)k;
}
__device__ __forceinline__ void first_insert(const ScanBuffers& sb, unsigned long long key,
                                             uint32_t idx) {
  uint32_t h = hash_key48(key) & sb.first_mask;
  for (;;) {  // table is sized >= 2n: always terminates
    unsigned long long w = __ldcg(sb.first_keys + h);
    if (w == kEmpty) {
      w = atomicCAS(sb.first_keys + h, kEmpty, key);
      if (w == kEmpty) w = key;
    }
    if (w == key) {
      atomicMin(sb.first_idx + h, idx);
      return;
    }
    h = (h + 1u) & sb.first_mask;
  }
}
__device__ __forceinline__ uint32_t first_lookup(const ScanBuffers& sb, unsigned long long key) {
  uint32_t h = hash_key48(key) & sb.first_mask;
  for (;;) {
    const unsigned long long w = sb.first_keys[h];
    if (w == kEmpty) return 0xFFFFFFFFu;
    if (w == key) return sb.first_idx[h];
    h = (h + 1u) & sb.first_mask;
  }
}

// Common tail of the three front-ends: classify world point `p` of index i.
__device__ __forceinline__ void classify_point(const DevParams& P, const ScanBuffers& sb,
                                               Counters* cnt, const float* origin, uint32_t i,
                                               float px, float py, float pz, uint32_t* s_acc = nullptr) {
  sb.world[3 * i + 0] = px;
  sb.world[3 * i + 1] = py;
  sb.world[3 * i + 2] = pz;
  const PointClass c = classify_math(P, origin, px, py, pz);
  const unsigned long long key = c.key;
  const bool bound = c.bound, in_range = c.in_range;
  sb.ray_end[3 * i + 0] = c.ex;
  sb.ray_end[3 * i + 1] = c.ey;
  sb.ray_end[3 * i + 2] = c.ez;
  sb.len[i] = c.len;
  sb.key[i] = key;
  sb.flags[i] = (uint8_t)(kFlagValid | (in_range ? kFlagRange : 0u) | (bound ? kFlagBound : 0u));
  if (in_range && bound) first_insert(sb, key, i);
  // statistics (warp-aggregated)
  const unsigned act = __activemask();
  const unsigned in_b = __ballot_sync(act, in_range);
  if ((threadIdx.x & 31u) == (uint32_t)(__ffs(act) - 1)) {
    // (s_acc: the CTA's own pair of counters in shared memory -- the caller adds them to the scan's once)
    atomicAdd(s_acc ? s_acc : &cnt->points_valid, (uint32_t)__popc(act));
    if (in_b) atomicAdd(s_acc ? s_acc + 1 : &cnt->points_in_range, (uint32_t)__popc(in_b));
  }
}

// K1 (cloud): pcl::removeNaNFromPointCloud + pcl::transformPointCloud(Matrix4d) + classify.
__global__ void __launch_bounds__(256)
k_front_cloud(DevParams P, ScanBuffers sb, Counters* cnt, Transform T, const uint8_t* __restrict__ in,
              uint32_t n, uint32_t stride, int is_dense) {
  pdl_launch();
  // statistics: one pair of global atomics per CTA (thousands of per-warp atomics on one address
  // serialise in the L2 and show up as a tail on a kernel this short)
  __shared__ uint32_t s_acc[2];
  if (threadIdx.x < 2) s_acc[threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float* p = reinterpret_cast<const float*>(in + (size_t)i * stride);
    const float fx = p[0], fy = p[1], fz = p[2];
    if (!is_dense && !(isfinite(fx) && isfinite(fy) && isfinite(fz))) {
      sb.flags[i] = 0;
      sb.world[3 * i] = fx; sb.world[3 * i + 1] = fy; sb.world[3 * i + 2] = fz;
    } else {
      float wx, wy, wz;
      matrix_transform(T, fx, fy, fz, &wx, &wy, &wz);
      classify_point(P, sb, cnt, T.origin, i, wx, wy, wz, s_acc);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_acc[0]) atomicAdd(&cnt->points_valid, s_acc[0]);
    if (s_acc[1]) atomicAdd(&cnt->points_in_range, s_acc[1]);
  }
}

__device__ __forceinline__ void projected_point(const DevParams& P, const ScanBuffers& sb,
                                                Counters* cnt, const Transform& T, uint32_t i,
                                                float sx, float sy, float sz) {
  // isValidPoint(pt) && !(pt.z < 0)   (octomap_world.cc:156, 232-236)
  const bool valid = projected_valid(sz);
  if (!valid) {
    sb.flags[i] = 0;
    return;
  }
  double wx, wy, wz;
  quat_transform(T, (double)sx, (double)sy, (double)sz, &wx, &wy, &wz);
  classify_point(P, sb, cnt, T.origin, i, (float)wx, (float)wy, (float)wz);
}

// K1 (projected image): CV_32FC3 rows with a pitch.
__global__ void __launch_bounds__(256)
k_front_projected(DevParams P, ScanBuffers sb, Counters* cnt, Transform T,
                  const uint8_t* __restrict__ img, int rows, int cols, size_t pitch) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)rows * (uint32_t)cols) return;
  const uint32_t v = i / (uint32_t)cols, u = i - v * (uint32_t)cols;
  const float* p = reinterpret_cast<const float*>(img + (size_t)v * pitch) + 3 * u;
  projected_point(P, sb, cnt, T, i, p[0], p[1], p[2]);
}

// image minimum for reprojectImageTo3D(handleMissingValues=true)
__global__ void __launch_bounds__(256)
k_min_disparity(Counters* cnt, const uint8_t* __restrict__ img, int rows, int cols, size_t pitch) {
  uint32_t best = 0xFFFFFFFFu;
  const uint32_t total = (uint32_t)rows * (uint32_t)cols;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t v = i / (uint32_t)cols, u = i - v * (uint32_t)cols;
    const float d = reinterpret_cast<const float*>(img + (size_t)v * pitch)[u];
    if (d == d) best = min(best, float_to_ordered(d));
  }
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
  if ((threadIdx.x & 31) == 0 && best != 0xFFFFFFFFu) atomicMin(&cnt->min_disparity_ord, best);
}

// K1 (disparity): projection fused with validity test, quaternion transform and classification.
__global__ void __launch_bounds__(256)
k_front_disparity(DevParams P, ScanBuffers sb, Counters* cnt, Transform T, QMat Q,
                  const uint8_t* __restrict__ img, int rows, int cols, size_t pitch,
                  float* __restrict__ xyz_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)rows * (uint32_t)cols) return;
  const uint32_t v = i / (uint32_t)cols, u = i - v * (uint32_t)cols;
  const float d = reinterpret_cast<const float*>(img + (size_t)v * pitch)[u];
  const uint32_t mo = cnt->min_disparity_ord;
  const double min_disp = (mo == 0xFFFFFFFFu) ? 3.4028234663852886e38 : (double)ordered_to_float(mo);
  float X, Y, Z;
  reproject_pixel(Q.q, min_disp, u, v, d, &X, &Y, &Z);
  if (xyz_out) {  // projection-only entry point (vmap_reproject)
    xyz_out[3 * i] = X; xyz_out[3 * i + 1] = Y; xyz_out[3 * i + 2] = Z;
    return;
  }
  projected_point(P, sb, cnt, T, i, X, Y, Z);
}

// mark voxel (kx,ky,kz) in a block mask array (table mode)
__device__ __forceinline__ void mark_voxel(const MapView& m, Counters* cnt, uint32_t* masks,
                                           uint32_t kx, uint32_t ky, uint32_t kz) {
  const uint32_t h = find_or_insert_touch(m, cnt, block_key(kx, ky, kz));
  if (h == kNotFound) return;
  const uint32_t l = local_index(kx, ky, kz);
  atomicOr(masks + (size_t)h * kMaskWords + (l >> 5), 1u << (l & 31u));
}
// the same in the dense window (occupied endpoints)
__device__ __forceinline__ void mark_voxel_dense(const MarkSet& ms, Counters* cnt, uint32_t* masks,
                                                 uint32_t kx, uint32_t ky, uint32_t kz) {
  const uint32_t idx = window_index(ms, kx, ky, kz);
  if (idx == kNotFound) { atomicExch(&cnt->overflow_window, 1u); return; }
  const uint32_t l = local_index(kx, ky, kz);
  atomicOr(masks + (size_t)idx * kMaskWords + (l >> 5), 1u << (l & 31u));
  atomicOr(ms.touched_bits + (idx >> 5), 1u << (idx & 31u));
}

// Statistics counters are bumped once per CTA, not once per warp: thousands of atomics on one
// address serialise in the L2 and show up as a tail on short kernels.
__device__ __forceinline__ void cta_add_u32(uint32_t* s_acc, uint32_t v) {   // s_acc: shared, zeroed
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31u) == 0 && v) atomicAdd(s_acc, v);
}
__device__ __forceinline__ void cta_max_u32(uint32_t* s_acc, uint32_t v) {
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  if ((threadIdx.x & 31u) == 0 && v) atomicMax(s_acc, v);
}

__device__ __forceinline__ uint32_t ray_bin(const ScanBuffers& sb, uint32_t i) {
  return fine_bin(sb.len[i]);
}

// K1b: skip rule + work list + occupied marks.
//   occupied_cells = { K_i : valid, in range, in bounds };  first[K] = min such i.
//   castRay(i) is invoked  <=>  K_i not yet in occupied_cells when point i is visited
//                           <=>  in-range&bound ? i == first[K_i] : i < first[K_i].
__global__ void __launch_bounds__(256)
k_resolve(ScanBuffers sb, MapView m, MarkSet ms, Counters* cnt, uint32_t n, int mark_occupied) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool cast = false, occ = false;
  unsigned long long key = 0;
  if (i < n) {
    const uint32_t f = sb.flags[i];
    if (f & kFlagValid) {
      key = sb.key[i];
      const uint32_t first = first_lookup(sb, key);
      const bool rb = (f & kFlagRange) && (f & kFlagBound);
      cast = rb ? (i == first) : (i < first);
      occ = rb && cast;
    }
  }
  if (occ) {
    if (mark_occupied && ms.dense)
      mark_voxel_dense(ms, cnt, ms.occ_mask, (uint32_t)(key & 0xFFFFu), (uint32_t)((key >> 16) & 0xFFFFu),
                       (uint32_t)((key >> 32) & 0xFFFFu));
    else if (mark_occupied)
      mark_voxel(m, cnt, m.occ_mask, (uint32_t)(key & 0xFFFFu), (uint32_t)((key >> 16) & 0xFFFFu),
                 (uint32_t)((key >> 32) & 0xFFFFu));
    else
      sb.flags[i] |= 8u;   // kFlagOcc (vmap_sort.cuh): emitted as a packed key by the sort mode
  }
  // cast-ray work list, pass 1: position of the ray inside its length bin
  // (only a handful of bins are populated: one atomic per distinct bin per warp)
  const unsigned cast_b = __ballot_sync(0xFFFFFFFFu, cast);
  if (cast) {
    sb.flags[i] |= kFlagCast;
    const uint32_t bin = ray_bin(sb, i);
    const unsigned peers = __match_any_sync(cast_b, bin);
    const uint32_t lane_ = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane_ == (uint32_t)(__ffs(peers) - 1)) base = atomicAdd(sb.len_hist + bin, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, __ffs(peers) - 1);
    sb.rank[i] = base + __popc(peers & ((1u << lane_) - 1u));
  }
  const unsigned occ_b = __ballot_sync(0xFFFFFFFFu, occ);
  if ((threadIdx.x & 31u) == 0 && cast_b) {
    atomicAdd(&cnt->n_rays, (uint32_t)__popc(cast_b));
    atomicAdd(&cnt->rays_cast, (uint32_t)__popc(cast_b));
    if (occ_b) atomicAdd(&cnt->occupied_emitted, (uint32_t)__popc(occ_b));
  }
}

// Descending exclusive prefix of the fine-bin histogram (a CTA of 256 threads): off_s[b] = number of
// cast rays in bins above b, i.e. the position of bin b's first ray in the ordered list (longest
// rays first, so that the 32 rays of a warp take about the same number of DDA steps).
template <bool kGlobalSource = true>
__device__ __forceinline__ void cta_bin_offsets(const uint32_t* len_hist, uint32_t* off_s, uint32_t* warp_sum) {
  constexpr int kPer = (kFineBins + 255) / 256;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // thread t owns the kPer bins starting at descending position t * kPer
  uint32_t local[kPer], sum = 0;
#pragma unroll
  for (int j = 0; j < kPer; j++) {
    const int pos = threadIdx.x * kPer + j;
    local[j] = pos < kFineBins ? (kGlobalSource ? __ldcg(len_hist + (kFineBins - 1 - pos)) : len_hist[kFineBins - 1 - pos]) : 0u;
    sum += local[j];
  }
  uint32_t incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  uint32_t base = incl - sum;
  for (uint32_t w2 = 0; w2 < warp; w2++) base += warp_sum[w2];
#pragma unroll
  for (int j = 0; j < kPer; j++) {
    const int pos = threadIdx.x * kPer + j;
    if (pos < kFineBins) off_s[kFineBins - 1 - pos] = base;
    base += local[j];
  }
  __syncthreads();
}
// Work slots of the segment walker, level-major: level j holds one slot for every ray with more
// than j segments -- a prefix of the ordered list, as long as the rays that precede the highest
// fine bin of segment count j.  (CTA of 256 threads; off_s from cta_bin_offsets.)
__device__ __forceinline__ void cta_seg_base(const uint32_t* off_s, uint32_t* warp_sum, uint32_t* seg_base) {
  constexpr int kLv = kLenBins / 256;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t lv[kLv], lsum = 0;
#pragma unroll
  for (int j = 0; j < kLv; j++) {
    lv[j] = off_s[fine_bin_top(threadIdx.x * kLv + j)];
    lsum += lv[j];
  }
  uint32_t linc = lsum;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, linc, o);
    if (lane >= (uint32_t)o) linc += t;
  }
  __syncthreads();
  if (lane == 31) warp_sum[warp] = linc;
  __syncthreads();
  uint32_t lbase = linc - lsum;
  for (uint32_t w2 = 0; w2 < warp; w2++) lbase += warp_sum[w2];
#pragma unroll
  for (int j = 0; j < kLv; j++) {
    seg_base[threadIdx.x * kLv + j] = lbase;
    lbase += lv[j];
  }
  if (threadIdx.x == 255) seg_base[kLenBins] = lbase;
}

// K1b for the segment walker's chain: k_resolve with the work-list bookkeeping folded in.
//   * position inside a bin: one shared-memory counter per bin and CTA (warp-aggregated), then ONE
//     global atomic per populated bin and CTA -- k_resolve's per-warp returning atomics on a dozen
//     hot histogram words were 45 % of its stall samples (profiles/README.md, round 2);
//   * the CTA also leaves its own cast rays, longest first, in sb.rays[256 * blockIdx.x ...] (padded
//     with kNoRay): k_checkpoints_pts runs CTA for CTA over these lists, so its warps get rays of
//     about the same segment count without a globally ordered list;
//   * the last CTA to finish turns the finished histogram into sb.bin_off (first list position of
//     every bin) and sb.seg_base, so no k_order_rays launch follows: a ray's position in the ordered
//     list is bin_off[bin] + rank.
constexpr uint32_t kNoRay = 0xFFFFFFFFu;
__global__ void __launch_bounds__(256)
k_resolve_fold(ScanBuffers sb, MapView m, MarkSet ms, Counters* cnt, uint32_t n, int mark_occupied) {
  __shared__ uint32_t s_hist[kFineBins];     // the CTA's rays per bin, then their first position inside the bin
  __shared__ uint32_t s_off[kFineBins];      // rays of the CTA in higher bins
  __shared__ uint32_t s_stat[3];
  __shared__ uint32_t warp_sum[8];
  __shared__ bool s_last;
  for (int k = threadIdx.x; k < kFineBins; k += 256) s_hist[k] = 0u;
  if (threadIdx.x < 3) s_stat[threadIdx.x] = 0u;
  pdl_wait();
  pdl_launch();
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane_ = threadIdx.x & 31u;
  bool cast = false, occ = false;
  unsigned long long key = 0;
  uint32_t f = 0;
  if (i < n) {
    f = sb.flags[i];
    if (f & kFlagValid) {
      key = sb.key[i];
      const uint32_t first = first_lookup(sb, key);
      const bool rb = (f & kFlagRange) && (f & kFlagBound);
      cast = rb ? (i == first) : (i < first);
      occ = rb && cast;
    }
  }
  if (occ) {
    if (mark_occupied && ms.dense)
      mark_voxel_dense(ms, cnt, ms.occ_mask, (uint32_t)(key & 0xFFFFu), (uint32_t)((key >> 16) & 0xFFFFu),
                       (uint32_t)((key >> 32) & 0xFFFFu));
    else if (mark_occupied)
      mark_voxel(m, cnt, m.occ_mask, (uint32_t)(key & 0xFFFFu), (uint32_t)((key >> 16) & 0xFFFFu),
                 (uint32_t)((key >> 32) & 0xFFFFu));
    else
      f |= 8u;   // kFlagOcc
  }
  const unsigned cast_b = __ballot_sync(0xFFFFFFFFu, cast);
  uint32_t bin = 0, local = 0;
  if (cast) {
    sb.flags[i] = (uint8_t)(f | kFlagCast);
    bin = ray_bin(sb, i);
    const unsigned peers = __match_any_sync(cast_b, bin);
    const int leader = __ffs(peers) - 1;
    if (lane_ == (uint32_t)leader) local = atomicAdd(&s_hist[bin], (uint32_t)__popc(peers));
    local = __shfl_sync(peers, local, leader) + __popc(peers & ((1u << lane_) - 1u));
  }
  const unsigned occ_b = __ballot_sync(0xFFFFFFFFu, occ);
  if (lane_ == 0 && cast_b) {
    atomicAdd(&s_stat[0], (uint32_t)__popc(cast_b));
    if (occ_b) atomicAdd(&s_stat[1], (uint32_t)__popc(occ_b));
  }
  __syncthreads();
  cta_bin_offsets<false>(s_hist, s_off, warp_sum);
  // the CTA's share of every populated bin -> its first position inside the bin
  for (int k = threadIdx.x; k < kFineBins; k += 256) {
    const uint32_t c = s_hist[k];
    if (c) s_hist[k] = atomicAdd(sb.len_hist + k, c);
  }
  __syncthreads();
  if (cast) {
    sb.rank[i] = s_hist[bin] + local;
    sb.rays[blockIdx.x * 256u + s_off[bin] + local] = i;
  }
  if (threadIdx.x >= s_stat[0]) sb.rays[blockIdx.x * 256u + threadIdx.x] = kNoRay;
  if (threadIdx.x == 0) {
    if (s_stat[0]) {
      atomicAdd(&cnt->n_rays, s_stat[0]);
      atomicAdd(&cnt->rays_cast, s_stat[0]);
    }
    if (s_stat[1]) atomicAdd(&cnt->occupied_emitted, s_stat[1]);
    __threadfence();
    s_last = atomicAdd(&cnt->done_resolve, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  cta_bin_offsets(sb.len_hist, s_hist, warp_sum);
  for (int k = threadIdx.x; k < kFineBins; k += 256) sb.bin_off[k] = s_hist[k];
  cta_seg_base(s_hist, warp_sum, sb.seg_base);
}

// Cast-ray work list, pass 2 (walkers that read the ordered list sb.rays: table mode, k_walk_dense,
// sort mode; the segment walker's chain gets its positions from k_resolve_fold instead).
__global__ void __launch_bounds__(256) k_order_rays(ScanBuffers sb, uint32_t n) {
  __shared__ uint32_t off[kFineBins];
  __shared__ uint32_t warp_sum[8];
  cta_bin_offsets(sb.len_hist, off, warp_sum);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (sb.flags[i] & kFlagCast))
    sb.rays[off[ray_bin(sb, i)] + sb.rank[i]] = i;
  if (blockIdx.x != 0) return;
  cta_seg_base(off, warp_sum, sb.seg_base);
}

// K2: one thread per cast ray.  Every visited voxel is one bit of its block's 512-bit free mask
// (the KeySet insert of the reference).  The warp alternates between two converged phases:
//   (A) every live lane looks up / inserts the block of its pending voxel (hash probe, all lanes
//       in flight together), then
//   (B) walks the voxels of the ray inside that block with the branch-free DDA, collecting the
//       bits of one 32-voxel mask word in a register and flushing it with a single RED.OR when
//       the ray moves on to another word.
template <bool kFilter>
__global__ void __launch_bounds__(128)
k_walk_mask(DevParams P, ScanBuffers sb, MapView m, Counters* cnt, float ox, float oy, float oz) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  RayWalk w;
  w.live = false;
  if (r < cnt->n_rays) {
    const uint32_t i = sb.rays[r];
    w.init(P, ox, oy, oz, sb.ray_end[3 * i], sb.ray_end[3 * i + 1], sb.ray_end[3 * i + 2]);
  }
  uint32_t count = 0;
  while (__any_sync(0xFFFFFFFFu, w.live)) {
    // (A)
    uint32_t* mask = nullptr;
    if (w.live) {
      const uint32_t h = find_or_insert_touch(m, cnt, block_key(w.c0, w.c1, w.c2));
      if (h == kNotFound) w.live = false;   // table full: the host grows it and re-runs the scan
      else mask = m.free_mask + (size_t)h * kMaskWords;
    }
    // (B)
    if (w.live) {
      const uint32_t b0 = w.c0 >> 3, b1 = w.c1 >> 3, b2 = w.c2 >> 3;
      uint32_t word = 0xFFFFFFFFu, bits = 0, seen = 0;   // `seen`: see k_walk_dense
      for (;;) {
        ++count;
        if (!kFilter || free_filter_pass(P, ox, oy, oz, w.c0, w.c1, w.c2)) {
          const uint32_t l = local_index(w.c0, w.c1, w.c2);
          if ((l >> 5) != word) {
            if (bits & ~seen) atomicOr(mask + word, bits);
            word = l >> 5;
            bits = 0;
            seen = mask[word];
          }
          bits |= 1u << (l & 31u);
        }
        if (!w.advance()) break;
        if (count >= kKeyRayMax - 1u) { w.live = false; break; }   // KeyRay capacity
        if (((w.c0 >> 3) != b0) | ((w.c1 >> 3) != b1) | ((w.c2 >> 3) != b2)) break;
      }
      if (bits & ~seen) atomicOr(mask + word, bits);
    }
  }
  // statistics: traversals (sum) and longest ray (max), warp-reduced
  uint32_t sum = count, mx = count;
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  if ((threadIdx.x & 31u) == 0 && sum) {
    atomicAdd(&cnt->traversals, (unsigned long long)sum);
    atomicMax(&cnt->longest_ray, mx);
  }
}

// K2, dense-window mode: no lookups at all.  The free mask word of every visited voxel is
// addressed directly from its key; the bits of one 32-voxel word are collected in a register
// and flushed with a single RED.OR when the ray moves on to another word, and the block's bit
// in the window's touched bitmap is set when the ray enters a block.
template <bool kFilter>
__global__ void __launch_bounds__(128)
k_walk_dense(DevParams P, ScanBuffers sb, MarkSet ms, Counters* cnt, float ox, float oy, float oz, int variant) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  RayWalk w;
  w.live = false;
  if (r < cnt->n_rays) {
    const uint32_t i = sb.rays[r];
    w.init(P, ox, oy, oz, sb.ray_end[3 * i], sb.ray_end[3 * i + 1], sb.ray_end[3 * i + 2]);
  }
  uint32_t count = 0;
  // word index (window block * 16 + word), the bits collected for it, and the word's value as
  // read when the ray entered it.  Near the sensor every ray sets the same voxels: a RED is only
  // issued for bits the preloaded value does not show yet (a stale value costs one redundant
  // RED, never a missed bit), which keeps tens of thousands of same-address atomics off the L2.
  uint32_t cur = 0xFFFFFFFFu, bits = 0, seen = 0;
  while (w.live) {
    ++count;
    if (!kFilter || free_filter_pass(P, ox, oy, oz, w.c0, w.c1, w.c2)) {
      const uint32_t idx = window_index(ms, w.c0, w.c1, w.c2);
      if (idx == kNotFound) {   // outside the window: the host re-runs the scan in table mode
        atomicExch(&cnt->overflow_window, 1u);
        break;
      }
      const uint32_t l = local_index(w.c0, w.c1, w.c2);
      const uint32_t wi = (idx << 4) | (l >> 5);
      if (wi != cur) {
        if ((bits & ~seen) && variant != 1) atomicOr(ms.free_mask + cur, bits);
        bits = 0;
        if ((wi >> 4) != (cur >> 4)) {
          uint32_t* tb = ms.touched_bits + (idx >> 5);
          if (!((*tb >> (idx & 31u)) & 1u)) atomicOr(tb, 1u << (idx & 31u));
        }
        cur = wi;
        seen = (variant == 0 || (variant == 3 && count < 48u)) ? ms.free_mask[wi] : 0u;
      }
      bits |= 1u << (l & 31u);
    }
    if (!w.advance()) break;
    if (count >= kKeyRayMax - 1u) break;   // KeyRay capacity
  }
  if ((bits & ~seen) && variant != 1) atomicOr(ms.free_mask + cur, bits);
  uint32_t sum = count, mx = count;
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  }
  if ((threadIdx.x & 31u) == 0 && sum) {
    atomicAdd(&cnt->traversals, (unsigned long long)sum);
    atomicMax(&cnt->longest_ray, mx);
  }
}

// K2a: one thread per cast ray.  computeRayKeys' preamble, then the state of the walk at every
// segment boundary by exact jump-ahead of the three tMax sequences (vmap_seq.cuh) -- no DDA
// stepping here.  Level j's checkpoint of the ray at ordered position r goes to slot
// seg_base[j] + r.
__global__ void __launch_bounds__(128)
k_checkpoints(DevParams P, ScanBuffers sb, Counters* cnt, float ox, float oy, float oz) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cnt->n_rays) return;
  const uint32_t i = sb.rays[r];
  ray_checkpoints(P, ox, oy, oz, sb.ray_end[3 * i], sb.ray_end[3 * i + 1], sb.ray_end[3 * i + 2], sb.len[i], i,
                  &sb.ray_rec[r], [&](uint32_t j, const Checkpoint& cp) {
                    const uint32_t slot = sb.seg_base[j] + r;
                    if (slot >= sb.ckpt_cap) {
                      atomicExch(&cnt->overflow_ckpt, 1u);
                      return false;
                    }
                    sb.ckpt[slot] = cp;
                    return true;
                  });
}

// K2a of the folded chain: CTA b takes the rays k_resolve_fold's CTA b left in sb.rays[256 b ...]
// (its cast rays, longest first), one thread per ray; the ray's position in the ordered list is
// bin_off[bin] + rank.
// (Three lanes per ray -- one per axis, checkpoints assembled with shuffles -- was built and measured:
// 42 us instead of 25, the three axes take different paths through seq_advance and serialise inside
// the warp.  Rays taken in point order instead of longest first: +14 us, a warp then waits for its
// longest ray.  profiles/README.md, round 2.)
__global__ void __launch_bounds__(256)
k_checkpoints_pts(DevParams P, ScanBuffers sb, Counters* cnt, float ox, float oy, float oz) {
  pdl_wait();
  const uint32_t i = sb.rays[blockIdx.x * 256u + threadIdx.x];
  if (i != kNoRay) {
    const uint32_t steps = sb.len[i];
    const uint32_t r = sb.bin_off[fine_bin(steps)] + sb.rank[i];
    ray_checkpoints(P, ox, oy, oz, sb.ray_end[3 * i], sb.ray_end[3 * i + 1], sb.ray_end[3 * i + 2], steps, i,
                    &sb.ray_rec[r], [&](uint32_t j, const Checkpoint& cp) {
                      const uint32_t slot = sb.seg_base[j] + r;
                      if (slot >= sb.ckpt_cap) {
                        atomicExch(&cnt->overflow_ckpt, 1u);
                        return false;
                      }
                      sb.ckpt[slot] = cp;
                      return true;
                    });
  }
  pdl_launch();
}

// K2b: one thread per (segment level, ray) work slot, level-major so that the lanes of a warp
// walk neighbouring rays at the same depth.  Segment j performs the DDA steps whose crossing
// parameter is in [j, j+1) * dtau and emits the voxels those steps enter (segment 0 also the
// origin voxel); only the last segment can see computeRayKeys' termination tests fire.  Marks
// go to the dense window exactly as in k_walk_dense.
template <bool kFilter>
__global__ void __launch_bounds__(256)
k_walk_segments(DevParams P, ScanBuffers sb, MarkSet ms, Counters* cnt, float ox, float oy, float oz,
                int variant) {
  // some checkpoints were never written (buffer too small): nothing may be walked; the host
  // enlarges the buffer and runs the scan again
  if (cnt->overflow_ckpt) return;
  __shared__ uint32_t base_s[kLenBins + 1];
  for (int k = threadIdx.x; k <= kLenBins; k += blockDim.x) base_s[k] = sb.seg_base[k];
  __syncthreads();
  const uint32_t total = min(base_s[kLenBins], sb.ckpt_cap);
  uint32_t emitted = 0, longest = 0;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x) {
    // level of the slot: largest j with base_s[j] <= slot
    uint32_t lo = 0, hi = kLenBins;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (base_s[mid] <= slot) lo = mid; else hi = mid;
    }
    const uint32_t j = lo, r = slot - base_s[lo];
    const Checkpoint cp = sb.ckpt[slot];
    if (cp.count == 0xFFFFFFFFu) continue;
    walk_slot<kFilter>(P, ms, cnt, cp, sb.ray_rec[r], j, ox, oy, oz, variant, &emitted, &longest);
  }
  __shared__ uint32_t s_stat[2];
  __syncthreads();                       // (base_s is dead by now)
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  __syncthreads();
  cta_add_u32(&s_stat[0], emitted);      // < 2^32 per CTA: <= 256 threads x a few segments x 100k keys
  cta_max_u32(&s_stat[1], longest);
  __syncthreads();
  if (threadIdx.x == 0 && s_stat[0]) {
    atomicAdd(&cnt->traversals, (unsigned long long)s_stat[0]);
    atomicMax(&cnt->longest_ray, s_stat[1]);
  }
}

// K2b with dilated window addressing (vmap_walk.cuh: walk_slot_dilated); same slots, same marks.
// Only launched for windows whose x / y extents are powers of two (VMAP_WALK_ADDR >= 1).
// order == 1 (VMAP_WALK_ORDER, experimental): the 256-slot chunks are visited in golden-ratio
// stride order instead of level-major order, so that the first-segment slots -- which all mark the
// same words around the sensor -- are spread over the kernel's run time instead of filling its
// first wave (marks do not depend on the order).
template <bool kFilter, int kMode, int kMinCtas, bool kWide>
__global__ void __launch_bounds__(256, kMinCtas)
k_walk_segments_dilated(DevParams P, ScanBuffers sb, MarkSet ms, Counters* cnt, float ox, float oy, float oz,
                        int variant, int order, DilatedWindow dw) {
  // (dw = dilated_window(ms), computed by the host: as kernel parameters the axis masks are
  // constant-bank operands of the step's LOP3s instead of values rebuilt at every step)
  if (cnt->overflow_ckpt) return;
  __shared__ uint32_t base_s[kLenBins + 1];
  for (int k = threadIdx.x; k <= kLenBins; k += blockDim.x) base_s[k] = sb.seg_base[k];
  __syncthreads();
  const uint32_t total = min(base_s[kLenBins], sb.ckpt_cap);
  const uint32_t near_limit = (variant == 0) ? 0xFFFFFFFFu : (variant == 2 ? 0u : (uint32_t)variant);
  const uint32_t n_chunks = (total + 255u) >> 8;
  uint32_t stride = 1;
  const bool counted = (order & 2) != 0;   // bit 1: non-last segments run a counted loop (VMAP_WALK_COUNTED, default)
  if ((order & 1) && n_chunks > 2u) {
    stride = (uint32_t)((double)n_chunks * 0.6180339887498949);
    if (stride < 1u) stride = 1u;
    for (;; stride++) {           // a stride coprime to n_chunks visits every chunk exactly once
      uint32_t a = stride, b = n_chunks;
      while (b) { const uint32_t t = a % b; a = b; b = t; }
      if (a == 1u) break;
    }
  }
  uint32_t emitted = 0, longest = 0;
  for (uint32_t q = blockIdx.x; q < n_chunks; q += gridDim.x) {
    const uint32_t chunk = (uint32_t)(((unsigned long long)q * stride) % n_chunks);
    const uint32_t slot = chunk * 256u + threadIdx.x;
    if (slot >= total) continue;
    uint32_t lo = 0, hi = kLenBins;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (base_s[mid] <= slot) lo = mid; else hi = mid;
    }
    const Checkpoint cp = sb.ckpt[slot];
    if (cp.count == 0xFFFFFFFFu) continue;
    // steps of a non-last segment = keys emitted up to the next level's checkpoint - up to this one
    // (level lo + 1 holds a slot for this ray exactly when this is not its last segment)
    const uint32_t r = slot - base_s[lo];
    uint32_t n_inner = 0xFFFFFFFFu;
    // (lo <= kLenBins - 2: a ray has at most kLenBins - 1 segments)
    const RayRec& rr = sb.ray_rec[r];
    if (counted) {
      if (r < base_s[lo + 2] - base_s[lo + 1]) n_inner = sb.ckpt[base_s[lo + 1] + r].count - cp.count;
      else if (rr.n_keys != 0xFFFFFFFFu) n_inner = rr.n_keys > cp.count ? rr.n_keys - cp.count : 0u;
    }
    walk_slot_dilated<kFilter, kMode, false, kWide>(P, ms, dw, cnt, cp, rr, lo, ox, oy, oz, near_limit, &emitted, &longest, NearGrid(),
                                      n_inner, (order & 4) != 0);
  }
  __shared__ uint32_t s_stat[2];
  __syncthreads();
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  __syncthreads();
  cta_add_u32(&s_stat[0], emitted);
  cta_max_u32(&s_stat[1], longest);
  __syncthreads();
  if (threadIdx.x == 0 && s_stat[0]) {
    atomicAdd(&cnt->traversals, (unsigned long long)s_stat[0]);
    atomicMax(&cnt->longest_ray, s_stat[1]);
  }
}

// K2b, dilated addressing + queued marks (vmap_walk.cuh: walk_slot_queued) -- the default walker.
// Same slots and marks as k_walk_segments_dilated; the lanes of a warp run one loop nest (every slot
// knows its step count up front) and drain their parked mask words together.
template <bool kFilter, int kMinCtas, bool kBranchFree>
__global__ void __launch_bounds__(256, kMinCtas)
k_walk_segments_queued(DevParams P, ScanBuffers sb, MarkSet ms, Counters* cnt, float ox, float oy, float oz, int variant,
                       DilatedWindow dw) {
  pdl_wait();
  if (cnt->overflow_ckpt) return;
  __shared__ uint32_t base_s[kLenBins + 1];
  __shared__ uint32_t queue_s[3 * kQueueSteps * 256];
  for (int k = threadIdx.x; k <= kLenBins; k += blockDim.x) base_s[k] = sb.seg_base[k];
  __syncthreads();
  const uint32_t total = min(base_s[kLenBins], sb.ckpt_cap);
  const uint32_t near_limit = (variant == 0) ? 0xFFFFFFFFu : (variant == 2 ? 0u : (uint32_t)variant);
  uint32_t* const mq = queue_s + threadIdx.x;
  uint32_t emitted = 0, longest = 0;
  // (Handing the slots out dynamically, 32 at a time to whichever warp asks next, measured 8 % slower
  // than this static deal of 256-slot chunks: profiles/r02_ab_walk6.json.)
  const uint32_t n_chunks = (total + 255u) >> 8;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const uint32_t slot = chunk * 256u + threadIdx.x;
    bool active = slot < total;
    uint32_t lo = 0, r = 0, n_steps = 0;
    Checkpoint cp;
    cp.count = 0xFFFFFFFFu;
    if (active) {
      uint32_t hi = kLenBins;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (base_s[mid] <= slot) lo = mid; else hi = mid;
      }
      r = slot - base_s[lo];
      cp = sb.ckpt[slot];
      active = cp.count != 0xFFFFFFFFu;
    }
    const RayRec& rr = sb.ray_rec[r];
    if (active) {
      // steps of a non-last segment = keys emitted up to the next level's checkpoint - up to this one
      // (level lo + 1 holds a slot for this ray exactly when this is not its last segment; lo <= kLenBins - 2)
      if (r < base_s[lo + 2] - base_s[lo + 1]) n_steps = sb.ckpt[base_s[lo + 1] + r].count - cp.count;
      else if (rr.n_keys != 0xFFFFFFFFu) n_steps = rr.n_keys > cp.count ? rr.n_keys - cp.count : 0u;
      else {
        // (a ray whose key count ray_checkpoints could not bound: walked the plain way, by this lane alone)
        walk_slot_dilated<kFilter, 0, false, true>(P, ms, dw, cnt, cp, rr, lo, ox, oy, oz, near_limit, &emitted, &longest);
        active = false;
      }
    }
    walk_slot_queued<kFilter, 256u, kBranchFree>(P, ms, dw, cnt, cp, rr, lo, ox, oy, oz, near_limit, &emitted, &longest, n_steps, active, mq);
  }
  pdl_launch();
  __shared__ uint32_t s_stat[2];
  __syncthreads();
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  __syncthreads();
  cta_add_u32(&s_stat[0], emitted);
  cta_max_u32(&s_stat[1], longest);
  __syncthreads();
  if (threadIdx.x == 0 && s_stat[0]) {
    atomicAdd(&cnt->traversals, (unsigned long long)s_stat[0]);
    atomicMax(&cnt->longest_ray, s_stat[1]);
  }
}

// K2b, dilated addressing + near-field combining (vmap_walk.cuh: NearGrid).  Chunks that consist
// of first-segment slots only collect the marks around the sensor in a 32 KiB shared bit grid and
// flush it once per chunk; every other chunk runs exactly like k_walk_segments_dilated.
__global__ void __launch_bounds__(256, 4)
k_walk_segments_near(DevParams P, ScanBuffers sb, MarkSet ms, Counters* cnt, float ox, float oy, float oz,
                     int variant, int order) {
  if (cnt->overflow_ckpt) return;
  __shared__ uint32_t base_s[kLenBins + 1];
  __shared__ uint32_t near_s[kNearWords];
  for (int k = threadIdx.x; k <= kLenBins; k += blockDim.x) base_s[k] = sb.seg_base[k];
  for (uint32_t k = threadIdx.x; k < kNearWords; k += blockDim.x) near_s[k] = 0u;
  __syncthreads();
  const uint32_t total = min(base_s[kLenBins], sb.ckpt_cap);
  const DilatedWindow dw = dilated_window(ms);
  const uint32_t near_limit = (variant == 0) ? 0xFFFFFFFFu : (variant == 2 ? 0u : (uint32_t)variant);
  NearGrid near;
  near.words = near_s;
  bool near_ok = false;
  {
    // every ray starts in the sensor's voxel (coordToKeyChecked(origin), as in RayWalk::init)
    const int s0 = scaled_coord(P.res_factor, (double)ox), s1 = scaled_coord(P.res_factor, (double)oy),
              s2 = scaled_coord(P.res_factor, (double)oz);
    if (key_in_range(s0) && key_in_range(s1) && key_in_range(s2))
      near_ok = near_grid_place(ms, (uint32_t)s0, (uint32_t)s1, (uint32_t)s2, &near);
  }
  const uint32_t n_chunks = (total + 255u) >> 8;
  uint32_t stride = 1;
  if ((order & 1) && n_chunks > 2u) {
    stride = (uint32_t)((double)n_chunks * 0.6180339887498949);
    if (stride < 1u) stride = 1u;
    for (;; stride++) {
      uint32_t a = stride, b = n_chunks;
      while (b) { const uint32_t t = a % b; a = b; b = t; }
      if (a == 1u) break;
    }
  }
  uint32_t emitted = 0, longest = 0;
  for (uint32_t q = blockIdx.x; q < n_chunks; q += gridDim.x) {
    const uint32_t chunk = (uint32_t)(((unsigned long long)q * stride) % n_chunks);
    const uint32_t slot = chunk * 256u + threadIdx.x;
    // (uniform over the CTA) the whole chunk lies in level 0
    const bool first_level = near_ok && (chunk * 256u + 256u <= base_s[1]);
    if (slot < total) {
      uint32_t lo = 0, hi = kLenBins;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (base_s[mid] <= slot) lo = mid; else hi = mid;
      }
      const Checkpoint cp = sb.ckpt[slot];
      if (cp.count != 0xFFFFFFFFu) {
        if (first_level)
          walk_slot_dilated<false, 0, true>(P, ms, dw, cnt, cp, sb.ray_rec[slot - base_s[lo]], lo, ox, oy, oz, near_limit,
                                            &emitted, &longest, near);
        else
          walk_slot_dilated<false, 0, false>(P, ms, dw, cnt, cp, sb.ray_rec[slot - base_s[lo]], lo, ox, oy, oz, near_limit,
                                             &emitted, &longest);
      }
    }
    if (first_level) {
      __syncthreads();
      for (uint32_t k = threadIdx.x; k < kNearWords; k += blockDim.x) near_flush_word(ms, near, k);
      __syncthreads();
    }
  }
  __shared__ uint32_t s_stat[2];
  __syncthreads();
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  __syncthreads();
  cta_add_u32(&s_stat[0], emitted);
  cta_max_u32(&s_stat[1], longest);
  __syncthreads();
  if (threadIdx.x == 0 && s_stat[0]) {
    atomicAdd(&cnt->traversals, (unsigned long long)s_stat[0]);
    atomicMax(&cnt->longest_ray, s_stat[1]);
  }
}

// Compact the window's touched bitmap into the block list (cnt->n_touched entries).
__global__ void __launch_bounds__(256) k_list_touched(MarkSet ms, Counters* cnt, uint32_t n_words) {
  pdl_wait();
  pdl_launch();
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t wbase = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; wbase < n_words;
       wbase += gridDim.x * blockDim.x) {
    const uint32_t wi = wbase + lane;
    const uint32_t v = wi < n_words ? ms.touched_bits[wi] : 0u;
    const uint32_t c = __popc(v);
    uint32_t incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (!total) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&cnt->n_touched, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    uint32_t pos = base + incl - c;
    uint32_t rest = v;
    while (rest) {
      const uint32_t b = __ffs(rest) - 1;
      rest &= rest - 1;
      if (pos < ms.list_cap) ms.list[pos] = wi * 32u + b;
      else atomicExch(&cnt->overflow_keys, 1u);   // list too small: the host grows it
      ++pos;
    }
  }
}

// Admission of a scan's update, decided ONCE by one thread before anything of the scan touches the
// map (dense mode: before k_locate_blocks inserts entries).  Refused -- cnt->stall != 0, marks left
// intact -- when an earlier scan is stalled (the order of the clamped updates must hold), when the
// marks are incomplete, or when the listed blocks might push the table over load 1/2.  `poison`
// (pipelined insertion): a refusal also stalls every later scan until the host has replayed it.
__global__ void k_admit(MarkSet ms, MapView m, Counters* cnt, int poison) {
  pdl_wait();
  pdl_launch();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t why = 0;
  if (m.persist->stalled) why |= kStallEarlier;
  if (cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys) why |= kStallMarks;
  if (((unsigned long long)m.persist->n_entries + cnt->n_touched) * 2ull > (unsigned long long)m.cap_mask + 1ull) why |= kStallTable;
  cnt->stall = why;
  if (why && poison) m.persist->stalled = 1u;
}

// Table entry of every listed window block (inserted when new).  The last CTA to finish decides,
// once, whether the pool has a slot for every listed block that lacks one (k_update only reads
// the verdict).
__global__ void __launch_bounds__(256) k_locate_blocks(MarkSet ms, MapView m, Counters* cnt, int poison) {
  pdl_wait();
  pdl_launch();
  if (cnt->stall) return;
  const uint32_t n = cnt->n_touched;
  uint32_t need = 0;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const uint32_t h = find_or_insert(m, cnt, window_block_key(ms, ms.list[t]));
    ms.locate[t] = h;
    if (h != kNotFound && m.slot[h] == kUnassigned) ++need;
  }
  __shared__ uint32_t s_need;
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_need = 0;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) need += __shfl_xor_sync(0xFFFFFFFFu, need, o);
  if ((threadIdx.x & 31u) == 0 && need) atomicAdd(&s_need, need);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_need) atomicAdd(&cnt->need_slots, s_need);
    __threadfence();
    s_last = atomicAdd(&cnt->done_ctas, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    const uint32_t total = *(volatile uint32_t*)&cnt->need_slots;
    if (*(volatile uint32_t*)&cnt->overflow_table) {
      cnt->stall |= kStallTable;                      // (probe limit hit: cannot happen below load 1/2)
      if (poison) m.persist->stalled = 1u;
    } else if ((unsigned long long)m.persist->n_slots + total > (unsigned long long)m.pool_cap) {
      cnt->stall |= kStallPool;
      cnt->overflow_pool = 1u;
      if (poison) m.persist->stalled = 1u;
    }
  }
}

// Table mode (marks in the table's mask columns, entries already inserted by the walk): the pool
// verdict by one thread -- at most one new slot per touched block, and (single GPU, where every
// entry ends up owning a slot) never more than there are entries.
__global__ void k_admit_pool(MapView m, Counters* cnt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (cnt->overflow_table || cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys) return;
  const uint32_t by_touch = m.persist->n_slots + cnt->n_touched;
  const uint32_t slots_bound = m.sharded ? by_touch : min(by_touch, m.persist->n_entries);
  if (slots_bound > m.pool_cap) {
    cnt->overflow_pool = 1u;
    cnt->stall |= kStallPool;
  }
}

// updateOccupancy on explicit key lists (vmap_apply_keys): mark phase.
__global__ void __launch_bounds__(256)
k_mark_keys(MapView m, Counters* cnt, const uint16_t* __restrict__ k3, uint32_t n, int occupied) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mark_voxel(m, cnt, occupied ? m.occ_mask : m.free_mask, k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]);
}

// Parity hook: expand this scan's masks into explicit key lists (before k_update clears them).
__global__ void __launch_bounds__(256)
k_capture_keys(MapView m, MarkSet ms, Counters* cnt, uint16_t* __restrict__ free_k3, uint32_t* n_free,
               uint32_t free_cap, uint16_t* __restrict__ occ_k3, uint32_t* n_occ, uint32_t occ_cap) {
  if (cnt->stall || cnt->overflow_table || cnt->overflow_pool || cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys) return;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < cnt->n_touched; t += warps) {
    uint32_t e, h;
    const unsigned long long b = listed_block(ms, m, t, &e, &h);
    const uint32_t bx = (uint32_t)(b & 0x1FFFu) << 3, by = (uint32_t)((b >> 13) & 0x1FFFu) << 3,
                   bz = (uint32_t)((b >> 26) & 0x1FFFu) << 3;
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t o = ms.occ_mask[(size_t)e * kMaskWords + j];
      const uint32_t f = ms.free_mask[(size_t)e * kMaskWords + j] & ~o;
      const uint32_t l = (uint32_t)j * 32u + lane;
      const uint32_t kx = bx + (l & 7u), ky = by + ((l >> 3) & 7u), kz = bz + (l >> 6);
      if ((o >> lane) & 1u) {
        const uint32_t p = atomicAdd(n_occ, 1u);
        if (p < occ_cap) { occ_k3[3 * p] = kx; occ_k3[3 * p + 1] = ky; occ_k3[3 * p + 2] = kz; }
      } else if ((f >> lane) & 1u) {
        const uint32_t p = atomicAdd(n_free, 1u);
        if (p < free_cap) { free_k3[3 * p] = kx; free_k3[3 * p + 1] = ky; free_k3[3 * p + 2] = kz; }
      }
    }
  }
}

// K4: one warp per touched block.  free \ occupied, one clamped update per marked voxel
// (occupied: +hit, free: +miss), masks cleared.  Every voxel has exactly one writer.
template <int kVariant>
__global__ void __launch_bounds__(256)
k_update(DevParams P, MapView m, MarkSet ms, Counters* cnt) {
  pdl_wait();
  // Abort the whole scan's update (uniformly) when a map structure must grow first: the verdict was
  // reached once, before this launch (k_admit / k_locate_blocks / k_admit_pool), every thread reads
  // the same flags.
  if (cnt->stall || cnt->overflow_table || cnt->overflow_pool || cnt->overflow_window || cnt->overflow_ckpt ||
      cnt->overflow_keys)
    return;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_touched = cnt->n_touched;
  uint32_t n_free = 0, n_occ = 0, lane_free = 0, lane_occ = 0;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_touched; t += warps) {
    uint32_t e, h;
    listed_block(ms, m, t, &e, &h);
    uint32_t f = 0, o = 0;
    if (lane < kMaskWords) {
      uint32_t* pf = ms.free_mask + (size_t)e * kMaskWords + lane;
      uint32_t* po = ms.occ_mask + (size_t)e * kMaskWords + lane;
      f = *pf; o = *po;
      if (f) *pf = 0u;
      if (o) *po = 0u;
      f &= ~o;  // occupied wins (octomap_world.cc:252-254)
    }
    if (!__any_sync(0xFFFFFFFFu, (f | o) != 0u)) continue;   // e.g. every voxel filtered out
    uint32_t slot = 0;
    if (lane == 0) {
      slot = m.slot[h];
      if (slot == kUnassigned) {
        slot = atomicAdd(&m.persist->n_slots, 1u);
        m.slot[h] = slot;
      }
    }
    slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
    uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    if (kVariant == 1) {
      // every row's load is issued before the first update: one dependent round trip per block
      // instead of sixteen (the kernel is latency-bound otherwise)
      uint32_t val[kMaskWords];
#pragma unroll
      for (int j = 0; j < kMaskWords; j++) {
        const uint32_t any = __shfl_sync(0xFFFFFFFFu, f | o, j);
        val[j] = ((any >> lane) & 1u) ? __ldcg(vox + j * 32 + lane) : 0u;
      }
#pragma unroll
      for (int j = 0; j < kMaskWords; j++) {
        const uint32_t fw = __shfl_sync(0xFFFFFFFFu, f, j);
        const uint32_t ow = __shfl_sync(0xFFFFFFFFu, o, j);
        const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
        if (is_occ || is_free) vox[j * 32 + lane] = update_leaf(val[j], is_occ ? P.hit : P.miss, P.cmin, P.cmax);
        if (lane == 0) { n_free += __popc(fw); n_occ += __popc(ow); }
      }
    } else if (m.chg0 != nullptr) {
      // change detection on (OccupancyOcTreeBase::updateNodeRecurs with use_change_detection):
      // a created leaf is recorded as `true`; an occupancy flip inserts `false`, or erases an
      // earlier `false`; a `true` entry is never changed.  Early-aborted updates record nothing.
      uint32_t* c0 = m.chg0 + (size_t)slot * kMaskWords;
      uint32_t* c1 = m.chg1 + (size_t)slot * kMaskWords;
      for (int j = 0; j < kMaskWords; j++) {
        const uint32_t fw = __shfl_sync(0xFFFFFFFFu, f, j);
        const uint32_t ow = __shfl_sync(0xFFFFFFFFu, o, j);
        if ((fw | ow) == 0u) continue;
        const uint32_t o0 = c0[j], o1 = c1[j];
        uint32_t code = ((o0 >> lane) & 1u) | (((o1 >> lane) & 1u) << 1);
        const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
        if (is_occ || is_free) {
          uint32_t* p = vox + j * 32 + lane;
          const uint32_t before = *p;
          const uint32_t after = update_leaf(before, is_occ ? P.hit : P.miss, P.cmin, P.cmax);
          *p = after;
          if (before == kUnknownBits) {
            code = 1u;
          } else if ((__uint_as_float(before) >= P.occ_thres) != (__uint_as_float(after) >= P.occ_thres)) {
            if (code == 0u) code = 2u;
            else if (code == 2u) code = 0u;
          }
        }
        const uint32_t n0 = __ballot_sync(0xFFFFFFFFu, code & 1u), n1 = __ballot_sync(0xFFFFFFFFu, code >> 1);
        if (lane == 0) {
          if (n0 != o0) c0[j] = n0;
          if (n1 != o1) c1[j] = n1;
          n_free += __popc(fw); n_occ += __popc(ow);
        }
      }
    } else {
      // (statistics: every lane counts its own mask word once per block instead of lane 0 counting
      // every row -- the row loop is what the kernel's issue slots go to)
      lane_free += __popc(f);
      lane_occ += __popc(o);
#pragma unroll
      for (int j = 0; j < kMaskWords; j++) {
        const uint32_t fw = __shfl_sync(0xFFFFFFFFu, f, j);
        const uint32_t ow = __shfl_sync(0xFFFFFFFFu, o, j);
        if ((fw | ow) == 0u) continue;
        const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
        if (is_occ || is_free) {
          uint32_t* p = vox + j * 32 + lane;
          *p = update_leaf(*p, is_occ ? P.hit : P.miss, P.cmin, P.cmax);
        }
      }
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    lane_free += __shfl_xor_sync(0xFFFFFFFFu, lane_free, o2);
    lane_occ += __shfl_xor_sync(0xFFFFFFFFu, lane_occ, o2);
  }
  n_free += lane_free;
  n_occ += lane_occ;
  __shared__ uint32_t s_stat[2];
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  __syncthreads();
  if (lane == 0) {
    if (n_free) atomicAdd(&s_stat[0], n_free);
    if (n_occ) atomicAdd(&s_stat[1], n_occ);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_stat[0]) atomicAdd(&cnt->unique_free, s_stat[0]);
    if (s_stat[1]) atomicAdd(&cnt->unique_occupied, s_stat[1]);
  }
  // The scan is applied: the window's touched bitmap is free for the next scan (nobody in this
  // kernel reads it -- the blocks come from the list).  A refused update returns before this point,
  // so its bitmap, list and marks all survive for the replay.
  if (ms.dense) {
    const uint32_t n_words = (ms.nx * ms.ny * ms.nz + 31u) >> 5;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
      if (ms.touched_bits[i]) ms.touched_bits[i] = 0u;
  }
}

// K4 with the memory accesses of a block in flight together (the default without change detection).
// k_update<0> walks a block row by row -- load, update, store -- so a warp has ONE 128-byte load in
// flight at a time and sixteen dependent round trips per block (85 % warps active, 20 warps stalled on
// long scoreboard per issued instruction, ~55 % of the HBM time for the traffic it moves).  Here a warp
//   * keeps the list entry of the block after next and the masks + pool slot of the next block in
//     registers while it works on the current one (two-deep software pipeline over its blocks), and
//   * issues the loads of every marked row of the current block before the first update.
// Launched with as many CTAs as are resident at once (the register count is what it is).
__global__ void __launch_bounds__(256)
k_update_mlp(DevParams P, MapView m, MarkSet ms, Counters* cnt) {
  pdl_wait();
  if (cnt->stall || cnt->overflow_table || cnt->overflow_pool || cnt->overflow_window || cnt->overflow_ckpt ||
      cnt->overflow_keys)
    return;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_touched = cnt->n_touched;
  uint32_t lane_free = 0, lane_occ = 0;
  uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // stage A: list entry; stage B: masks (cleared as they are read) + pool slot
  uint32_t eA = 0, hA = 0, eB = 0, hB = 0, fB = 0, oB = 0, slotB = 0;
  auto stage_a = [&](uint32_t tt) {
    if (tt < n_touched) listed_block(ms, m, tt, &eA, &hA);
  };
  auto stage_b = [&](uint32_t tt) {       // from the entry in (eA, hA)
    eB = eA; hB = hA; fB = 0; oB = 0; slotB = 0;
    if (tt >= n_touched) return;
    if (lane < kMaskWords) {
      uint32_t* pf = ms.free_mask + (size_t)eB * kMaskWords + lane;
      uint32_t* po = ms.occ_mask + (size_t)eB * kMaskWords + lane;
      fB = *pf; oB = *po;
      if (fB) *pf = 0u;
      if (oB) *po = 0u;
      fB &= ~oB;  // occupied wins (octomap_world.cc:252-254)
    }
    if (lane == 0) slotB = m.slot[hB];
  };
  stage_a(t);
  stage_b(t);
  stage_a(t + warps);
  for (; t < n_touched; t += warps) {
    const uint32_t f = fB, o = oB, h = hB;
    uint32_t slot = slotB;
    stage_b(t + warps);
    stage_a(t + 2u * warps);
    const uint32_t rows = __ballot_sync(0xFFFFFFFFu, (f | o) != 0u);   // bit j: row j has marks (lanes >= 16 hold 0)
    if (!rows) continue;   // e.g. every voxel filtered out
    if (lane == 0 && slot == kUnassigned) {
      slot = atomicAdd(&m.persist->n_slots, 1u);
      m.slot[h] = slot;
    }
    slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
    uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    uint32_t val[kMaskWords];
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) val[j] = ((rows >> j) & 1u) ? __ldcg(vox + j * 32 + lane) : 0u;
    lane_free += __popc(f);
    lane_occ += __popc(o);
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) {
      if (!((rows >> j) & 1u)) continue;
      const uint32_t fw = __shfl_sync(0xFFFFFFFFu, f, j);
      const uint32_t ow = __shfl_sync(0xFFFFFFFFu, o, j);
      const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
      if (is_occ || is_free) vox[j * 32 + lane] = update_leaf(val[j], is_occ ? P.hit : P.miss, P.cmin, P.cmax);
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    lane_free += __shfl_xor_sync(0xFFFFFFFFu, lane_free, o2);
    lane_occ += __shfl_xor_sync(0xFFFFFFFFu, lane_occ, o2);
  }
  __shared__ uint32_t s_stat[2];
  if (threadIdx.x < 2) s_stat[threadIdx.x] = 0;
  __syncthreads();
  if (lane == 0) {
    if (lane_free) atomicAdd(&s_stat[0], lane_free);
    if (lane_occ) atomicAdd(&s_stat[1], lane_occ);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_stat[0]) atomicAdd(&cnt->unique_free, s_stat[0]);
    if (s_stat[1]) atomicAdd(&cnt->unique_occupied, s_stat[1]);
  }
  if (ms.dense) {   // (as in k_update: the window's touched bitmap is free for the next scan)
    const uint32_t n_words = (ms.nx * ms.ny * ms.nz + 31u) >> 5;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
      if (ms.touched_bits[i]) ms.touched_bits[i] = 0u;
  }
}

// Drop this scan's marks (used when a scan has to be re-run after a grow).
__global__ void __launch_bounds__(256) k_clear_masks(MapView m, MarkSet ms, Counters* cnt) {
  const uint32_t total = cnt->n_touched * kMaskWords;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t e = ms.list[i / kMaskWords];
    ms.free_mask[(size_t)e * kMaskWords + (i % kMaskWords)] = 0u;
    ms.occ_mask[(size_t)e * kMaskWords + (i % kMaskWords)] = 0u;
  }
}

// Dense window: drop every mark by walking the touched bitmap itself (the list may be
// incomplete when this runs), then the bitmap.
__global__ void __launch_bounds__(256) k_clear_window(MarkSet ms, uint32_t n_words) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < n_words; wi += warps) {
    uint32_t v = ms.touched_bits[wi];
    while (v) {
      const uint32_t b = __ffs(v) - 1;
      v &= v - 1;
      const size_t e = (size_t)wi * 32u + b;
      if (lane < kMaskWords) ms.free_mask[e * kMaskWords + lane] = 0u;
      else ms.occ_mask[e * kMaskWords + lane - kMaskWords] = 0u;
    }
    __syncwarp();
    if (lane == 0) ms.touched_bits[wi] = 0u;
  }
}

// Rehash into a larger table (epochs reset to 0, slots carried over).
__global__ void __launch_bounds__(256)
k_rehash(const unsigned long long* __restrict__ old_keys, const uint32_t* __restrict__ old_slot,
         uint32_t old_cap, unsigned long long* new_keys, uint32_t* new_slot, uint32_t new_mask) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  const unsigned long long w = old_keys[i];
  if (w == kEmpty) return;
  const unsigned long long b = w >> kEpochBits;
  uint32_t h = hash_block(b) & new_mask;
  for (;;) {
    if (atomicCAS(new_keys + h, kEmpty, b << kEpochBits) == kEmpty) {
      new_slot[h] = old_slot[i];
      return;
    }
    h = (h + 1u) & new_mask;
  }
}

__global__ void __launch_bounds__(256) k_reset_epochs(unsigned long long* keys, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap && keys[i] != kEmpty) keys[i] &= ~(unsigned long long)kEpochMask;
}

// Give every touched entry a pool slot (import path; k_update does this inline).
__global__ void __launch_bounds__(256) k_assign_slots(MapView m, Counters* cnt) {
  if (cnt->overflow_table) return;
  if (m.persist->n_entries > m.pool_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->overflow_pool = 1u;
    return;
  }
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt->n_touched) return;
  const uint32_t h = m.touched[t];
  if (m.slot[h] == kUnassigned) m.slot[h] = atomicAdd(&m.persist->n_slots, 1u);
}

__global__ void __launch_bounds__(256)
k_touch_keys(MapView m, Counters* cnt, const uint16_t* __restrict__ k3, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  find_or_insert_touch(m, cnt, block_key(k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]));
}

__global__ void __launch_bounds__(256)
k_write_leaves(MapView m, Counters* cnt, const uint16_t* __restrict__ k3,
               const float* __restrict__ val, uint32_t n) {
  if (cnt->overflow_table || cnt->overflow_pool) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t kx = k3[3 * i], ky = k3[3 * i + 1], kz = k3[3 * i + 2];
  const uint32_t h = find_block(m, block_key(kx, ky, kz));
  if (h == kNotFound) return;
  m.pool[(size_t)m.slot[h] * kBlockVoxels + local_index(kx, ky, kz)] = val[i];
}

// Export: count, then write, the known voxels of every resident block (one warp per entry).
__global__ void __launch_bounds__(256)
k_export(MapView m, Counters* cnt, uint16_t* __restrict__ k3, float* __restrict__ val, uint32_t cap) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h <= m.cap_mask; h += warps) {
    const unsigned long long w = m.keys[h];
    if (w == kEmpty) continue;
    const uint32_t slot = m.slot[h];
    if (slot == kUnassigned) continue;
    const unsigned long long b = w >> kEpochBits;
    const uint32_t bx = (uint32_t)(b & 0x1FFFu) << 3, by = (uint32_t)((b >> 13) & 0x1FFFu) << 3,
                   bz = (uint32_t)((b >> 26) & 0x1FFFu) << 3;
    const uint32_t* vox = reinterpret_cast<const uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t bits = vox[j * 32 + lane];
      const bool known = bits != kUnknownBits;
      const unsigned kb = __ballot_sync(0xFFFFFFFFu, known);
      if (!kb) continue;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&cnt->n_leaves, (uint32_t)__popc(kb));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (known && k3) {
        const uint32_t p = base + __popc(kb & ((1u << lane) - 1u));
        if (p < cap) {
          const uint32_t l = (uint32_t)j * 32u + lane;
          k3[3 * p] = (uint16_t)(bx + (l & 7u));
          k3[3 * p + 1] = (uint16_t)(by + ((l >> 3) & 7u));
          k3[3 * p + 2] = (uint16_t)(bz + (l >> 6));
          val[p] = __uint_as_float(bits);
        }
      }
    }
  }
}

// Order-independent digest of the map: the 64-bit sum over all known voxels of
// mix(mix(packed 48-bit key) ^ value bits).  Two maps hold the same leaves with the same float bit
// patterns iff count and digest agree (up to a 2^-64 collision); used to compare a sharded build
// with a single-device build without moving the leaves (bench.py, tests).
__device__ __forceinline__ unsigned long long digest_mix(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__global__ void __launch_bounds__(256)
k_digest(MapView m, unsigned long long* __restrict__ out /* [0] = count, [1] = digest */) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long sum = 0, cnt = 0;
  for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h <= m.cap_mask; h += warps) {
    const unsigned long long w = m.keys[h];
    if (w == kEmpty) continue;
    const uint32_t slot = m.slot[h];
    if (slot == kUnassigned) continue;
    const unsigned long long b = w >> kEpochBits;
    const uint32_t bx = (uint32_t)(b & 0x1FFFu) << 3, by = (uint32_t)((b >> 13) & 0x1FFFu) << 3,
                   bz = (uint32_t)((b >> 26) & 0x1FFFu) << 3;
    const uint32_t* vox = reinterpret_cast<const uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t bits = vox[j * 32 + lane];
      if (bits == kUnknownBits) continue;
      const uint32_t l = (uint32_t)j * 32u + lane;
      const unsigned long long key = (unsigned long long)(bx + (l & 7u)) | ((unsigned long long)(by + ((l >> 3) & 7u)) << 16) |
                                     ((unsigned long long)(bz + (l >> 6)) << 32);
      sum += digest_mix(digest_mix(key) ^ (unsigned long long)bits);
      ++cnt;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
  }
  if (lane == 0 && cnt) {
    atomicAdd(out, cnt);
    atomicAdd(out + 1, sum);
  }
}

// getChangedPoints (octomap_world.cc:1413-1440): every key of octomap's changed_keys with the
// CURRENT occupancy of its node; `clear` = resetChangeDetection().  Pass k3 == NULL to count.
__global__ void __launch_bounds__(256)
k_changed(DevParams P, MapView m, Counters* cnt, uint16_t* __restrict__ k3, uint8_t* __restrict__ state,
          uint32_t cap, int clear) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h <= m.cap_mask; h += warps) {
    const unsigned long long w = m.keys[h];
    if (w == kEmpty) continue;
    const uint32_t slot = m.slot[h];
    if (slot == kUnassigned) continue;
    const unsigned long long b = w >> kEpochBits;
    const uint32_t bx = (uint32_t)(b & 0x1FFFu) << 3, by = (uint32_t)((b >> 13) & 0x1FFFu) << 3,
                   bz = (uint32_t)((b >> 26) & 0x1FFFu) << 3;
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t any = m.chg0[(size_t)slot * kMaskWords + j] | m.chg1[(size_t)slot * kMaskWords + j];
      if (!any) continue;
      const bool mine = (any >> lane) & 1u;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&cnt->n_leaves, (uint32_t)__popc(any));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (mine && k3) {
        const uint32_t p = base + __popc(any & ((1u << lane) - 1u));
        if (p < cap) {
          const uint32_t l = (uint32_t)j * 32u + lane;
          k3[3 * p] = (uint16_t)(bx + (l & 7u));
          k3[3 * p + 1] = (uint16_t)(by + ((l >> 3) & 7u));
          k3[3 * p + 2] = (uint16_t)(bz + (l >> 6));
          const uint32_t bits = reinterpret_cast<const uint32_t*>(m.pool)[(size_t)slot * kBlockVoxels + l];
          state[p] = (bits != kUnknownBits && __uint_as_float(bits) >= P.occ_thres) ? 1 : 0;
        }
      }
      __syncwarp();
      if (clear && lane == 0) {
        m.chg0[(size_t)slot * kMaskWords + j] = 0u;
        m.chg1[(size_t)slot * kMaskWords + j] = 0u;
      }
    }
  }
}

// OccupancyOcTreeBase::toMaxLikelihood on the leaves: occupied -> clamping max, else clamping min.
__global__ void __launch_bounds__(256) k_to_max_likelihood(DevParams P, MapView m) {
  const size_t total = (size_t)m.persist->n_slots * kBlockVoxels;
  uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t bits = vox[i];
    if (bits == kUnknownBits) continue;
    vox[i] = __float_as_uint(__uint_as_float(bits) >= P.occ_thres ? P.cmax : P.cmin);
  }
}

// ---------------------------------------------------------------------------------------------
// K5 queries
// ---------------------------------------------------------------------------------------------
// getCellStatusPoint: octree_->search(x,y,z) quantises the DOUBLE coordinates.
//   true_status == 0 : getCellStatusPoint      (octomap_world.cc:346-360; unknown -> occupied when configured)
//   true_status == 1 : getCellTrueStatusPoint  (:363-373; unknown stays unknown)
//   logodds != NULL  : the node's log-odds (NaN when there is no node) for getCellProbabilityPoint (:375-393)
__global__ void __launch_bounds__(256)
k_query_points(DevParams P, MapView m, const double* __restrict__ xyz, uint32_t n,
               uint8_t* __restrict__ status, int true_status, float* __restrict__ logodds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t bits;
  const uint8_t st = point_status(P, m, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], true_status, &bits);
  status[i] = st;
  if (logodds) logodds[i] = __uint_as_float(bits == kUnknownBits ? 0x7FC00000u : bits);
}

__global__ void __launch_bounds__(128)
k_query_lines(DevParams P, MapView m, const double* __restrict__ ab, uint32_t n,
              uint8_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // pointEigenToOctomap narrows to float
  status[i] = line_status(P, m, (float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2],
                          (float)ab[6 * i + 3], (float)ab[6 * i + 4], (float)ab[6 * i + 5]);
}

// Measurement twin of k_query_lines: same walk, and the number of voxels probed (search() calls of
// the reference's loop, octomap_world.cc:405-415) summed into *probes -- the `probes` of SURVEY 8d's
// B_alg = 49 + 12 * probes per segment.
__global__ void __launch_bounds__(128)
k_query_lines_probes(DevParams P, MapView m, const double* __restrict__ ab, uint32_t n,
                     uint8_t* __restrict__ status, unsigned long long* __restrict__ probes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mine = 0;
  if (i < n) {
    uint8_t st = 0;
    unsigned long long cached_b = kEmpty;
    uint32_t cached_h = kNotFound;
    walk_ray(P, (float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2], (float)ab[6 * i + 3], (float)ab[6 * i + 4],
             (float)ab[6 * i + 5], [&](uint32_t kx, uint32_t ky, uint32_t kz) {
               const unsigned long long b = block_key(kx, ky, kz);
               if (b != cached_b) { cached_b = b; cached_h = find_block(m, b); }
               ++mine;
               const uint8_t s = voxel_status(P, m, cached_h, kx, ky, kz);
               if (s != 0) { st = s; return false; }
               return true;
             });
    status[i] = st;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(probes, (unsigned long long)mine);
}

// getVisibility (octomap_world.cc:420-449): occupied (or, when asked, unknown) voxels between the
// view point and the voxel under test, the voxel under test itself excepted.  Unknown is reported
// as unknown here whatever treat_unknown_as_occupied says, like the reference.
__global__ void __launch_bounds__(128)
k_query_visibility(DevParams P, MapView m, const double* __restrict__ ab, uint32_t n, int stop_at_unknown,
                   uint8_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  status[i] = visibility_status(P, m, (float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2], (float)ab[6 * i + 3],
                                (float)ab[6 * i + 4], (float)ab[6 * i + 5], stop_at_unknown);
}

// getLineStatusBoundingBox (octomap_world.cc:451-493): the reference shifts the segment by every
// offset of a grid spanning the box, in a fixed order, and returns the first non-free result.
// One thread per (segment, offset); the earliest offset with a non-free line wins.
__global__ void __launch_bounds__(128)
k_query_lines_bbox(DevParams P, MapView m, const double* __restrict__ ab, uint32_t n,
                   const double* __restrict__ offsets, uint32_t n_off, uint32_t* __restrict__ packed) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned long long)n * n_off) return;
  const uint32_t i = (uint32_t)(t / n_off), j = (uint32_t)(t % n_off);
  if (packed[i] < (j << 2)) return;     // an earlier offset already decided this segment
  const double ox = offsets[3 * j], oy = offsets[3 * j + 1], oz = offsets[3 * j + 2];
  const uint8_t s = line_status(P, m, (float)__dadd_rn(ab[6 * i], ox), (float)__dadd_rn(ab[6 * i + 1], oy),
                                (float)__dadd_rn(ab[6 * i + 2], oz), (float)__dadd_rn(ab[6 * i + 3], ox),
                                (float)__dadd_rn(ab[6 * i + 4], oy), (float)__dadd_rn(ab[6 * i + 5], oz));
  if (s != 0) atomicMin(packed + i, (j << 2) | s);
}

// getCellStatusBoundingBox (octomap_world.cc:274-344), batched: one warp per box, the lanes share
// the box's voxels (vmap_query.cuh).
__global__ void __launch_bounds__(256)
k_query_bbox(DevParams P, MapView m, const double* __restrict__ xyz, uint32_t n, BoxQuery q,
             uint8_t* __restrict__ status) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const BoxScan r = bbox_scan(P, m, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q, lane, 32u);
    const bool occ = __any_sync(0xFFFFFFFFu, r.occupied), unk = __any_sync(0xFFFFFFFFu, r.unknown);
    if (lane == 0) status[i] = bbox_combine(P, r.decided, occ, unk);
  }
}

// (first-hit index << 2 | status) -> CellStatus after the MIN-reduction over shards
__global__ void __launch_bounds__(256)
k_unpack_line_status(const uint32_t* __restrict__ packed, uint32_t n, uint8_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = packed[i];
  status[i] = (v == 0xFFFFFFFFu) ? (uint8_t)0 : (uint8_t)(v & 3u);
}

}  // namespace vmapdev
