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

namespace vmapdev {

constexpr uint32_t kFlagValid = 1u;   // point takes part in the scan
constexpr uint32_t kFlagRange = 2u;   // castRay's in-range branch
constexpr uint32_t kFlagBound = 4u;   // coordToKeyChecked(point) succeeds

struct ScanBuffers {
  float* world;             // [n*3] world-frame points (what the reference leaves in the cloud)
  float* ray_end;           // [n*3] ray end: the point, or the clipped end when out of range
  unsigned long long* key;  // [n]   unchecked endpoint key k0 | k1<<16 | k2<<32
  uint8_t* flags;           // [n]
  uint32_t* rays;           // [n]   indices of cast rays
  unsigned long long* first_keys;  // [first_cap] per-scan table: endpoint key -> first index
  uint32_t* first_idx;             // [first_cap]
  uint32_t first_mask;
};

struct Transform {
  double m[12];   // row-major 3x4 of getTransformationMatrix()
  double q[4];    // w x y z
  double t[3];
  float origin[3];
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
                                               float px, float py, float pz) {
  sb.world[3 * i + 0] = px;
  sb.world[3 * i + 1] = py;
  sb.world[3 * i + 2] = pz;
  const int s0 = scaled_coord(P.res_factor, (double)px);
  const int s1 = scaled_coord(P.res_factor, (double)py);
  const int s2 = scaled_coord(P.res_factor, (double)pz);
  const unsigned long long key = (unsigned long long)((uint32_t)s0 & 0xFFFFu) |
                                 ((unsigned long long)((uint32_t)s1 & 0xFFFFu) << 16) |
                                 ((unsigned long long)((uint32_t)s2 & 0xFFFFu) << 32);
  const bool bound = key_in_range(s0) && key_in_range(s1) && key_in_range(s2);
  // castRay: (point - sensor_origin).norm() <= sensor_max_range, norm() in double on a float sum
  float dx = __fsub_rn(px, origin[0]), dy = __fsub_rn(py, origin[1]), dz = __fsub_rn(pz, origin[2]);
  const double len = __dsqrt_rn((double)norm_sq_f(dx, dy, dz));
  const bool in_range = (P.max_range < 0.0) || (len <= P.max_range);
  float ex = px, ey = py, ez = pz;
  if (!in_range) {
    // new_end = origin + (point - origin).normalized() * sensor_max_range   (all float)
    if (len > 0) {
      const float fl = (float)len;
      dx = __fdiv_rn(dx, fl); dy = __fdiv_rn(dy, fl); dz = __fdiv_rn(dz, fl);
    }
    const float mr = (float)P.max_range;
    ex = __fadd_rn(origin[0], __fmul_rn(dx, mr));
    ey = __fadd_rn(origin[1], __fmul_rn(dy, mr));
    ez = __fadd_rn(origin[2], __fmul_rn(dz, mr));
  }
  sb.ray_end[3 * i + 0] = ex;
  sb.ray_end[3 * i + 1] = ey;
  sb.ray_end[3 * i + 2] = ez;
  sb.key[i] = key;
  sb.flags[i] = (uint8_t)(kFlagValid | (in_range ? kFlagRange : 0u) | (bound ? kFlagBound : 0u));
  if (in_range && bound) first_insert(sb, key, i);
  // statistics (warp-aggregated)
  const unsigned act = __activemask();
  const unsigned in_b = __ballot_sync(act, in_range);
  if ((threadIdx.x & 31u) == (uint32_t)(__ffs(act) - 1)) {
    atomicAdd(&cnt->points_valid, (uint32_t)__popc(act));
    if (in_b) atomicAdd(&cnt->points_in_range, (uint32_t)__popc(in_b));
  }
}

// K1 (cloud): pcl::removeNaNFromPointCloud + pcl::transformPointCloud(Matrix4d) + classify.
__global__ void __launch_bounds__(256)
k_front_cloud(DevParams P, ScanBuffers sb, Counters* cnt, Transform T, const uint8_t* __restrict__ in,
              uint32_t n, uint32_t stride, int is_dense) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(in + (size_t)i * stride);
  const float fx = p[0], fy = p[1], fz = p[2];
  if (!is_dense && !(isfinite(fx) && isfinite(fy) && isfinite(fz))) {
    sb.flags[i] = 0;
    sb.world[3 * i] = fx; sb.world[3 * i + 1] = fy; sb.world[3 * i + 2] = fz;
    return;
  }
  const double x = fx, y = fy, z = fz;
  // static_cast<float>(m(r,0)*x + m(r,1)*y + m(r,2)*z + m(r,3)), left to right, double
  const float wx = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[0], x), __dmul_rn(T.m[1], y)), __dmul_rn(T.m[2], z)), T.m[3]);
  const float wy = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[4], x), __dmul_rn(T.m[5], y)), __dmul_rn(T.m[6], z)), T.m[7]);
  const float wz = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[8], x), __dmul_rn(T.m[9], y)), __dmul_rn(T.m[10], z)), T.m[11]);
  classify_point(P, sb, cnt, T.origin, i, wx, wy, wz);
}

// Eigen::Quaterniond::_transformVector(v) + t  (kindr QuatTransformation::operator*)
__device__ __forceinline__ void quat_transform(const Transform& T, double vx, double vy, double vz,
                                               double* ox, double* oy, double* oz) {
  const double w = T.q[0], x = T.q[1], y = T.q[2], z = T.q[3];
  double u0 = __dsub_rn(__dmul_rn(y, vz), __dmul_rn(z, vy));
  double u1 = __dsub_rn(__dmul_rn(z, vx), __dmul_rn(x, vz));
  double u2 = __dsub_rn(__dmul_rn(x, vy), __dmul_rn(y, vx));
  u0 = __dadd_rn(u0, u0); u1 = __dadd_rn(u1, u1); u2 = __dadd_rn(u2, u2);
  const double c0 = __dsub_rn(__dmul_rn(y, u2), __dmul_rn(z, u1));
  const double c1 = __dsub_rn(__dmul_rn(z, u0), __dmul_rn(x, u2));
  const double c2 = __dsub_rn(__dmul_rn(x, u1), __dmul_rn(y, u0));
  *ox = __dadd_rn(__dadd_rn(__dadd_rn(vx, __dmul_rn(w, u0)), c0), T.t[0]);
  *oy = __dadd_rn(__dadd_rn(__dadd_rn(vy, __dmul_rn(w, u1)), c1), T.t[1]);
  *oz = __dadd_rn(__dadd_rn(__dadd_rn(vz, __dmul_rn(w, u2)), c2), T.t[2]);
}

__device__ __forceinline__ void projected_point(const DevParams& P, const ScanBuffers& sb,
                                                Counters* cnt, const Transform& T, uint32_t i,
                                                float sx, float sy, float sz) {
  // isValidPoint(pt) && !(pt.z < 0)   (octomap_world.cc:156, 232-236)
  const bool valid = (sz != 10000.0f) && !isinf(sz) && !(sz < 0.0f);
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

// cv::reprojectImageTo3D, OpenCV 4.13 arithmetic: one pixel.
__device__ __forceinline__ void reproject_pixel(const double* Q, double min_disp, uint32_t u,
                                                uint32_t v, float disp, float* X, float* Y, float* Z) {
  const double x = (double)u, y = (double)v, d = (double)disp;
  double h[4];
#pragma unroll
  for (int r = 0; r < 4; r++)
    h[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Q[4 * r], x), __dmul_rn(Q[4 * r + 1], y)),
                               __dmul_rn(Q[4 * r + 2], d)),
                     __dmul_rn(Q[4 * r + 3], 1.0));
  const double iw = __ddiv_rn(1.0, h[3]);
  *X = (float)__dmul_rn((double)(float)h[0], iw);
  *Y = (float)__dmul_rn((double)(float)h[1], iw);
  *Z = (float)__dmul_rn((double)(float)h[2], iw);
  if (fabs(__dsub_rn(d, min_disp)) <= 1.1920928955078125e-07) *Z = 10000.0f;
}

struct QMat { double q[16]; };

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

// mark voxel (kx,ky,kz) in a block mask array
__device__ __forceinline__ void mark_voxel(const MapView& m, Counters* cnt, uint32_t* masks,
                                           uint32_t kx, uint32_t ky, uint32_t kz) {
  const uint32_t h = find_or_insert_touch(m, cnt, block_key(kx, ky, kz));
  if (h == kNotFound) return;
  const uint32_t l = local_index(kx, ky, kz);
  atomicOr(masks + (size_t)h * kMaskWords + (l >> 5), 1u << (l & 31u));
}

// K1b: skip rule + work list + occupied marks.
//   occupied_cells = { K_i : valid, in range, in bounds };  first[K] = min such i.
//   castRay(i) is invoked  <=>  K_i not yet in occupied_cells when point i is visited
//                           <=>  in-range&bound ? i == first[K_i] : i < first[K_i].
__global__ void __launch_bounds__(256)
k_resolve(ScanBuffers sb, MapView m, Counters* cnt, uint32_t n, int mark_occupied) {
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
    if (mark_occupied)
      mark_voxel(m, cnt, m.occ_mask, (uint32_t)(key & 0xFFFFu), (uint32_t)((key >> 16) & 0xFFFFu),
                 (uint32_t)((key >> 32) & 0xFFFFu));
    else
      sb.flags[i] |= 8u;   // kFlagOcc (vmap_sort.cuh): emitted as a packed key by the sort mode
  }
  // order-preserving-within-warp compaction of the cast list
  const unsigned cast_b = __ballot_sync(0xFFFFFFFFu, cast);
  const unsigned occ_b = __ballot_sync(0xFFFFFFFFu, occ);
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0 && cast_b) {
    base = atomicAdd(&cnt->n_rays, (uint32_t)__popc(cast_b));
    atomicAdd(&cnt->rays_cast, (uint32_t)__popc(cast_b));
    if (occ_b) atomicAdd(&cnt->occupied_emitted, (uint32_t)__popc(occ_b));
  }
  base = __shfl_sync(0xFFFFFFFFu, base, 0);
  if (cast) sb.rays[base + __popc(cast_b & ((1u << lane) - 1u))] = i;
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
      uint32_t word = 0, bits = 0;
      for (;;) {
        ++count;
        if (!kFilter || free_filter_pass(P, ox, oy, oz, w.c0, w.c1, w.c2)) {
          const uint32_t l = local_index(w.c0, w.c1, w.c2);
          if ((l >> 5) != word) {
            if (bits) atomicOr(mask + word, bits);
            word = l >> 5;
            bits = 0;
          }
          bits |= 1u << (l & 31u);
        }
        if (!w.advance()) break;
        if (count >= kKeyRayMax - 1u) { w.live = false; break; }   // KeyRay capacity
        if (((w.c0 >> 3) != b0) | ((w.c1 >> 3) != b1) | ((w.c2 >> 3) != b2)) break;
      }
      if (bits) atomicOr(mask + word, bits);
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

// updateOccupancy on explicit key lists (vmap_apply_keys): mark phase.
__global__ void __launch_bounds__(256)
k_mark_keys(MapView m, Counters* cnt, const uint16_t* __restrict__ k3, uint32_t n, int occupied) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mark_voxel(m, cnt, occupied ? m.occ_mask : m.free_mask, k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]);
}

// Parity hook: expand this scan's masks into explicit key lists (before k_update clears them).
__global__ void __launch_bounds__(256)
k_capture_keys(MapView m, Counters* cnt, uint16_t* __restrict__ free_k3, uint32_t* n_free,
               uint32_t free_cap, uint16_t* __restrict__ occ_k3, uint32_t* n_occ, uint32_t occ_cap) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < cnt->n_touched; t += warps) {
    const uint32_t h = m.touched[t];
    const unsigned long long b = m.keys[h] >> kEpochBits;
    const uint32_t bx = (uint32_t)(b & 0x1FFFu) << 3, by = (uint32_t)((b >> 13) & 0x1FFFu) << 3,
                   bz = (uint32_t)((b >> 26) & 0x1FFFu) << 3;
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t o = m.occ_mask[(size_t)h * kMaskWords + j];
      const uint32_t f = m.free_mask[(size_t)h * kMaskWords + j] & ~o;
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
__global__ void __launch_bounds__(256)
k_update(DevParams P, MapView m, Counters* cnt) {
  // Abort the whole scan's update (uniformly) when a map structure must grow first.
  if (cnt->overflow_table) return;
  // every entry owns a slot (single GPU); a sharded table also lists foreign blocks, so bound
  // the slots by what this update can add at most
  const uint32_t slots_bound = m.sharded ? m.persist->n_slots + cnt->n_touched : m.persist->n_entries;
  if (slots_bound > m.pool_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->overflow_pool = 1u;
    return;
  }
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_touched = cnt->n_touched;
  uint32_t n_free = 0, n_occ = 0;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_touched; t += warps) {
    const uint32_t h = m.touched[t];
    uint32_t slot = 0;
    if (lane == 0) {
      slot = m.slot[h];
      if (slot == kUnassigned) {
        slot = atomicAdd(&m.persist->n_slots, 1u);
        m.slot[h] = slot;
      }
    }
    slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
    uint32_t f = 0, o = 0;
    if (lane < kMaskWords) {
      uint32_t* pf = m.free_mask + (size_t)h * kMaskWords + lane;
      uint32_t* po = m.occ_mask + (size_t)h * kMaskWords + lane;
      f = *pf; o = *po;
      if (f) *pf = 0u;
      if (o) *po = 0u;
      f &= ~o;  // occupied wins (octomap_world.cc:252-254)
    }
    uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
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
      if (lane == 0) { n_free += __popc(fw); n_occ += __popc(ow); }
    }
  }
  if (lane == 0) {
    if (n_free) atomicAdd(&cnt->unique_free, n_free);
    if (n_occ) atomicAdd(&cnt->unique_occupied, n_occ);
  }
}

// Drop this scan's marks (used when a scan has to be re-run on a grown table).
__global__ void __launch_bounds__(256) k_clear_masks(MapView m, Counters* cnt) {
  const uint32_t total = cnt->n_touched * kMaskWords;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t h = m.touched[i / kMaskWords];
    m.free_mask[(size_t)h * kMaskWords + (i % kMaskWords)] = 0u;
    m.occ_mask[(size_t)h * kMaskWords + (i % kMaskWords)] = 0u;
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

// ---------------------------------------------------------------------------------------------
// K5 queries
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t voxel_status(const DevParams& P, const MapView& m, uint32_t h,
                                                uint32_t kx, uint32_t ky, uint32_t kz) {
  uint32_t bits = kUnknownBits;
  if (h != kNotFound) {
    const uint32_t slot = m.slot[h];
    if (slot != kUnassigned)
      bits = reinterpret_cast<const uint32_t*>(m.pool)[(size_t)slot * kBlockVoxels + local_index(kx, ky, kz)];
  }
  if (bits == kUnknownBits) return P.treat_unknown_as_occupied ? 1 : 2;
  return (__uint_as_float(bits) >= P.occ_thres) ? 1 : 0;   // isNodeOccupied
}

// getCellStatusPoint: octree_->search(x,y,z) quantises the DOUBLE coordinates.
__global__ void __launch_bounds__(256)
k_query_points(DevParams P, MapView m, const double* __restrict__ xyz, uint32_t n,
               uint8_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s0 = scaled_coord(P.res_factor, xyz[3 * i]);
  const int s1 = scaled_coord(P.res_factor, xyz[3 * i + 1]);
  const int s2 = scaled_coord(P.res_factor, xyz[3 * i + 2]);
  if (!(key_in_range(s0) && key_in_range(s1) && key_in_range(s2))) {
    status[i] = P.treat_unknown_as_occupied ? 1 : 2;   // search() returns NULL
    return;
  }
  const uint32_t h = find_block(m, block_key(s0, s1, s2));
  status[i] = voxel_status(P, m, h, s0, s1, s2);
}

// getLineStatus: status of the first non-free voxel in DDA order, end voxel excluded.
__global__ void __launch_bounds__(128)
k_query_lines(DevParams P, MapView m, const double* __restrict__ ab, uint32_t n,
              uint8_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // pointEigenToOctomap narrows to float
  const float ax = (float)ab[6 * i], ay = (float)ab[6 * i + 1], az = (float)ab[6 * i + 2];
  const float bx = (float)ab[6 * i + 3], by = (float)ab[6 * i + 4], bz = (float)ab[6 * i + 5];
  uint8_t st = 0;
  unsigned long long cached_b = kEmpty;
  uint32_t cached_h = kNotFound;
  walk_ray(P, ax, ay, az, bx, by, bz, [&](uint32_t kx, uint32_t ky, uint32_t kz) {
    const unsigned long long b = block_key(kx, ky, kz);
    if (b != cached_b) { cached_b = b; cached_h = find_block(m, b); }
    const uint8_t s = voxel_status(P, m, cached_h, kx, ky, kz);
    if (s != 0) { st = s; return false; }
    return true;
  });
  status[i] = st;
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
