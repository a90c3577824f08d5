// vmap_query.cuh -- per-lane body of the bounding-box status query (SURVEY 8 f1:
// getCellStatusBoundingBox behind checkCollisionWithRobot / checkPathForCollisionsWithRobot,
// octomap_world.cc:274-344, 1372-1412).  Like vmap_walk.cuh it uses no block- or warp-level
// primitives, so the CPU test-suite runs the same source against the oracle
// (tests/native/query_check.cc).
//
// The reference walks octomap's pruned pointer tree; here the map is flat (8^3 blocks of
// finest-level voxels), so the two tree walks are restated on voxels:
//   * leaf_bbx_iterator + isNodeOccupied: for every occupied voxel in the key range the LEAF of
//     the canonical pruned tree that contains it is reconstructed (the largest aligned cube of
//     bit-identical known voxels, leaf_level below) and put through the reference's own two
//     tests with that leaf's centre key, half width and cube: the iterator's overlap test --
//     which is conservative by one key on the lower side for leaves above the finest depth, so
//     the scan starts one voxel below minKey -- and the "is it really inside" test in double
//     arithmetic;
//   * getUnknownLeafCenters: octomap's float loop, step added BEFORE each query (the layer of
//     the min key is never tested), every centre looked up through coordToKeyChecked.
// This is exact whenever the reference's tree is canonically pruned, which updateNode keeps true
// on every insertion path; after lazy box edits (setNodeValue with lazy_eval, no prune) the
// reference's tree may hold eight equal un-merged siblings and the lower-side quirk then depends
// on that history.  isSpeckleNode (:858-880) never fires for an occupied
// node as written upstream (it searches `key`, not `current_key`), so filter_speckles has no
// effect here either.
#pragma once

#include "vmap_device.cuh"

namespace vmapdev {

// ---------------------------------------------------------------------------------------------
// K5 per-thread bodies: point, line and visibility queries
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
//   true_status == 0 : getCellStatusPoint      (octomap_world.cc:346-360; unknown -> occupied when configured)
//   true_status == 1 : getCellTrueStatusPoint  (:363-373; unknown stays unknown)
// *bits_out: the leaf's value bits (kUnknownBits without a node).
__device__ __forceinline__ uint8_t point_status(const DevParams& P, const MapView& m, double x, double y, double z,
                                                int true_status, uint32_t* bits_out) {
  const int s0 = scaled_coord(P.res_factor, x);
  const int s1 = scaled_coord(P.res_factor, y);
  const int s2 = scaled_coord(P.res_factor, z);
  uint32_t bits = kUnknownBits;
  if (key_in_range(s0) && key_in_range(s1) && key_in_range(s2)) {   // else search() returns NULL
    const uint32_t h = find_block(m, block_key(s0, s1, s2));
    if (h != kNotFound) {
      const uint32_t slot = m.slot[h];
      if (slot != kUnassigned)
        bits = reinterpret_cast<const uint32_t*>(m.pool)[(size_t)slot * kBlockVoxels + local_index(s0, s1, s2)];
    }
  }
  uint8_t st;
  if (bits == kUnknownBits) st = (P.treat_unknown_as_occupied && !true_status) ? 1 : 2;
  else st = (__uint_as_float(bits) >= P.occ_thres) ? 1 : 0;
  *bits_out = bits;
  return st;
}

// getLineStatus on float end points: status of the first non-free voxel in DDA order, end voxel
// excluded (octomap_world.cc:395-418).
__device__ __forceinline__ uint8_t line_status(const DevParams& P, const MapView& m, float ax, float ay,
                                               float az, float bx, float by, float bz) {
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
  return st;
}

// getVisibility (octomap_world.cc:420-449): occupied (or, when asked, unknown) voxels between the
// view point and the voxel under test, the voxel under test itself excepted.  Unknown is reported
// as unknown here whatever treat_unknown_as_occupied says, like the reference.
__device__ __forceinline__ uint8_t visibility_status(const DevParams& P, const MapView& m, float ax, float ay, float az,
                                                     float bx, float by, float bz, int stop_at_unknown) {
  // coordToKey (unchecked, wraps) of the voxel under test
  const uint32_t t0 = (uint32_t)scaled_coord(P.res_factor, (double)bx) & 0xFFFFu;
  const uint32_t t1 = (uint32_t)scaled_coord(P.res_factor, (double)by) & 0xFFFFu;
  const uint32_t t2 = (uint32_t)scaled_coord(P.res_factor, (double)bz) & 0xFFFFu;
  uint8_t st = 0;
  unsigned long long cached_b = kEmpty;
  uint32_t cached_h = kNotFound;
  walk_ray(P, ax, ay, az, bx, by, bz, [&](uint32_t kx, uint32_t ky, uint32_t kz) {
    if (kx == t0 && ky == t1 && kz == t2) return true;
    const unsigned long long b = block_key(kx, ky, kz);
    if (b != cached_b) { cached_b = b; cached_h = find_block(m, b); }
    uint32_t bits = kUnknownBits;
    if (cached_h != kNotFound) {
      const uint32_t slot = m.slot[cached_h];
      if (slot != kUnassigned)
        bits = reinterpret_cast<const uint32_t*>(m.pool)[(size_t)slot * kBlockVoxels + local_index(kx, ky, kz)];
    }
    if (bits == kUnknownBits) {
      if (stop_at_unknown) { st = 2; return false; }
    } else if (__uint_as_float(bits) >= P.occ_thres) {
      st = 1;
      return false;
    }
    return true;
  });
  return st;
}

struct BoxQuery {
  double size[3];   // bounding_box_size
};
struct BoxScan {
  uint8_t decided;  // 0xFF: the scans below decide; else the status is final (same on every lane)
  bool occupied;    // this lane saw an occupied voxel inside the box
  bool unknown;     // this lane saw an unknown leaf centre
};

// leaf value bits of voxel (kx, ky, kz), kUnknownBits without a node; one-entry block cache
__device__ __forceinline__ uint32_t leaf_bits_cached(const MapView& m, uint32_t kx, uint32_t ky, uint32_t kz,
                                                     unsigned long long* cached_b, uint32_t* cached_h) {
  const unsigned long long b = block_key(kx, ky, kz);
  if (b != *cached_b) {
    *cached_b = b;
    *cached_h = find_block(m, b);
  }
  if (*cached_h == kNotFound) return kUnknownBits;
  const uint32_t slot = m.slot[*cached_h];
  if (slot == kUnassigned) return kUnknownBits;
  return reinterpret_cast<const uint32_t*>(m.pool)[(size_t)slot * kBlockVoxels + local_index(kx, ky, kz)];
}

// Level of the canonical pruned leaf that contains voxel (kx, ky, kz) whose value bits are
// `bits`: the largest l such that the aligned cube of 2^l voxels per side around it holds only
// known voxels with exactly these bits (OcTreeBaseImpl::isNodeCollapsible, applied bottom-up).
// Levels 1..3 stay inside the voxel's 8^3 block; higher levels need whole uniform blocks.
__device__ __forceinline__ uint32_t leaf_level(const MapView& m, uint32_t kx, uint32_t ky, uint32_t kz, uint32_t bits) {
  uint32_t level = 0;
  for (uint32_t l = 1; l <= 15u; l++) {
    const uint32_t side = 1u << l, half = side >> 1;
    const uint32_t lo[3] = {kx & ~(side - 1u), ky & ~(side - 1u), kz & ~(side - 1u)};
    // the sub-cube that contains the voxel is known to be uniform: test the other seven
    const uint32_t own = ((kx & half) ? 1u : 0u) | ((ky & half) ? 2u : 0u) | ((kz & half) ? 4u : 0u);
    bool uniform = true;
    for (uint32_t c = 0; c < 8u && uniform; c++) {
      if (c == own) continue;
      const uint32_t o[3] = {lo[0] + ((c & 1u) ? half : 0u), lo[1] + ((c & 2u) ? half : 0u), lo[2] + ((c & 4u) ? half : 0u)};
      if (half <= 4u) {           // sub-cube inside one block
        const uint32_t h = find_block(m, block_key(o[0], o[1], o[2]));
        const uint32_t slot = h == kNotFound ? kUnassigned : m.slot[h];
        if (slot == kUnassigned) { uniform = false; break; }
        const uint32_t* vox = reinterpret_cast<const uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
        for (uint32_t z = 0; z < half && uniform; z++)
          for (uint32_t y = 0; y < half && uniform; y++)
            for (uint32_t x = 0; x < half; x++)
              if (vox[local_index(o[0] + x, o[1] + y, o[2] + z)] != bits) { uniform = false; break; }
      } else {                    // sub-cube = (half / 8)^3 whole blocks
        const uint32_t nb = half >> 3;
        for (uint32_t bz = 0; bz < nb && uniform; bz++)
          for (uint32_t by = 0; by < nb && uniform; by++)
            for (uint32_t bx = 0; bx < nb && uniform; bx++) {
              const uint32_t h = find_block(m, block_key(o[0] + 8u * bx, o[1] + 8u * by, o[2] + 8u * bz));
              const uint32_t slot = h == kNotFound ? kUnassigned : m.slot[h];
              if (slot == kUnassigned) { uniform = false; break; }
              const uint32_t* vox = reinterpret_cast<const uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
              for (uint32_t i = 0; i < (uint32_t)kBlockVoxels; i++)
                if (vox[i] != bits) { uniform = false; break; }
            }
      }
    }
    if (!uniform) break;
    level = l;
  }
  return level;
}

// One lane's share of getCellStatusBoundingBox(point, size).  Lanes 0 .. n_lanes-1 together cover
// the box; the caller ORs `occupied` / `unknown` over the lanes:
//   status = decided != 0xFF ? decided : any(occupied) ? occupied : any(unknown) ? unknown_as : free.
__device__ __forceinline__ BoxScan bbox_scan(const DevParams& P, const MapView& m, double px, double py, double pz,
                                             const BoxQuery& q, uint32_t lane, uint32_t n_lanes) {
  BoxScan r;
  r.decided = 0xFF;
  r.occupied = false;
  r.unknown = false;
  const uint8_t unknown_as = P.treat_unknown_as_occupied ? 1 : 2;
  unsigned long long cb = kEmpty;
  uint32_t ch = kNotFound;
  const double p[3] = {px, py, pz};
  // centre point: getCellStatusPoint quantises the DOUBLE coordinates (:276-280)
  {
    const int s0 = scaled_coord(P.res_factor, px), s1 = scaled_coord(P.res_factor, py), s2 = scaled_coord(P.res_factor, pz);
    uint32_t bits = kUnknownBits;
    if (key_in_range(s0) && key_in_range(s1) && key_in_range(s2))
      bits = leaf_bits_cached(m, (uint32_t)s0, (uint32_t)s1, (uint32_t)s2, &cb, &ch);
    const uint8_t st = bits == kUnknownBits ? unknown_as : ((__uint_as_float(bits) >= P.occ_thres) ? 1 : 0);
    if (st != 0) { r.decided = st; return r; }
  }
  // coordToKeyChecked(pointEigenToOctomap(point)) on the float-narrowed point (:283-290)
  for (int a = 0; a < 3; a++)
    if (!key_in_range(scaled_coord(P.res_factor, (double)(float)p[a]))) { r.decided = unknown_as; return r; }
  // bbx_min / bbx_max: double arithmetic, narrowed to float (:293-297)
  float mn[3], mx[3];
  for (int a = 0; a < 3; a++) {
    const double half = __ddiv_rn(q.size[a], 2.0);
    mn[a] = (float)__dsub_rn(p[a], half);
    mx[a] = (float)__dadd_rn(p[a], half);
  }
  // ---- occupied leaves in the box (:299-330) ----
  int kmin[3], kmax[3];
  bool keys_ok = true;
  for (int a = 0; a < 3; a++) {
    kmin[a] = scaled_coord(P.res_factor, (double)mn[a]);
    kmax[a] = scaled_coord(P.res_factor, (double)mx[a]);
    keys_ok = keys_ok && key_in_range(kmin[a]) && key_in_range(kmax[a]);
  }
  if (keys_ok && kmin[0] <= kmax[0] && kmin[1] <= kmax[1] && kmin[2] <= kmax[2]) {   // else: end iterator / no overlap
    // one extra layer below minKey: a pruned leaf ending there still passes the iterator's test
    const int emin[3] = {kmin[0] > 0 ? kmin[0] - 1 : 0, kmin[1] > 0 ? kmin[1] - 1 : 0, kmin[2] > 0 ? kmin[2] - 1 : 0};
    const unsigned long long nx = (unsigned long long)(kmax[0] - emin[0] + 1), ny = (unsigned long long)(kmax[1] - emin[1] + 1),
                             nz = (unsigned long long)(kmax[2] - emin[2] + 1);
    for (unsigned long long t = lane; t < nx * ny * nz; t += n_lanes) {
      const uint32_t k[3] = {(uint32_t)(emin[0] + (int)(t % nx)), (uint32_t)(emin[1] + (int)((t / nx) % ny)),
                             (uint32_t)(emin[2] + (int)(t / (nx * ny)))};
      const uint32_t bits = leaf_bits_cached(m, k[0], k[1], k[2], &cb, &ch);
      if (bits == kUnknownBits || !(__uint_as_float(bits) >= P.occ_thres)) continue;
      const bool in_range = (int)k[0] >= kmin[0] && (int)k[1] >= kmin[1] && (int)k[2] >= kmin[2];
      // finest-depth leaves in the extra layer fail the (exact) iterator test: only a pruned leaf matters there
      const uint32_t level = leaf_level(m, k[0], k[1], k[2], bits);
      if (level == 0u && !in_range) continue;
      // the leaf's key (centre key above the finest depth), half width in keys, cube size
      const uint32_t side = 1u << level, half_keys = side >> 1;
      const double cube_size = __dmul_rn(P.res, (double)side);               // getNodeSize(depth)
      const double half_cube = __dmul_rn(__ddiv_rn(cube_size, 2.0), 1.0);    // (cube_size / 2) * Ones()
      bool visit = true;   // iterator overlap test, then "Check if it is really inside bounding box"
      for (int a = 0; a < 3; a++) {
        const int ck = (int)((k[a] & ~(side - 1u)) + half_keys);
        if (!(kmin[a] <= ck + (int)half_keys && kmax[a] >= ck - (int)half_keys)) visit = false;
        // keyToCoord(key, depth)
        const double c = level == 0u ? key_to_coord(P.res, k[a])
                                     : __dmul_rn(__dadd_rn(floor(__ddiv_rn(__dsub_rn((double)ck, 32768.0), (double)side)), 0.5), cube_size);
        const double lower = __dsub_rn(c, half_cube), upper = __dadd_rn(c, half_cube);
        if (upper < (double)mn[a] || lower > (double)mx[a]) visit = false;
      }
      if (visit) { r.occupied = true; break; }
    }
  }
  // ---- unknown leaf centres (:332-342; octomap getUnknownLeafCenters, float arithmetic) ----
  float c0[3], step_size = (float)P.res;
  uint32_t steps[3];
  for (int a = 0; a < 3; a++) {
    // keyToCoord(coordToKey(p)): unchecked key (wraps to 16 bits), float-narrowed centre
    const uint32_t k_lo = (uint32_t)scaled_coord(P.res_factor, (double)mn[a]) & 0xFFFFu;
    const uint32_t k_hi = (uint32_t)scaled_coord(P.res_factor, (double)mx[a]) & 0xFFFFu;
    c0[a] = (float)key_to_coord(P.res, k_lo);
    const float diff = __fsub_rn((float)key_to_coord(P.res, k_hi), c0[a]);
    const float fl = floorf(__fdiv_rn(diff, step_size));
    steps[a] = fl > 0.0f ? (uint32_t)fl : 0u;
  }
  const unsigned long long n_xy = (unsigned long long)steps[0] * steps[1];
  if (steps[2] != 0u) {
    for (unsigned long long t = lane; t < n_xy && !r.unknown; t += n_lanes) {
      const uint32_t ix = (uint32_t)(t / steps[1]), iy = (uint32_t)(t % steps[1]);
      // p.x() / p.y() after (ix + 1) / (iy + 1) additions of step_size, one rounding each
      float x = c0[0], y = c0[1], z = c0[2];
      for (uint32_t i = 0; i <= ix; i++) x = __fadd_rn(x, step_size);
      for (uint32_t i = 0; i <= iy; i++) y = __fadd_rn(y, step_size);
      const int sx = scaled_coord(P.res_factor, (double)x), sy = scaled_coord(P.res_factor, (double)y);
      const bool xy_ok = key_in_range(sx) && key_in_range(sy);
      for (uint32_t iz = 0; iz < steps[2]; iz++) {
        z = __fadd_rn(z, step_size);
        const int sz = scaled_coord(P.res_factor, (double)z);
        if (!xy_ok || !key_in_range(sz)) { r.unknown = true; break; }   // search(p) == NULL outside the tree
        if (leaf_bits_cached(m, (uint32_t)sx, (uint32_t)sy, (uint32_t)sz, &cb, &ch) == kUnknownBits) {
          r.unknown = true;
          break;
        }
      }
    }
  }
  return r;
}

__device__ __forceinline__ uint8_t bbox_combine(const DevParams& P, uint8_t decided, bool any_occupied, bool any_unknown) {
  if (decided != 0xFF) return decided;
  if (any_occupied) return 1;
  if (any_unknown) return P.treat_unknown_as_occupied ? 1 : 2;
  return 0;
}

}  // namespace vmapdev
