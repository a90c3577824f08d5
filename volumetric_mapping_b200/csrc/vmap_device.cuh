// vmap_device.cuh -- device-side building blocks shared by the scan-insertion and query kernels.
//
// Arithmetic in this file is written with explicit round-to-nearest intrinsics (__dadd_rn,
// __fmul_rn, ...) wherever the reference's results depend on the exact IEEE operation order:
// those intrinsics are never contracted into FMAs, so the bit patterns match the x86 build of
// the reference (octomap / PCL / Eigen compiled without FMA) irrespective of nvcc's -fmad flag.
#pragma once

#if !defined(VMAP_HOST_EMULATION)   // (tests/native/cuda_host_shim.h stands in for the CUDA headers)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace vmapdev {

// ---------------------------------------------------------------------------------------------
// Parameters that every kernel needs (derived from vmap_params on the host; passed by value).
// ---------------------------------------------------------------------------------------------
struct DevParams {
  double res;           // OcTree::resolution
  double res_factor;    // 1.0 / resolution
  double max_range;     // sensor_max_range (negative: no limit)
  double max_free_space;
  double min_height_free_space;
  float hit, miss, cmin, cmax, occ_thres;  // (float)log(p/(1-p))
  int treat_unknown_as_occupied;
  int free_filter;      // max_free_space != 0.0
};

// ---------------------------------------------------------------------------------------------
// Sparse voxel-block hash.  One entry per 8x8x8 block of finest-level voxels.
//   keys[h]      = (block_key39 << 25) | scan_epoch25, or kEmpty
//   slot[h]      = index of the block's 512 floats in `pool` (kUnassigned until first update)
//   free_mask/occ_mask[h*16 .. h*16+15] = this scan's free / occupied voxel sets (512 bits each)
//   touched[]    = entries whose epoch was moved to the current scan (the per-scan work list)
// Leaf log-odds are float32; the bit pattern 0xFFFFFFFF (a NaN no update can produce) marks a
// voxel that was never updated ("unknown", i.e. no node in the reference's tree).
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kUnassigned = 0xFFFFFFFFu;
constexpr uint32_t kUnknownBits = 0xFFFFFFFFu;
constexpr int kEpochBits = 25;
constexpr uint32_t kEpochMask = (1u << kEpochBits) - 1u;
constexpr uint32_t kEpochMax = kEpochMask - 1u;  // never store all-ones (would alias kEmpty)
constexpr uint32_t kNotFound = 0xFFFFFFFFu;
constexpr int kMaxProbe = 2048;
constexpr int kBlockVoxels = 512;
constexpr int kMaskWords = 16;
constexpr uint32_t kKeyRayMax = 100000;  // octomap::KeyRay::sizeMax

struct Persist {            // survives across scans
  uint32_t n_entries;       // occupied hash entries (blocks known to the table)
  uint32_t n_slots;         // pool slots handed out
  uint32_t stalled;         // pipelined insertion: an earlier scan's update was refused (the map must
                            // grow first); every later update refuses too until the host has replayed it
  uint32_t pad;
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization
// attribute may become resident while its predecessor in the stream is still running; pdl_wait()
// blocks until that predecessor has completed and its memory is visible (a no-op for an ordinary
// launch), pdl_launch() lets the successor's CTAs in early.  Every kernel of the scan chain calls
// pdl_wait() before it touches global memory.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

struct Counters {           // zeroed at the start of every scan
  unsigned long long traversals;
  uint32_t points_valid;
  uint32_t points_in_range;
  uint32_t rays_cast;
  uint32_t n_rays;          // length of the cast-ray work list
  uint32_t occupied_emitted;
  uint32_t unique_free;
  uint32_t unique_occupied;
  uint32_t longest_ray;
  uint32_t n_touched;
  uint32_t overflow_table;  // probe limit hit: table must grow, scan is re-run
  uint32_t overflow_pool;   // pool too small for the new blocks: pool must grow, update re-run
  uint32_t overflow_keys;   // sort mode: key buffer too small
  uint32_t overflow_window; // dense mode: a voxel fell outside the window
  uint32_t overflow_ckpt;   // dense mode: segment checkpoint buffer too small
  uint32_t overflow_pack;   // sharded: record buffer too small (value = records needed)
  uint32_t n_keys;          // sort mode: packed keys emitted
  uint32_t n_leaves;        // export
  uint32_t min_disparity_ord;  // ordered-int encoding of the image minimum
  uint32_t stall;           // the scan's update was refused as a whole (kStall* bits); its marks are intact
  uint32_t need_slots;      // listed blocks that have no pool slot yet (k_locate_blocks)
  uint32_t done_ctas;       // CTAs of k_locate_blocks that have finished (last one decides about the pool)
  uint32_t occ_cursor;      // sort mode: occupied keys appended after the free ones
  uint32_t round_blocks;    // sharded apply: distinct blocks named by the round's records
  uint32_t done_resolve;    // CTAs of k_resolve_fold that have finished (last one builds bin_off / seg_base)
  uint32_t done_list;       // CTAs of k_list_touched that have finished (last one decides the admission)
};
constexpr uint32_t kStallEarlier = 1u;   // an earlier scan is stalled (Persist::stalled)
constexpr uint32_t kStallMarks = 2u;     // the marks are incomplete (window / checkpoint / list overflow)
constexpr uint32_t kStallTable = 4u;     // the table could exceed load 1/2
constexpr uint32_t kStallPool = 8u;      // not enough free pool slots

struct MapView {
  unsigned long long* keys;
  uint32_t* slot;
  uint32_t* free_mask;
  uint32_t* occ_mask;
  uint32_t* touched;
  float* pool;
  Persist* persist;
  uint32_t cap_mask;   // capacity - 1 (capacity is a power of two)
  uint32_t pool_cap;   // blocks
  uint32_t epoch;      // current scan epoch (1..kEpochMax)
  uint32_t sharded;    // 1: the table also holds blocks owned by other ranks (no pool slot)
  // octomap's change detection (changed_keys), two bit planes per pool block (16 words each):
  // code 1 = node created since the last reset (value `true` in the KeyBoolMap), code 2 =
  // occupancy flipped (`false`), 0 = not in the map.  NULL when change detection is off.
  uint32_t* chg0;
  uint32_t* chg1;
};

__device__ __forceinline__ unsigned long long block_key(uint32_t kx, uint32_t ky, uint32_t kz) {
  return (unsigned long long)(kx >> 3) | ((unsigned long long)(ky >> 3) << 13) |
         ((unsigned long long)(kz >> 3) << 26);
}
__device__ __forceinline__ uint32_t local_index(uint32_t kx, uint32_t ky, uint32_t kz) {
  return (kx & 7u) | ((ky & 7u) << 3) | ((kz & 7u) << 6);
}
__device__ __forceinline__ uint32_t hash_block(unsigned long long b) {
  // 64-bit finaliser (splitmix64): neighbouring blocks land in unrelated table lines
  b ^= b >> 30; b *= 0xbf58476d1ce4e5b9ull;
  b ^= b >> 27; b *= 0x94d049bb133111ebull;
  b ^= b >> 31;
  return (uint32_t)b;
}

__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* p) {
  return __ldcg(p);
}

// Find the entry of block `b`, inserting it when absent, and make sure it is on this scan's
// touched list.  Returns the entry index or kNotFound when the probe limit is hit (table full:
// cnt->overflow_table is raised and the host re-runs the scan on a larger table).
__device__ __forceinline__ uint32_t find_or_insert_touch(const MapView& m, Counters* cnt,
                                                         unsigned long long b) {
  uint32_t h = hash_block(b) & m.cap_mask;
  const unsigned long long want = (b << kEpochBits) | (unsigned long long)m.epoch;
  unsigned long long w = ld_cg_u64(m.keys + h);
  for (int probe = 0; probe < kMaxProbe;) {
    if (w == kEmpty) {
      const unsigned long long old = atomicCAS(m.keys + h, kEmpty, want);
      if (old == kEmpty) {
        atomicAdd(&m.persist->n_entries, 1u);
        m.touched[atomicAdd(&cnt->n_touched, 1u)] = h;
        return h;
      }
      w = old;  // somebody else filled the slot: re-examine it
      continue;
    }
    if ((w >> kEpochBits) == b) {
      if ((uint32_t)(w & kEpochMask) != m.epoch) {
        const unsigned long long old = atomicCAS(m.keys + h, w, want);
        if (old == w) m.touched[atomicAdd(&cnt->n_touched, 1u)] = h;
        // else: another thread moved the epoch (the key part never changes)
      }
      return h;
    }
    h = (h + 1u) & m.cap_mask;
    ++probe;
    w = ld_cg_u64(m.keys + h);
  }
  atomicExch(&cnt->overflow_table, 1u);
  return kNotFound;
}

// Read-only lookup (queries, export, value writes after the blocks exist).
__device__ __forceinline__ uint32_t find_block(const MapView& m, unsigned long long b) {
  uint32_t h = hash_block(b) & m.cap_mask;
  for (int probe = 0; probe < kMaxProbe; ++probe) {
    const unsigned long long w = m.keys[h];
    if (w == kEmpty) return kNotFound;
    if ((w >> kEpochBits) == b) return h;
    h = (h + 1u) & m.cap_mask;
  }
  return kNotFound;
}

// Insert-or-find without touching the per-scan list (dense-window mode locates each touched
// block once per scan, after the walk).
__device__ __forceinline__ uint32_t find_or_insert(const MapView& m, Counters* cnt,
                                                   unsigned long long b) {
  uint32_t h = hash_block(b) & m.cap_mask;
  const unsigned long long want = (b << kEpochBits) | (unsigned long long)m.epoch;
  unsigned long long w = ld_cg_u64(m.keys + h);
  for (int probe = 0; probe < kMaxProbe;) {
    if (w == kEmpty) {
      const unsigned long long old = atomicCAS(m.keys + h, kEmpty, want);
      if (old == kEmpty) {
        atomicAdd(&m.persist->n_entries, 1u);
        return h;
      }
      w = old;
      continue;
    }
    if ((w >> kEpochBits) == b) return h;
    h = (h + 1u) & m.cap_mask;
    ++probe;
    w = ld_cg_u64(m.keys + h);
  }
  atomicExch(&cnt->overflow_table, 1u);
  return kNotFound;
}

// ---------------------------------------------------------------------------------------------
// Where the marks (free / occupied voxel sets) of the scan in flight live.
//   table mode : mask columns of the hash table, entry = table index, list = m.touched
//   dense mode : a directly addressed window of 8^3 blocks around the sensor (x fastest), entry =
//                window block index, list = compacted from the window's touched bitmap; `locate`
//                holds the table index of every listed block (k_locate_blocks)
// ---------------------------------------------------------------------------------------------
struct MarkSet {
  uint32_t* free_mask;       // [entries * kMaskWords]
  uint32_t* occ_mask;        // [entries * kMaskWords]
  uint32_t* list;            // [n_touched] entries
  uint32_t* locate;          // dense: [n_touched] table indices
  uint32_t* touched_bits;    // dense: 1 bit per window block
  int x0, y0, z0;            // dense: window origin in block units
  uint32_t nx, ny, nz;       // dense: window size in blocks
  uint32_t list_cap;
  uint32_t dense;
};

// window block index of the block containing key (kx,ky,kz), or kNotFound outside the window
__device__ __forceinline__ uint32_t window_index(const MarkSet& ms, uint32_t kx, uint32_t ky, uint32_t kz) {
  const uint32_t rx = (uint32_t)((int)(kx >> 3) - ms.x0);
  const uint32_t ry = (uint32_t)((int)(ky >> 3) - ms.y0);
  const uint32_t rz = (uint32_t)((int)(kz >> 3) - ms.z0);
  if (rx >= ms.nx || ry >= ms.ny || rz >= ms.nz) return kNotFound;
  return (rz * ms.ny + ry) * ms.nx + rx;
}
__device__ __forceinline__ unsigned long long window_block_key(const MarkSet& ms, uint32_t idx) {
  const uint32_t rx = idx % ms.nx;
  const uint32_t t = idx / ms.nx;
  const uint32_t ry = t % ms.ny, rz = t / ms.ny;
  return (unsigned long long)(uint32_t)((int)rx + ms.x0) | ((unsigned long long)(uint32_t)((int)ry + ms.y0) << 13) |
         ((unsigned long long)(uint32_t)((int)rz + ms.z0) << 26);
}
// block key and table index of the t-th listed entry
__device__ __forceinline__ unsigned long long listed_block(const MarkSet& ms, const MapView& m, uint32_t t,
                                                           uint32_t* entry, uint32_t* table_h) {
  const uint32_t e = ms.list[t];
  *entry = e;
  if (ms.dense) {
    *table_h = ms.locate[t];
    return window_block_key(ms, e);
  }
  *table_h = e;
  return m.keys[e] >> kEpochBits;
}

// ---------------------------------------------------------------------------------------------
// octomap key arithmetic (OcTreeBaseImpl::coordToKey / coordToKeyChecked / keyToCoord).
// ---------------------------------------------------------------------------------------------
// x86-64 `(int)floor(v)` is cvttsd2si: NaN and out-of-range produce INT_MIN.
__device__ __forceinline__ int x86_floor_to_int(double v) {
  const double f = floor(v);
  if (!(f >= -2147483648.0 && f < 2147483648.0)) return (int)0x80000000;
  return (int)f;
}
// scaled coordinate before the range check / uint16 wrap
__device__ __forceinline__ int scaled_coord(double res_factor, double c) {
  return x86_floor_to_int(__dmul_rn(res_factor, c)) + 32768;
}
__device__ __forceinline__ bool key_in_range(int s) { return s >= 0 && (unsigned)s < 65536u; }
__device__ __forceinline__ double key_to_coord(double res, uint32_t k) {
  return __dmul_rn(__dadd_rn((double)((int)k - 32768), 0.5), res);
}

// float3 helpers with octomath::Vector3 semantics (every op rounded to float, no FMA)
__device__ __forceinline__ float norm_sq_f(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// ---------------------------------------------------------------------------------------------
// Amanatides-Woo DDA over OcTreeKey space == octomap::OcTreeBaseImpl::computeRayKeys.
// `emit(kx,ky,kz)` is called for the origin key and every intermediate key, never for the end
// key.  Returns the number of keys emitted (0 also when computeRayKeys would return false).
// `visit` may return false to stop early (queries).
// ---------------------------------------------------------------------------------------------
template <class Visit>
__device__ __forceinline__ uint32_t walk_ray(const DevParams& P, float ox, float oy, float oz,
                                             float ex, float ey, float ez, Visit&& visit) {
  const int so0 = scaled_coord(P.res_factor, (double)ox);
  const int so1 = scaled_coord(P.res_factor, (double)oy);
  const int so2 = scaled_coord(P.res_factor, (double)oz);
  const int se0 = scaled_coord(P.res_factor, (double)ex);
  const int se1 = scaled_coord(P.res_factor, (double)ey);
  const int se2 = scaled_coord(P.res_factor, (double)ez);
  if (!(key_in_range(so0) && key_in_range(so1) && key_in_range(so2) && key_in_range(se0) &&
        key_in_range(se1) && key_in_range(se2)))
    return 0;
  uint32_t c0 = (uint32_t)so0, c1 = (uint32_t)so1, c2 = (uint32_t)so2;
  const uint32_t e0 = (uint32_t)se0, e1 = (uint32_t)se1, e2 = (uint32_t)se2;
  if (c0 == e0 && c1 == e1 && c2 == e2) return 0;

  uint32_t count = 1;
  if (!visit(c0, c1, c2)) return count;

  float d0 = __fsub_rn(ex, ox), d1 = __fsub_rn(ey, oy), d2 = __fsub_rn(ez, oz);
  const float length = (float)__dsqrt_rn((double)norm_sq_f(d0, d1, d2));
  d0 = __fdiv_rn(d0, length);
  d1 = __fdiv_rn(d1, length);
  d2 = __fdiv_rn(d2, length);

  const double kDblMax = 1.7976931348623157e308;
  int s0 = d0 > 0.0f ? 1 : (d0 < 0.0f ? -1 : 0);
  int s1 = d1 > 0.0f ? 1 : (d1 < 0.0f ? -1 : 0);
  int s2 = d2 > 0.0f ? 1 : (d2 < 0.0f ? -1 : 0);
  double t0 = kDblMax, t1 = kDblMax, t2 = kDblMax, dt0 = kDblMax, dt1 = kDblMax, dt2 = kDblMax;
  // voxelBorder = keyToCoord(key) + (float)(step * res * 0.5)
  if (s0 != 0) {
    const double border =
        __dadd_rn(key_to_coord(P.res, c0), (double)(float)__dmul_rn(__dmul_rn((double)s0, P.res), 0.5));
    t0 = __ddiv_rn(__dsub_rn(border, (double)ox), (double)d0);
    dt0 = __ddiv_rn(P.res, fabs((double)d0));
  }
  if (s1 != 0) {
    const double border =
        __dadd_rn(key_to_coord(P.res, c1), (double)(float)__dmul_rn(__dmul_rn((double)s1, P.res), 0.5));
    t1 = __ddiv_rn(__dsub_rn(border, (double)oy), (double)d1);
    dt1 = __ddiv_rn(P.res, fabs((double)d1));
  }
  if (s2 != 0) {
    const double border =
        __dadd_rn(key_to_coord(P.res, c2), (double)(float)__dmul_rn(__dmul_rn((double)s2, P.res), 0.5));
    t2 = __ddiv_rn(__dsub_rn(border, (double)oz), (double)d2);
    dt2 = __ddiv_rn(P.res, fabs((double)d2));
  }
  const double dlen = (double)length;

  for (;;) {
    // find minimum tMax (strict '<': ties go to the later axis)
    if (t0 < t1) {
      if (t0 < t2) { c0 = (c0 + (uint32_t)s0) & 0xFFFFu; t0 = __dadd_rn(t0, dt0); }
      else         { c2 = (c2 + (uint32_t)s2) & 0xFFFFu; t2 = __dadd_rn(t2, dt2); }
    } else {
      if (t1 < t2) { c1 = (c1 + (uint32_t)s1) & 0xFFFFu; t1 = __dadd_rn(t1, dt1); }
      else         { c2 = (c2 + (uint32_t)s2) & 0xFFFFu; t2 = __dadd_rn(t2, dt2); }
    }
    if (c0 == e0 && c1 == e1 && c2 == e2) break;
    // std::min(std::min(t0,t1),t2) with std::min(a,b) = (b < a) ? b : a
    const double m01 = (t1 < t0) ? t1 : t0;
    const double dist = (t2 < m01) ? t2 : m01;
    if (dist > dlen) break;
    if (count >= kKeyRayMax - 1u) break;  // KeyRay capacity (octomap asserts; see oracle.cc)
    ++count;
    if (!visit(c0, c1, c2)) break;
  }
  return count;
}

// ---------------------------------------------------------------------------------------------
// The same DDA as an explicit state machine, written branch-free so that the lanes of a warp
// (one ray each) stay converged: `emit pending voxel -> advance()` reproduces computeRayKeys'
//   addKey(origin); loop { pick dim; step; if (cur==end) stop; if (min(tMax) > length) stop;
//                          addKey(cur); }
// exactly, including the tie rule (strict '<', later axis wins) and the sequential double
// accumulation of tMax.  min(tMax) of one step is the value the next step's axis choice
// selects, so both are computed once.
// ---------------------------------------------------------------------------------------------
struct RayWalk {
  uint32_t c0, c1, c2;   // current key (the voxel that is emitted next while `live`)
  uint32_t e0, e1, e2;   // end key (never emitted)
  int s0, s1, s2;
  double t0, t1, t2, dt0, dt1, dt2, dlen;
  bool d0, d1;           // axis choice of the NEXT step (d2 = !d0 && !d1)
  bool live;

  __device__ __forceinline__ void pick() {
    const bool a01 = t0 < t1, a02 = t0 < t2, a12 = t1 < t2;
    d0 = a01 && a02;
    d1 = !a01 && a12;
  }
  __device__ __forceinline__ double tmin() const { return d0 ? t0 : (d1 ? t1 : t2); }

  // computeRayKeys' preamble; live == false when the ray has no keys at all.
  __device__ __forceinline__ void init(const DevParams& P, float ox, float oy, float oz, float ex,
                                       float ey, float ez) {
    live = false;
    const int so0 = scaled_coord(P.res_factor, (double)ox);
    const int so1 = scaled_coord(P.res_factor, (double)oy);
    const int so2 = scaled_coord(P.res_factor, (double)oz);
    const int se0 = scaled_coord(P.res_factor, (double)ex);
    const int se1 = scaled_coord(P.res_factor, (double)ey);
    const int se2 = scaled_coord(P.res_factor, (double)ez);
    c0 = (uint32_t)so0; c1 = (uint32_t)so1; c2 = (uint32_t)so2;
    e0 = (uint32_t)se0; e1 = (uint32_t)se1; e2 = (uint32_t)se2;
    s0 = s1 = s2 = 0;
    const double kDblMax = 1.7976931348623157e308;
    t0 = t1 = t2 = dt0 = dt1 = dt2 = kDblMax;
    dlen = 0.0;
    d0 = d1 = false;
    if (!(key_in_range(so0) && key_in_range(so1) && key_in_range(so2) && key_in_range(se0) &&
          key_in_range(se1) && key_in_range(se2)))
      return;
    if (c0 == e0 && c1 == e1 && c2 == e2) return;
    live = true;
    float f0 = __fsub_rn(ex, ox), f1 = __fsub_rn(ey, oy), f2 = __fsub_rn(ez, oz);
    const float length = (float)__dsqrt_rn((double)norm_sq_f(f0, f1, f2));
    f0 = __fdiv_rn(f0, length);
    f1 = __fdiv_rn(f1, length);
    f2 = __fdiv_rn(f2, length);
    s0 = f0 > 0.0f ? 1 : (f0 < 0.0f ? -1 : 0);
    s1 = f1 > 0.0f ? 1 : (f1 < 0.0f ? -1 : 0);
    s2 = f2 > 0.0f ? 1 : (f2 < 0.0f ? -1 : 0);
    if (s0 != 0) {
      const double border = __dadd_rn(key_to_coord(P.res, c0),
                                      (double)(float)__dmul_rn(__dmul_rn((double)s0, P.res), 0.5));
      t0 = __ddiv_rn(__dsub_rn(border, (double)ox), (double)f0);
      dt0 = __ddiv_rn(P.res, fabs((double)f0));
    }
    if (s1 != 0) {
      const double border = __dadd_rn(key_to_coord(P.res, c1),
                                      (double)(float)__dmul_rn(__dmul_rn((double)s1, P.res), 0.5));
      t1 = __ddiv_rn(__dsub_rn(border, (double)oy), (double)f1);
      dt1 = __ddiv_rn(P.res, fabs((double)f1));
    }
    if (s2 != 0) {
      const double border = __dadd_rn(key_to_coord(P.res, c2),
                                      (double)(float)__dmul_rn(__dmul_rn((double)s2, P.res), 0.5));
      t2 = __ddiv_rn(__dsub_rn(border, (double)oz), (double)f2);
      dt2 = __ddiv_rn(P.res, fabs((double)f2));
    }
    dlen = (double)length;
    pick();
  }

  // One DDA step where computeRayKeys' two termination tests cannot fire (every segment of a
  // ray but its last one, see k_walk_segments).
  __device__ __forceinline__ void advance_inner() {
    const bool d2 = !(d0 || d1);
    c0 = d0 ? ((c0 + (uint32_t)s0) & 0xFFFFu) : c0;
    c1 = d1 ? ((c1 + (uint32_t)s1) & 0xFFFFu) : c1;
    c2 = d2 ? ((c2 + (uint32_t)s2) & 0xFFFFu) : c2;
    t0 = d0 ? __dadd_rn(t0, dt0) : t0;
    t1 = d1 ? __dadd_rn(t1, dt1) : t1;
    t2 = d2 ? __dadd_rn(t2, dt2) : t2;
    pick();
  }

  // One DDA step after the current voxel has been emitted.  Returns false (and clears `live`)
  // when the walk is over, i.e. the new current voxel must NOT be emitted.
  __device__ __forceinline__ bool advance() {
    const bool d2 = !(d0 || d1);
    c0 = d0 ? ((c0 + (uint32_t)s0) & 0xFFFFu) : c0;
    c1 = d1 ? ((c1 + (uint32_t)s1) & 0xFFFFu) : c1;
    c2 = d2 ? ((c2 + (uint32_t)s2) & 0xFFFFu) : c2;
    t0 = d0 ? __dadd_rn(t0, dt0) : t0;
    t1 = d1 ? __dadd_rn(t1, dt1) : t1;
    t2 = d2 ? __dadd_rn(t2, dt2) : t2;
    pick();
    const bool at_end = (c0 == e0) & (c1 == e1) & (c2 == e2);
    live = !at_end && !(tmin() > dlen);
    return live;
  }
};

// castRay's free-space filter (octomap_world.cc:192-200 / 218-226): true -> key is inserted.
__device__ __forceinline__ bool free_filter_pass(const DevParams& P, float ox, float oy, float oz,
                                                 uint32_t kx, uint32_t ky, uint32_t kz) {
  const float vx = (float)key_to_coord(P.res, kx);
  const float vy = (float)key_to_coord(P.res, ky);
  const float vz = (float)key_to_coord(P.res, kz);
  const float dx = __fsub_rn(vx, ox), dy = __fsub_rn(vy, oy), dz = __fsub_rn(vz, oz);
  const double n = __dsqrt_rn((double)norm_sq_f(dx, dy, dz));
  return (n < P.max_free_space) ||
         ((double)vz > __dsub_rn((double)oz, P.min_height_free_space));
}

// updateNode's leaf arithmetic incl. the early abort (OccupancyOcTreeBase::updateNode /
// updateNodeLogOdds); `bits` is the stored word, kUnknownBits when the voxel has no node.
__device__ __forceinline__ uint32_t update_leaf(uint32_t bits, float u, float cmin, float cmax) {
  float v;
  if (bits == kUnknownBits) {
    v = 0.0f;  // new node value
  } else {
    v = __uint_as_float(bits);
    if ((u >= 0.0f && v >= cmax) || (u <= 0.0f && v <= cmin)) return bits;  // early abort
  }
  v = __fadd_rn(v, u);
  if (v < cmin) v = cmin;
  else if (v > cmax) v = cmax;
  return __float_as_uint(v);
}

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

}  // namespace vmapdev
