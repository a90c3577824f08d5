// vmap_walk.cuh -- per-thread bodies of the segmented ray walker (K2a / K2b), kept free of
// block- and warp-level primitives so that the same code also builds for the host: the CPU
// test-suite (tests/native/walk_check.cc) runs it slot by slot against the plain DDA.
//
//   ray_checkpoints      state of one ray's walk at every segment boundary (exact jump-ahead)
//   walk_slot            one (segment level, ray) work slot -> marks in the dense window
//   walk_slot_dilated    the same walk, carrying the window address of the current voxel as three
//                        per-axis bit fields instead of recomputing it from the key at every step
#pragma once

#include "vmap_device.cuh"
#include "vmap_seq.cuh"

namespace vmapdev {

constexpr int kLenBins = 2048;        // cast rays are ordered by segment count: bin = min(n_seg, 2047)
#ifndef VMAP_SEG_STEPS
#define VMAP_SEG_STEPS 64
#endif
constexpr uint32_t kSegSteps = VMAP_SEG_STEPS;    // DDA steps per ray segment (one thread walks one segment)

// A ray is cut into n_seg = max(1, steps / kSegSteps) segments of equal length in the ray
// parameter t; segment j owns the DDA steps whose crossing parameter lies in [j, j+1) * dtau.
struct RayRec {               // per cast ray (index = position in the ordered work list)
  double dt0, dt1, dt2;       // tDelta
  double dlen;                // (double)length
  double dtau;                // segment length in t
  uint32_t e01, e2s;          // end key: e0 | e1 << 16 ; e2 | step codes << 16 (2 bits per axis: 0, 1 = +1, 2 = -1)
  uint32_t n_seg;
  uint32_t point;             // index of the point / pixel
  uint32_t n_keys;            // keys the walk emits when neither the end key nor KeyRay's capacity stops it
                              // first (origin included); 0xFFFFFFFF: not known, the walker tests min(tMax) > length
  uint32_t pad;
};
struct Checkpoint {           // state of the walk at the start of a segment
  double t0, t1, t2;          // tMax
  uint32_t c01, c2;           // current key
  uint32_t count;             // keys emitted by the earlier segments; 0xFFFFFFFF: nothing to walk
  uint32_t pad;
};

__device__ __forceinline__ uint32_t ray_segments(uint32_t steps) {
  return min(max(steps / kSegSteps, 1u), (uint32_t)(kLenBins - 1));
}
// The ordered ray list is sorted by FINE bin, descending: primary key the segment count (so that
// level j of the walker's work slots is a prefix of the list), secondary key -- for rays of fewer
// than kFineSegs segments -- the steps one segment takes, in kFineSub classes, so that the 32 lanes
// of a warp walk segments of nearly the same length.  MEASURED (profiles/r02_ab_walk3.json, 16
// classes): the walk got 8 % SLOWER and k_resolve 25 % slower -- lanes that are balanced but no
// longer neighbours in the scan spread a warp's atomics over more mask words; what bounds the walker
// is atomic traffic, not idle lanes.  Hence one class (= order by segment count only).
#ifndef VMAP_FINE_SUB
#define VMAP_FINE_SUB 1
#endif
constexpr int kFineSub = VMAP_FINE_SUB;
constexpr int kFineSegs = 128;                                   // segment counts below this get sub-bins
constexpr int kFineBins = kFineSegs * kFineSub + (kLenBins - kFineSegs);
__device__ __forceinline__ uint32_t fine_bin_top(uint32_t n_seg) {  // highest fine bin of a segment count
  return n_seg < (uint32_t)kFineSegs ? n_seg * kFineSub + (kFineSub - 1) : (uint32_t)(kFineSegs * kFineSub) + (n_seg - kFineSegs);
}
__device__ __forceinline__ uint32_t fine_bin(uint32_t steps) {
  const uint32_t n_seg = ray_segments(steps);
  if (n_seg >= (uint32_t)kFineSegs) return (uint32_t)(kFineSegs * kFineSub) + (n_seg - kFineSegs);
  // one segment: 1 .. 2 * kSegSteps - 1 steps; several: kSegSteps .. < 2 * kSegSteps steps each
  const uint32_t sub = n_seg == 1u ? min(steps, 2u * kSegSteps - 1u) / (2u * kSegSteps / kFineSub)
                                   : min((steps / n_seg - kSegSteps) / (kSegSteps / kFineSub), (uint32_t)(kFineSub - 1));
  return n_seg * kFineSub + sub;
}

// K2a body.  computeRayKeys' preamble, then the state of the walk at every segment boundary by
// exact jump-ahead of the three tMax sequences (vmap_seq.cuh) -- no DDA stepping here.
// `put(j, cp)` stores the checkpoint of level j and returns false when it has no room.
template <class Put>
__device__ __forceinline__ void ray_checkpoints(const DevParams& P, float ox, float oy, float oz, float ex, float ey,
                                                float ez, uint32_t steps, uint32_t point, RayRec* rec, Put&& put) {
  RayWalk w;
  w.init(P, ox, oy, oz, ex, ey, ez);
  const uint32_t n_seg = ray_segments(steps);
  RayRec rr;
  rr.dt0 = w.dt0; rr.dt1 = w.dt1; rr.dt2 = w.dt2;
  rr.dlen = w.dlen;
  rr.dtau = n_seg > 1 ? __ddiv_rn(__dmul_rn(w.dlen, (double)kSegSteps), (double)steps) : 0.0;
  const uint32_t sc = (uint32_t)(w.s0 > 0 ? 1 : (w.s0 < 0 ? 2 : 0)) | ((uint32_t)(w.s1 > 0 ? 1 : (w.s1 < 0 ? 2 : 0)) << 2) |
                      ((uint32_t)(w.s2 > 0 ? 1 : (w.s2 < 0 ? 2 : 0)) << 4);
  rr.e01 = w.e0 | (w.e1 << 16);
  rr.e2s = w.e2 | (sc << 16);
  rr.n_seg = n_seg;
  rr.point = point;
  rr.n_keys = 0xFFFFFFFFu;
  rr.pad = 0;
  *rec = rr;
  uint32_t n0 = 0, n1 = 0, n2 = 0;
  double t0 = w.t0, t1 = w.t1, t2 = w.t2;
  for (uint32_t j = 0; j < n_seg; j++) {
    if (j > 0) {
      const double tau = __dmul_rn((double)j, rr.dtau);
      const unsigned long long cap = 1ull << 40;
      if (w.s0) n0 += (uint32_t)vmapseq::seq_advance(&t0, w.dt0, tau, cap);
      if (w.s1) n1 += (uint32_t)vmapseq::seq_advance(&t1, w.dt1, tau, cap);
      if (w.s2) n2 += (uint32_t)vmapseq::seq_advance(&t2, w.dt2, tau, cap);
    }
    Checkpoint cp;
    cp.t0 = t0; cp.t1 = t1; cp.t2 = t2;
    const uint32_t c0 = (w.c0 + (uint32_t)(w.s0 * (int)n0)) & 0xFFFFu;
    const uint32_t c1 = (w.c1 + (uint32_t)(w.s1 * (int)n1)) & 0xFFFFu;
    const uint32_t c2 = (w.c2 + (uint32_t)(w.s2 * (int)n2)) & 0xFFFFu;
    cp.c01 = c0 | (c1 << 16);
    cp.c2 = c2;
    cp.count = w.live ? 1u + n0 + n1 + n2 : 0xFFFFFFFFu;
    cp.pad = 0;
    if (!put(j, cp)) return;
  }
  // computeRayKeys stops after the first step that leaves min(tMax) > length: after k steps min(tMax)
  // is the (k + 1)-th smallest crossing parameter, so the walk takes max(L, 1) steps, L = crossings
  // <= length over the three axes, and emits the origin plus every step but the last: max(L, 1) keys
  // (fewer when it runs into the end key or KeyRay's capacity first -- the walker still tests those).
  if (w.live && w.dlen > 0.0 && w.dlen < 1e300) {
    const double past = vmapseq::seq_from_bits(vmapseq::seq_bits(w.dlen) + 1ull);   // t < past  <=>  t <= length
    const unsigned long long cap = 1ull << 40;
    unsigned long long L = (unsigned long long)n0 + n1 + n2;
    if (w.s0) L += vmapseq::seq_advance(&t0, w.dt0, past, cap);
    if (w.s1) L += vmapseq::seq_advance(&t1, w.dt1, past, cap);
    if (w.s2) L += vmapseq::seq_advance(&t2, w.dt2, past, cap);
    if (L < 1ull) L = 1ull;
    if (L < 0xFFFFFFFFull) rec->n_keys = (uint32_t)L;
  }
}

// Every voxel of a ray lies between the segment's first voxel and the ray's end voxel (each key
// coordinate moves monotonically; one voxel of slack for a walk that misses the end voxel): one
// window test per segment instead of one per step.
__device__ __forceinline__ bool segment_inside_window(const MarkSet& ms, uint32_t c0, uint32_t c1, uint32_t c2,
                                                      uint32_t e0, uint32_t e1, uint32_t e2) {
  const int lo0 = (int)min(c0, e0) - 1, hi0 = (int)max(c0, e0) + 1;
  const int lo1 = (int)min(c1, e1) - 1, hi1 = (int)max(c1, e1) + 1;
  const int lo2 = (int)min(c2, e2) - 1, hi2 = (int)max(c2, e2) + 1;
  return ((lo0 >> 3) >= ms.x0) && ((hi0 >> 3) < ms.x0 + (int)ms.nx) && ((lo1 >> 3) >= ms.y0) &&
         ((hi1 >> 3) < ms.y0 + (int)ms.ny) && ((lo2 >> 3) >= ms.z0) && ((hi2 >> 3) < ms.z0 + (int)ms.nz);
}

// K2b body, block addressing (the default walker).  One (segment level j, ray) work slot: performs
// the DDA steps whose crossing parameter is in [j, j+1) * dtau and emits the voxels those steps
// enter (segment 0 also the origin voxel); only the last segment can see computeRayKeys'
// termination tests fire.  Marks go to the dense window.
template <bool kFilter>
__device__ __forceinline__ void walk_slot(const DevParams& P, const MarkSet& ms, Counters* cnt, const Checkpoint& cp,
                                          const RayRec& rr, uint32_t j, float ox, float oy, float oz, int variant,
                                          uint32_t* emitted_io, uint32_t* longest_io) {
  uint32_t emitted = *emitted_io, longest = *longest_io;
  RayWalk w;
  w.c0 = cp.c01 & 0xFFFFu; w.c1 = cp.c01 >> 16; w.c2 = cp.c2;
  w.e0 = rr.e01 & 0xFFFFu; w.e1 = rr.e01 >> 16; w.e2 = rr.e2s & 0xFFFFu;
  const uint32_t sc = rr.e2s >> 16;
  w.s0 = (sc & 3u) == 1u ? 1 : ((sc & 3u) == 2u ? -1 : 0);
  w.s1 = ((sc >> 2) & 3u) == 1u ? 1 : (((sc >> 2) & 3u) == 2u ? -1 : 0);
  w.s2 = ((sc >> 4) & 3u) == 1u ? 1 : (((sc >> 4) & 3u) == 2u ? -1 : 0);
  w.t0 = cp.t0; w.t1 = cp.t1; w.t2 = cp.t2;
  w.dt0 = rr.dt0; w.dt1 = rr.dt1; w.dt2 = rr.dt2;
  w.dlen = rr.dlen;
  w.live = true;
  w.pick();
  const bool last = (j + 1 == rr.n_seg);
  const double tau_end = __dmul_rn((double)(j + 1), rr.dtau);
  uint32_t count = cp.count;
  const uint32_t count0 = count;
  // Every voxel of the ray lies between the segment's first voxel and the ray's end voxel
  // (each key coordinate moves monotonically; one voxel of slack for a walk that misses the
  // end voxel): one window test per segment instead of one per step.
  {
    const int lo0 = (int)min(w.c0, w.e0) - 1, hi0 = (int)max(w.c0, w.e0) + 1;
    const int lo1 = (int)min(w.c1, w.e1) - 1, hi1 = (int)max(w.c1, w.e1) + 1;
    const int lo2 = (int)min(w.c2, w.e2) - 1, hi2 = (int)max(w.c2, w.e2) + 1;
    const bool inside = ((lo0 >> 3) >= ms.x0) && ((hi0 >> 3) < ms.x0 + (int)ms.nx) && ((lo1 >> 3) >= ms.y0) &&
                        ((hi1 >> 3) < ms.y0 + (int)ms.ny) && ((lo2 >> 3) >= ms.z0) && ((hi2 >> 3) < ms.z0 + (int)ms.nz);
    if (!inside) {   // the host re-runs the scan in table mode
      atomicExch(&cnt->overflow_window, 1u);
      return;
    }
  }
  // cur: mask word in flight (window block * 16 + word), bits: voxels collected for it,
  // seen: its value when the ray entered it -- near the sensor every ray sets the same voxels,
  // so there a RED is only issued for bits not visible yet (a stale value costs a redundant
  // RED, never a lost bit).  The block's touched bit is tested before it is set for the same
  // reason.
  uint32_t cur = 0xFFFFFFFFu, bits = 0, seen = 0;
  const uint32_t near_limit = (variant == 0) ? 0xFFFFFFFFu : (variant == 2 ? 0u : (uint32_t)variant);
  auto emit = [&]() {
    if (kFilter && !free_filter_pass(P, ox, oy, oz, w.c0, w.c1, w.c2)) return;
    const uint32_t idx = (((w.c2 >> 3) - (uint32_t)ms.z0) * ms.ny + ((w.c1 >> 3) - (uint32_t)ms.y0)) * ms.nx +
                         ((w.c0 >> 3) - (uint32_t)ms.x0);
    const uint32_t l = local_index(w.c0, w.c1, w.c2);
    const uint32_t wi = (idx << 4) | (l >> 5);
    if (wi != cur) {
      if (bits & ~seen) atomicOr(ms.free_mask + cur, bits);
      if ((wi >> 4) != (cur >> 4)) {
        uint32_t* tb = ms.touched_bits + (idx >> 5);
        if (!((*tb >> (idx & 31u)) & 1u)) atomicOr(tb, 1u << (idx & 31u));
      }
      bits = 0;
      cur = wi;
      seen = (count < near_limit) ? ms.free_mask[wi] : 0u;
    }
    bits |= 1u << (l & 31u);
  };
  if (j == 0) emit();                 // the origin voxel belongs to segment 0
  if (last) {
    for (;;) {
      if (!w.advance()) break;                     // cur == end, or min(tMax) > length
      if (count >= kKeyRayMax - 1u) break;         // KeyRay capacity
      ++count;
      emit();
    }
  } else {
    // steps whose crossing parameter is below tau_end; the termination tests cannot fire here
    while (w.tmin() < tau_end) {
      w.advance_inner();
      if (count >= kKeyRayMax - 1u) break;
      ++count;
      emit();
    }
  }
  if (bits & ~seen) atomicOr(ms.free_mask + cur, bits);
  emitted += count - count0 + (j == 0 ? 1u : 0u);
  if (last) longest = max(longest, count);
  *emitted_io = emitted;
  *longest_io = longest;
}

// ---------------------------------------------------------------------------------------------
// K2b body, dilated addressing.  Needs a window whose x and y extents (in blocks) are powers of
// two: the bit address of a voxel in the window's free-mask array is then the concatenation
//      [ rz | ry : py | rx : px | lz : 3 | ly : 3 | lx : 3 ]         (r* = block, l* = voxel in block)
// and splits into one word per axis (A0 holds rx and lx, A1 ry and ly, A2 rz and lz) with the
// other axes' bits zero.  A step of +-1 voxel along an axis is   A = (A + K) & M   with M the
// axis' bit mask and K = -M (carry runs through the gaps) or -1 (borrow runs through them), so
// the DDA never touches the 16-bit keys again: address = A0 | A1 | A2, mask word = address >> 5,
// bit = address & 31, block = address >> 9, and "reached the end voxel" is one compare of
// addresses.  Marks are identical to walk_slot's.
// ---------------------------------------------------------------------------------------------
template <bool B> struct BoolTag { static constexpr bool value = B; };

struct DilatedWindow {
  uint32_t m0, m1, m2;   // bit masks of the three axes
  uint32_t sh1, sh2;     // position of the ry / rz fields
};
__host__ __device__ inline bool dilated_window_fits(uint32_t nx, uint32_t ny, uint32_t nz) {
  if (!nx || !ny || !nz || (nx & (nx - 1u)) || (ny & (ny - 1u))) return false;
  uint32_t px = 0, py = 0, pz = 0;
  while ((1u << px) < nx) px++;
  while ((1u << py) < ny) py++;
  while ((1u << pz) < nz) pz++;
  return 9u + px + py + pz <= 31u;
}
__host__ __device__ __forceinline__ DilatedWindow dilated_window(const MarkSet& ms) {
  uint32_t px = 0, py = 0;
  while ((1u << px) < ms.nx) px++;
  while ((1u << py) < ms.ny) py++;
  DilatedWindow d;
  d.sh1 = 9u + px;
  d.sh2 = 9u + px + py;
  d.m0 = ((ms.nx - 1u) << 9) | 0x007u;
  d.m1 = ((ms.ny - 1u) << d.sh1) | 0x038u;
  d.m2 = (0xFFFFFFFFu << d.sh2) | 0x1C0u;
  return d;
}
// rel = voxel coordinate relative to the window's first voxel
__device__ __forceinline__ uint32_t dilate_axis(uint32_t rel, uint32_t local_shift, uint32_t block_shift) {
  return ((rel & 7u) << local_shift) | ((rel >> 3) << block_shift);
}

// (t0 < tau) || (t1 < tau) || (t2 < tau)  ==  min(tMax) < tau, as three chained compares: the
// compiler would otherwise build fmin(t0, t1, t2) with its NaN fix-ups (some 18 instructions).
__device__ __forceinline__ bool any_below(double t0, double t1, double t2, double tau) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.lt.f64 p, %1, %4;\n\t"
      "setp.lt.or.f64 p, %2, %4, p;\n\t"
      "setp.lt.or.f64 p, %3, %4, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "d"(t0), "d"(t1), "d"(t2), "d"(tau));
  return r != 0u;
#else
  return (t0 < tau) || (t1 < tau) || (t2 < tau);
#endif
}
// (t0 > len) && (t1 > len) && (t2 > len)  ==  min(tMax) > len
__device__ __forceinline__ bool all_above(double t0, double t1, double t2, double len) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f64 p, %1, %4;\n\t"
      "setp.gt.and.f64 p, %2, %4, p;\n\t"
      "setp.gt.and.f64 p, %3, %4, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "d"(t0), "d"(t1), "d"(t2), "d"(len));
  return r != 0u;
#else
  return (t0 > len) && (t1 > len) && (t2 > len);
#endif
}

// Near-field combining (experimental, VMAP_WALK_ADDR=5): the first segments of all rays cross the
// same few thousand mask words around the sensor, and the same-address RED.OR traffic on them is
// what the walker waits for (profiles/README.md).  A CTA that works on a chunk of first-segment
// slots therefore collects the marks that fall into a cube of 8^3 blocks around the sensor's block
// in a bit grid of its own (shared memory on the device) and writes each non-zero word to the
// window once per chunk, after testing it against what is already set.
constexpr uint32_t kNearBlocks = 8;                                        // blocks per side of the cube
constexpr uint32_t kNearWords = kNearBlocks * kNearBlocks * kNearBlocks * kMaskWords;   // 8192 words = 32 KiB
struct NearGrid {
  uint32_t* words;            // [kNearWords], zero between chunks
  uint32_t bx0, by0, bz0;     // first block of the cube, window-relative block coordinates
};
// cube of blocks centred on the block of voxel (c0, c1, c2); false when it leaves the window
__device__ __forceinline__ bool near_grid_place(const MarkSet& ms, uint32_t c0, uint32_t c1, uint32_t c2, NearGrid* g) {
  const int bx = (int)(c0 >> 3) - ms.x0 - (int)(kNearBlocks / 2), by = (int)(c1 >> 3) - ms.y0 - (int)(kNearBlocks / 2),
            bz = (int)(c2 >> 3) - ms.z0 - (int)(kNearBlocks / 2);
  if (bx < 0 || by < 0 || bz < 0 || bx + (int)kNearBlocks > (int)ms.nx || by + (int)kNearBlocks > (int)ms.ny ||
      bz + (int)kNearBlocks > (int)ms.nz)
    return false;
  g->bx0 = (uint32_t)bx; g->by0 = (uint32_t)by; g->bz0 = (uint32_t)bz;
  return true;
}
// word i of the grid -> window (one caller per word, after every slot of the chunk is done)
__device__ __forceinline__ void near_flush_word(const MarkSet& ms, const NearGrid& g, uint32_t i) {
  const uint32_t bits = g.words[i];
  if (!bits) return;
  g.words[i] = 0u;
  const uint32_t b = i >> 4;
  const uint32_t idx = ((g.bz0 + (b >> 6)) * ms.ny + (g.by0 + ((b >> 3) & 7u))) * ms.nx + (g.bx0 + (b & 7u));
  uint32_t* w = ms.free_mask + ((size_t)idx * kMaskWords + (i & 15u));
  if (bits & ~*w) atomicOr(w, bits);
  uint32_t* tb = ms.touched_bits + (idx >> 5);
  const uint32_t bit = 1u << (idx & 31u);
  if (!(*tb & bit)) atomicOr(tb, bit);
}

// kMode 0: the block's touched bit is tested when the ray enters the block.
// kMode 1: the touched word is only LOADED at block entry and tested when the ray leaves the
//          block (or ends), so that the walk never waits for that load.
// kNear: marks inside `near`'s cube go to the CTA's own bit grid (see NearGrid) instead of the window.
// kWide: marks are collected per 64-voxel pair of mask words (one 8 x 8 z-slice of the block) and
//        flushed with one 64-bit RED.OR; a mostly horizontal ray stays in a slice two to three times as
//        long as in a 32-voxel word, and the number of REDs is what bounds the walker.
template <bool kWide> struct MaskWord { typedef uint32_t type; };
template <> struct MaskWord<true> { typedef unsigned long long type; };

template <bool kFilter, int kMode = 0, bool kNear = false, bool kWide = false>
__device__ __forceinline__ void walk_slot_dilated(const DevParams& P, const MarkSet& ms, const DilatedWindow& dw,
                                                  Counters* cnt, const Checkpoint& cp, const RayRec& rr, uint32_t j,
                                                  float ox, float oy, float oz, uint32_t near_limit,
                                                  uint32_t* emitted, uint32_t* longest, const NearGrid& near = NearGrid(),
                                                  uint32_t n_inner = 0xFFFFFFFFu, bool far_blind = false) {
  if (cp.count == 0xFFFFFFFFu) return;
  uint32_t c0 = cp.c01 & 0xFFFFu, c1 = cp.c01 >> 16, c2 = cp.c2;
  const uint32_t e0 = rr.e01 & 0xFFFFu, e1 = rr.e01 >> 16, e2 = rr.e2s & 0xFFFFu;
  if (!segment_inside_window(ms, c0, c1, c2, e0, e1, e2)) {   // the host re-runs the scan in table mode
    atomicExch(&cnt->overflow_window, 1u);
    return;
  }
  const uint32_t sc = rr.e2s >> 16;
  // step codes: 0 = the ray does not move along the axis (its tMax is DBL_MAX), 1 = +1, 2 = -1
  const uint32_t sc0 = sc & 3u, sc1 = (sc >> 2) & 3u, sc2 = (sc >> 4) & 3u;
  const uint32_t k0 = sc0 == 1u ? (0u - dw.m0) : (sc0 == 2u ? 0xFFFFFFFFu : 0u);
  const uint32_t k1 = sc1 == 1u ? (0u - dw.m1) : (sc1 == 2u ? 0xFFFFFFFFu : 0u);
  const uint32_t k2 = sc2 == 1u ? (0u - dw.m2) : (sc2 == 2u ? 0xFFFFFFFFu : 0u);
  const uint32_t wx = (uint32_t)ms.x0 << 3, wy = (uint32_t)ms.y0 << 3, wz = (uint32_t)ms.z0 << 3;
  uint32_t a0 = dilate_axis(c0 - wx, 0u, 9u), a1 = dilate_axis(c1 - wy, 3u, dw.sh1), a2 = dilate_axis(c2 - wz, 6u, dw.sh2);
  const uint32_t a_end = dilate_axis(e0 - wx, 0u, 9u) | dilate_axis(e1 - wy, 3u, dw.sh1) | dilate_axis(e2 - wz, 6u, dw.sh2);
  // key steps (only the free-space filter still needs the keys)
  const int s0 = sc0 == 1u ? 1 : (sc0 == 2u ? -1 : 0), s1 = sc1 == 1u ? 1 : (sc1 == 2u ? -1 : 0),
            s2 = sc2 == 1u ? 1 : (sc2 == 2u ? -1 : 0);
  double t0 = cp.t0, t1 = cp.t1, t2 = cp.t2;
  const double dt0 = rr.dt0, dt1 = rr.dt1, dt2 = rr.dt2;
  const bool last = (j + 1 == rr.n_seg);
  uint32_t count = cp.count;
  const uint32_t count0 = count;
  // cur / bits / seen: as in walk_slot (kWide: cur counts 64-bit words)
  typedef typename MaskWord<kWide>::type Word;
  constexpr uint32_t kWordShift = kWide ? 6u : 5u;             // address bits below the word index
  constexpr uint32_t kWordsPerBlock = 512u >> kWordShift;
  uint32_t cur = 0xFFFFFFFFu;
  Word bits = 0, seen = 0;
  // kMode 1: touched word of the block in flight as loaded at entry (all ones: nothing pending)
  uint32_t tb_val = 0xFFFFFFFFu, tb_blk = 0;
  // kNear: base word of the current block in the near grid, or kNotFound outside the cube
  uint32_t near_base = kNotFound;
  auto flush = [&]() {
    if (kNear && near_base != kNotFound) {
      if (kWide) {
        const uint32_t lo = (uint32_t)bits, hi = (uint32_t)((unsigned long long)bits >> 32);
        if (lo) atomicOr(near.words + (near_base | ((cur & 7u) << 1)), lo);
        if (hi) atomicOr(near.words + (near_base | ((cur & 7u) << 1) | 1u), hi);
      } else if (bits) {
        atomicOr(near.words + (near_base | (cur & 15u)), (uint32_t)bits);
      }
    } else if (bits & ~seen) {
      atomicOr(reinterpret_cast<Word*>(ms.free_mask) + cur, bits);
    }
  };
  auto emit = [&]() {
    if (kFilter && !free_filter_pass(P, ox, oy, oz, c0, c1, c2)) return;
    const uint32_t a = a0 | a1 | a2;
    const uint32_t wi = a >> kWordShift;
    if (wi != cur) {
      flush();
      if ((wi ^ cur) >= kWordsPerBlock) {            // another block
        if (kNear) {
          const uint32_t dx = (a0 >> 9) - near.bx0, dy = (a1 >> dw.sh1) - near.by0, dz = (a2 >> dw.sh2) - near.bz0;
          near_base = (dx < kNearBlocks && dy < kNearBlocks && dz < kNearBlocks)
                          ? ((((dz * kNearBlocks) + dy) * kNearBlocks + dx) << 4) : kNotFound;
        }
        if (kNear && near_base != kNotFound) {
          // touched bit: set when the grid is flushed
        } else if (kMode == 0) {
          uint32_t* tb = ms.touched_bits + (a >> 14);
          const uint32_t bit = 1u << ((a >> 9) & 31u);
          // far_blind (VMAP_WALK_FAR_BLIND=1, experiment): beyond the near field the RED goes out without
          // looking first.  MEASURED: the walk takes TWICE as long (0.19 -> 0.40 ms, profiles/
          // r02_ab_walk3.json) -- 1.7 M more REDs on a few thousand touched words; same-address atomics
          // are what the walker waits for, a load that filters them is cheap in comparison.
          if (far_blind && count >= near_limit) atomicOr(tb, bit);
          else if (!(*tb & bit)) atomicOr(tb, bit);
        } else {
          if (!((tb_val >> (tb_blk & 31u)) & 1u)) atomicOr(ms.touched_bits + (tb_blk >> 5), 1u << (tb_blk & 31u));
          tb_blk = a >> 9;
          tb_val = ms.touched_bits[tb_blk >> 5];
        }
      }
      bits = 0;
      cur = wi;
      seen = (!(kNear && near_base != kNotFound) && count < near_limit) ? reinterpret_cast<const Word*>(ms.free_mask)[wi] : (Word)0;
    }
    bits |= (Word)1 << (a & ((1u << kWordShift) - 1u));
  };
  // one DDA step: strict '<' comparisons, ties go to the later axis (computeRayKeys)
  auto step = [&]() {
    const bool a01 = t0 < t1, a02 = t0 < t2, a12 = t1 < t2;
    const bool d0 = a01 && a02, d1 = !a01 && a12;
    if (d0) {
      a0 = (a0 + k0) & dw.m0;
      t0 = __dadd_rn(t0, dt0);
      if (kFilter) c0 = (c0 + (uint32_t)s0) & 0xFFFFu;
    } else if (d1) {
      a1 = (a1 + k1) & dw.m1;
      t1 = __dadd_rn(t1, dt1);
      if (kFilter) c1 = (c1 + (uint32_t)s1) & 0xFFFFu;
    } else {
      a2 = (a2 + k2) & dw.m2;
      t2 = __dadd_rn(t2, dt2);
      if (kFilter) c2 = (c2 + (uint32_t)s2) & 0xFFFFu;
    }
  };
  if (j == 0) emit();                 // the origin voxel belongs to segment 0
  const double dlen = rr.dlen;
  const double tau_end = __dmul_rn((double)(j + 1), rr.dtau);
  auto run = [&](auto capped) {
    if (last && n_inner != 0xFFFFFFFFu) {
      // counted: RayRec::n_keys bounds the keys of the whole walk, so min(tMax) is never compared
      for (uint32_t k = n_inner; k; --k) {
        step();
        if ((a0 | a1 | a2) == a_end) break;                            // cur == end
        if (decltype(capped)::value && count >= kKeyRayMax - 1u) break;   // KeyRay capacity
        ++count;
        emit();
      }
    } else if (last) {
      for (;;) {
        step();
        if ((a0 | a1 | a2) == a_end) break;                            // cur == end
        if (all_above(t0, t1, t2, dlen)) break;                        // min(tMax) > length
        if (decltype(capped)::value && count >= kKeyRayMax - 1u) break;   // KeyRay capacity
        ++count;
        emit();
      }
    } else if (n_inner != 0xFFFFFFFFu) {
      // the same steps, counted: the next level's checkpoint says how many keys the walk has emitted
      // when it gets there (ray_checkpoints: count = 1 + steps), so no crossing parameter is compared
      for (uint32_t k = n_inner; k; --k) {
        step();
        if (decltype(capped)::value && count >= kKeyRayMax - 1u) break;
        ++count;
        emit();
      }
    } else {
      // steps whose crossing parameter is below tau_end; the termination tests cannot fire here
      while (any_below(t0, t1, t2, tau_end)) {
        step();
        if (decltype(capped)::value && count >= kKeyRayMax - 1u) break;
        ++count;
        emit();
      }
    }
  };
  // Only rays that could reach KeyRay's capacity carry the per-step test: a walk takes at most
  // (Manhattan distance of the keys) + 3 steps, and n_seg = steps / kSegSteps (saturating at
  // kLenBins - 1, which is far above the capacity).
  if ((rr.n_seg + 1u) * kSegSteps + 4u >= kKeyRayMax - 1u) run(BoolTag<true>());
  else run(BoolTag<false>());
  flush();
  if (kMode == 1 && !((tb_val >> (tb_blk & 31u)) & 1u)) atomicOr(ms.touched_bits + (tb_blk >> 5), 1u << (tb_blk & 31u));
  *emitted += count - count0 + (j == 0 ? 1u : 0u);
  if (last) *longest = max(*longest, count);
}

// ---------------------------------------------------------------------------------------------
// K2b body, dilated addressing, QUEUED marks (the walker that ships).  What costs issue slots in
// walk_slot_dilated is not the DDA step but the code behind "the ray has entered another mask word":
// with 32 lanes at different places of their walks some lane takes that branch in nearly every
// iteration, so the whole warp pays for the flush (RED, block test, touched load + test + RED, near
// preload) at a handful of active lanes, every step.  Here the branch only PARKS the finished word
// (word index, 64 bits) in a small per-lane queue in shared memory; after every kQueueSteps steps the
// warp drains the queues together -- all lanes issue their loads, tests and REDs side by side, one
// queue entry per pass.  Needs the step count of the slot up front (counted loops: the next level's
// checkpoint / RayRec::n_keys), so that all lanes of a warp run the same loop nest.
// Marks are the ones walk_slot_dilated makes (OR is commutative; every parked word's block gets its
// touched bit); the near-field test "are these bits set already" is done at drain time.
// WARP-COLLECTIVE: every lane of the warp calls it (`active` false: no slot); q: this lane's queue.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kQueueSteps = 8;                 // steps between drains = queue capacity
#if !defined(__CUDA_ARCH__) && defined(VMAP_HOST_EMULATION)
static inline bool vmap_warp_any(bool p) { return p; }      // one "lane" at a time on the host
#else
__device__ __forceinline__ bool vmap_warp_any(bool p) { return __any_sync(0xFFFFFFFFu, p) != 0; }
#endif

// q: this lane's queue; entry e, word k (0: index of the 64-bit mask word, 1: bits 0..31, 2: bits 32..63)
// at q[(3 * e + k) * kQStride] (kQStride = threads per CTA on the device: conflict-free banks).
template <bool kFilter, uint32_t kQStride, bool kBranchFree = false>
__device__ __forceinline__ void walk_slot_queued(const DevParams& P, const MarkSet& ms, const DilatedWindow& dw, Counters* cnt,
                                                 const Checkpoint& cp, const RayRec& rr, uint32_t j, float ox, float oy,
                                                 float oz, uint32_t near_limit, uint32_t* emitted, uint32_t* longest,
                                                 uint32_t n_steps, bool active, uint32_t* q) {
  uint32_t c0 = 0, c1 = 0, c2 = 0, a0 = 0, a1 = 0, a2 = 0, k0 = 0, k1 = 0, k2 = 0;
  uint32_t a_stop = 0xFFFFFFFFu;                    // address of the end voxel in a ray's last segment, else none
  int s0 = 0, s1 = 0, s2 = 0;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, dt0 = 0.0, dt1 = 0.0, dt2 = 0.0;
  bool last = false, capped = false;
  uint32_t count = 0, count0 = 0;
  if (active && cp.count == 0xFFFFFFFFu) active = false;
  if (active) {
    c0 = cp.c01 & 0xFFFFu; c1 = cp.c01 >> 16; c2 = cp.c2;
    const uint32_t e0 = rr.e01 & 0xFFFFu, e1 = rr.e01 >> 16, e2 = rr.e2s & 0xFFFFu;
    if (!segment_inside_window(ms, c0, c1, c2, e0, e1, e2)) {   // the host re-runs the scan in table mode
      atomicExch(&cnt->overflow_window, 1u);
      active = false;
    } else {
      const uint32_t sc = rr.e2s >> 16;
      const uint32_t sc0 = sc & 3u, sc1 = (sc >> 2) & 3u, sc2 = (sc >> 4) & 3u;
      k0 = sc0 == 1u ? (0u - dw.m0) : (sc0 == 2u ? 0xFFFFFFFFu : 0u);
      k1 = sc1 == 1u ? (0u - dw.m1) : (sc1 == 2u ? 0xFFFFFFFFu : 0u);
      k2 = sc2 == 1u ? (0u - dw.m2) : (sc2 == 2u ? 0xFFFFFFFFu : 0u);
      s0 = sc0 == 1u ? 1 : (sc0 == 2u ? -1 : 0); s1 = sc1 == 1u ? 1 : (sc1 == 2u ? -1 : 0); s2 = sc2 == 1u ? 1 : (sc2 == 2u ? -1 : 0);
      const uint32_t wx = (uint32_t)ms.x0 << 3, wy = (uint32_t)ms.y0 << 3, wz = (uint32_t)ms.z0 << 3;
      a0 = dilate_axis(c0 - wx, 0u, 9u); a1 = dilate_axis(c1 - wy, 3u, dw.sh1); a2 = dilate_axis(c2 - wz, 6u, dw.sh2);
      t0 = cp.t0; t1 = cp.t1; t2 = cp.t2;
      dt0 = rr.dt0; dt1 = rr.dt1; dt2 = rr.dt2;
      last = (j + 1 == rr.n_seg);
      // (a voxel address has at most 31 bits -- dilated_window_fits -- so all ones never matches)
      if (last) a_stop = dilate_axis(e0 - wx, 0u, 9u) | dilate_axis(e1 - wy, 3u, dw.sh1) | dilate_axis(e2 - wz, 6u, dw.sh2);
      capped = (rr.n_seg + 1u) * kSegSteps + 4u >= kKeyRayMax - 1u;    // only such rays can reach KeyRay's capacity
      count = count0 = cp.count;
    }
  }
  const bool any_capped = vmap_warp_any(capped);    // warp-uniform: the per-step capacity test is compiled out of the common loop
  uint32_t remaining = active ? n_steps : 0u;
  uint32_t cur = 0xFFFFFFFFu, lo = 0u, hi = 0u;      // 64-bit mask word in flight and the bits collected for it
  uint32_t* qp = q;                                  // next free queue entry
  uint32_t last_blk = 0xFFFFFFFFu;
  auto emit = [&](uint32_t a) {
    if (kFilter && !free_filter_pass(P, ox, oy, oz, c0, c1, c2)) return;
    const uint32_t wi = a >> 6;
    if (wi != cur) {
      if (cur != 0xFFFFFFFFu) {                     // park the finished word
        qp[0] = cur;
        qp[kQStride] = lo;
        qp[2u * kQStride] = hi;
        qp += 3u * kQStride;
      }
      lo = 0u; hi = 0u;
      cur = wi;
    }
    const uint32_t bit = 1u << (a & 31u);
    if (a & 32u) hi |= bit; else lo |= bit;
  };
  auto drain = [&]() {
    // near the sensor every ray sets the same voxels: there a RED only goes out for bits not visible
    // yet.  Decided per drain from the lane's key count (the entries are at most kQueueSteps keys old).
    const bool near = count < near_limit + kQueueSteps;
    for (const uint32_t* e = q; vmap_warp_any(e < qp); e += 3u * kQStride) {
      if (e < qp) {
        const uint32_t wi = e[0];
        const unsigned long long b = (unsigned long long)e[kQStride] | ((unsigned long long)e[2u * kQStride] << 32);
        unsigned long long* w = reinterpret_cast<unsigned long long*>(ms.free_mask) + wi;
        if (near) { if (b & ~*w) atomicOr(w, b); }
        else atomicOr(w, b);
        const uint32_t blk = wi >> 3;
        if (blk != last_blk) {
          last_blk = blk;
          uint32_t* tb = ms.touched_bits + (blk >> 5);
          const uint32_t bit = 1u << (blk & 31u);
          if (!(*tb & bit)) atomicOr(tb, bit);
        }
      }
    }
    qp = q;
  };
  if (active && j == 0) emit(a0 | a1 | a2);       // the origin voxel belongs to segment 0
  while (vmap_warp_any(remaining != 0u)) {
#pragma unroll 2
    for (uint32_t s = 0; s < kQueueSteps; s++) {
      if (kBranchFree) {
        // One DDA step written without branches (selects become predicated instructions; the only branch
        // left in a step is "the voxel lies in another mask word").  MEASURED against the branchy form
        // below: see profiles/README.md (round 2, "branch-free step").
        const bool act = remaining != 0u;
        const bool a01 = t0 < t1, a02 = t0 < t2, a12 = t1 < t2;
        const bool d0 = act && a01 && a02;
        const bool d1 = act && !a01 && a12;
        const bool d2 = act && !(a01 && a02) && !(!a01 && a12);
        a0 = d0 ? ((a0 + k0) & dw.m0) : a0;
        a1 = d1 ? ((a1 + k1) & dw.m1) : a1;
        a2 = d2 ? ((a2 + k2) & dw.m2) : a2;
        t0 = d0 ? __dadd_rn(t0, dt0) : t0;
        t1 = d1 ? __dadd_rn(t1, dt1) : t1;
        t2 = d2 ? __dadd_rn(t2, dt2) : t2;
        if (kFilter) {
          c0 = d0 ? ((c0 + (uint32_t)s0) & 0xFFFFu) : c0;
          c1 = d1 ? ((c1 + (uint32_t)s1) & 0xFFFFu) : c1;
          c2 = d2 ? ((c2 + (uint32_t)s2) & 0xFFFFu) : c2;
        }
        const uint32_t a = a0 | a1 | a2;
        bool go = act && a != a_stop;                                       // cur == end: this voxel is not emitted
        if (any_capped) go = go && !(capped && count >= kKeyRayMax - 1u);   // KeyRay capacity
        remaining = go ? remaining - 1u : 0u;
        count += go ? 1u : 0u;
        if (go) emit(a);
      } else if (remaining) {
        // one DDA step: strict '<' comparisons, ties go to the later axis (computeRayKeys)
        const bool a01 = t0 < t1, a02 = t0 < t2, a12 = t1 < t2;
        const bool d0 = a01 && a02, d1 = !a01 && a12;
        if (d0) {
          a0 = (a0 + k0) & dw.m0;
          t0 = __dadd_rn(t0, dt0);
          if (kFilter) c0 = (c0 + (uint32_t)s0) & 0xFFFFu;
        } else if (d1) {
          a1 = (a1 + k1) & dw.m1;
          t1 = __dadd_rn(t1, dt1);
          if (kFilter) c1 = (c1 + (uint32_t)s1) & 0xFFFFu;
        } else {
          a2 = (a2 + k2) & dw.m2;
          t2 = __dadd_rn(t2, dt2);
          if (kFilter) c2 = (c2 + (uint32_t)s2) & 0xFFFFu;
        }
        const uint32_t a = a0 | a1 | a2;
        bool stop = a == a_stop;                                          // cur == end: this voxel is not emitted
        if (any_capped) stop = stop || (capped && count >= kKeyRayMax - 1u);   // KeyRay capacity
        if (stop) {
          remaining = 0u;
        } else {
          --remaining;
          ++count;
          emit(a);
        }
      }
    }
    drain();
  }
  if (cur != 0xFFFFFFFFu) {
    qp[0] = cur;
    qp[kQStride] = lo;
    qp[2u * kQStride] = hi;
    qp += 3u * kQStride;
  }
  drain();
  if (active) {
    *emitted += count - count0 + (j == 0 ? 1u : 0u);
    if (last) *longest = max(*longest, count);
  }
}

}  // namespace vmapdev
