// Host build of the product's per-thread walker code (vmap_device.cuh / vmap_walk.cuh) for the CPU
// test-suite, driven by tests/test_walk_host_emulation.py.  The checker below marks one scan's
// free voxels three ways and the Python side compares them:
//   walk_plain        walk_ray (the straightforward DDA)      -> marks through window_index()
//   walk_segmented    ray_checkpoints + walk_slot (the default walker) or walk_slot_dilated
//                     (experimental walkers)                 -> marks through block / dilated addresses
// and hands individual rays' keys to Python for comparison with the oracle's computeRayKeys.
#include "../../volumetric_mapping_b200/csrc/vmap_walk.cuh"

#include <algorithm>
#include <vector>

using namespace vmapdev;

namespace {
DevParams make_params(double res, double max_free_space, double min_height_free_space) {
  DevParams P;
  memset(&P, 0, sizeof(P));
  P.res = res;
  P.res_factor = 1.0 / res;
  P.max_range = -1.0;
  P.max_free_space = max_free_space;
  P.min_height_free_space = min_height_free_space;
  P.free_filter = max_free_space != 0.0;
  return P;
}
MarkSet make_window(const int32_t origin_block[3], const uint32_t dims[3], uint32_t* free_mask, uint32_t* touched) {
  MarkSet ms;
  memset(&ms, 0, sizeof(ms));
  ms.free_mask = free_mask;
  ms.touched_bits = touched;
  ms.x0 = origin_block[0]; ms.y0 = origin_block[1]; ms.z0 = origin_block[2];
  ms.nx = dims[0]; ms.ny = dims[1]; ms.nz = dims[2];
  ms.dense = 1;
  return ms;
}
}  // namespace

extern "C" {

// keys of one ray by the plain DDA (k3 out, up to cap keys); returns the key count
uint32_t wc_ray_keys(double res, const float* origin, const float* end, uint16_t* k3, uint32_t cap) {
  const DevParams P = make_params(res, 0.0, 0.0);
  uint32_t n = 0;
  return walk_ray(P, origin[0], origin[1], origin[2], end[0], end[1], end[2], [&](uint32_t a, uint32_t b, uint32_t c) {
    if (n < cap) { k3[3 * n] = (uint16_t)a; k3[3 * n + 1] = (uint16_t)b; k3[3 * n + 2] = (uint16_t)c; }
    n++;
    return true;
  });
}

// number of DDA steps the front-end predicts for a ray (classify_point's `len`)
static uint32_t predicted_steps(const DevParams& P, const float* o, const float* e) {
  const int o0 = scaled_coord(P.res_factor, (double)o[0]), o1 = scaled_coord(P.res_factor, (double)o[1]),
            o2 = scaled_coord(P.res_factor, (double)o[2]);
  const int e0 = scaled_coord(P.res_factor, (double)e[0]), e1 = scaled_coord(P.res_factor, (double)e[1]),
            e2 = scaled_coord(P.res_factor, (double)e[2]);
  const bool walk = key_in_range(o0) && key_in_range(o1) && key_in_range(o2) && key_in_range(e0) && key_in_range(e1) &&
                    key_in_range(e2);
  return walk ? (uint32_t)(abs(e0 - o0) + abs(e1 - o1) + abs(e2 - o2)) : 0u;
}

// Plain marks.  stats: [traversals, longest, overflow_window]
int wc_walk_plain(double res, double max_free_space, double min_height, const float* origin, const float* ends,
                  uint32_t n_rays, const int32_t* origin_block, const uint32_t* dims, uint32_t* free_mask,
                  uint32_t* touched, uint64_t* stats) {
  const DevParams P = make_params(res, max_free_space, min_height);
  const MarkSet ms = make_window(origin_block, dims, free_mask, touched);
  uint64_t trav = 0, longest = 0, ovf = 0;
  for (uint32_t r = 0; r < n_rays; r++) {
    const float* e = ends + 3 * r;
    const uint32_t n = walk_ray(P, origin[0], origin[1], origin[2], e[0], e[1], e[2], [&](uint32_t a, uint32_t b, uint32_t c) {
      if (P.free_filter && !free_filter_pass(P, origin[0], origin[1], origin[2], a, b, c)) return true;
      const uint32_t idx = window_index(ms, a, b, c);
      if (idx == kNotFound) { ovf = 1; return true; }
      const uint32_t l = local_index(a, b, c);
      ms.free_mask[(size_t)idx * kMaskWords + (l >> 5)] |= 1u << (l & 31u);
      ms.touched_bits[idx >> 5] |= 1u << (idx & 31u);
      return true;
    });
    trav += n;
    longest = std::max<uint64_t>(longest, n);
  }
  stats[0] = trav; stats[1] = longest; stats[2] = ovf;
  return 0;
}

// Segmented marks: checkpoints for every ray, then every (level, ray) slot on its own, levels
// outermost (the order the kernel's level-major slots are walked in is irrelevant for the marks,
// but this one is the least like the plain walk).  stats: [traversals, longest, overflow_window, slots, near chunks]
int wc_walk_segmented(double res, double max_free_space, double min_height, const float* origin, const float* ends,
                      uint32_t n_rays, const int32_t* origin_block, const uint32_t* dims, uint32_t* free_mask,
                      uint32_t* touched, uint32_t near_limit, int mode, uint64_t* stats) {
  const DevParams P = make_params(res, max_free_space, min_height);
  const MarkSet ms = make_window(origin_block, dims, free_mask, touched);
  if (!dilated_window_fits(ms.nx, ms.ny, ms.nz)) return 1;
  const DilatedWindow dw = dilated_window(ms);
  Counters cnt;
  memset(&cnt, 0, sizeof(cnt));
  std::vector<RayRec> recs(n_rays);
  std::vector<std::vector<Checkpoint> > cps(n_rays);
  uint32_t max_seg = 0;
  for (uint32_t r = 0; r < n_rays; r++) {
    const float* e = ends + 3 * r;
    ray_checkpoints(P, origin[0], origin[1], origin[2], e[0], e[1], e[2], predicted_steps(P, origin, e), r, &recs[r],
                    [&](uint32_t, const Checkpoint& cp) { cps[r].push_back(cp); return true; });
    max_seg = std::max(max_seg, (uint32_t)cps[r].size());
  }
  uint32_t emitted = 0, longest = 0;
  uint64_t trav = 0, slots = 0;
  if (mode == 2) {
    // The near-field kernel's schedule: rays ordered by segment count (descending), slots
    // level-major, chunks of 256 slots; a chunk made of first-segment slots only collects the
    // marks around the sensor in its own bit grid and flushes it when the chunk is done.
    std::vector<uint32_t> order(n_rays);
    for (uint32_t r = 0; r < n_rays; r++) order[r] = r;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cps[a].size() > cps[b].size(); });
    std::vector<std::pair<uint32_t, uint32_t> > slot_list;   // (level, ray)
    for (uint32_t jj = 0; jj < max_seg; jj++)
      for (uint32_t r : order)
        if (jj < cps[r].size()) slot_list.push_back(std::make_pair(jj, r));
    size_t level0 = 0;
    while (level0 < slot_list.size() && slot_list[level0].first == 0) level0++;
    std::vector<uint32_t> grid(kNearWords, 0u);
    NearGrid near;
    near.words = grid.data();
    const int s0 = scaled_coord(P.res_factor, (double)origin[0]), s1 = scaled_coord(P.res_factor, (double)origin[1]),
              s2 = scaled_coord(P.res_factor, (double)origin[2]);
    const bool near_ok = key_in_range(s0) && key_in_range(s1) && key_in_range(s2) &&
                         near_grid_place(ms, (uint32_t)s0, (uint32_t)s1, (uint32_t)s2, &near);
    uint64_t near_chunks = 0;
    for (size_t c0 = 0; c0 < slot_list.size(); c0 += 256) {
      const bool first_level = near_ok && c0 + 256 <= level0;
      for (size_t t = c0; t < std::min(c0 + 256, slot_list.size()); t++) {
        const uint32_t jj = slot_list[t].first, r = slot_list[t].second;
        emitted = 0;
        if (first_level)
          walk_slot_dilated<false, 0, true>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, near);
        else
          walk_slot_dilated<false, 0, false>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest);
        trav += emitted;
        slots++;
      }
      if (first_level) {
        for (uint32_t k = 0; k < kNearWords; k++) near_flush_word(ms, near, k);
        near_chunks++;
      }
    }
    stats[4] = near_chunks;
  } else
  for (uint32_t jj = max_seg; jj-- > 0;) {       // deepest level first
    for (uint32_t r = 0; r < n_rays; r++) {
      if (jj >= cps[r].size()) continue;
      emitted = 0;
#define WC_SLOT(F, M) walk_slot_dilated<F, M>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest)
      // mode 3: the default walker (block addressing); its `variant` argument encodes the near limit
      const int variant = near_limit == 0xFFFFFFFFu ? 0 : (near_limit == 0u ? 2 : (int)near_limit);
      if (mode == 3) {
        if (P.free_filter) walk_slot<true>(P, ms, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], variant, &emitted, &longest);
        else walk_slot<false>(P, ms, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], variant, &emitted, &longest);
      } else if (mode == 6) {
        // the shipped walker: queued marks (one "lane": the queue is drained after every 8 steps)
        uint32_t n_steps = 0;
        bool active = true;
        if (jj + 1 < cps[r].size()) n_steps = cps[r][jj + 1].count - cps[r][jj].count;
        else if (recs[r].n_keys != 0xFFFFFFFFu) n_steps = recs[r].n_keys > cps[r][jj].count ? recs[r].n_keys - cps[r][jj].count : 0u;
        else active = false;
        uint32_t mq[3 * kQueueSteps];
        if (!active) {
          if (P.free_filter) walk_slot_dilated<true, 0, false, true>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest);
          else walk_slot_dilated<false, 0, false, true>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest);
        } else if (P.free_filter) walk_slot_queued<true, 1u>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, n_steps, true, mq);
        else walk_slot_queued<false, 1u>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, n_steps, true, mq);
      } else if (mode == 4 || mode == 5) {
        // the shipped kernel's counted loop: steps of a non-last segment = next checkpoint's count - this one's
        // ... of the last one = RayRec::n_keys - this checkpoint's count
        uint32_t n_inner = 0xFFFFFFFFu;
        if (jj + 1 < cps[r].size()) n_inner = cps[r][jj + 1].count - cps[r][jj].count;
        else if (recs[r].n_keys != 0xFFFFFFFFu) n_inner = recs[r].n_keys > cps[r][jj].count ? recs[r].n_keys - cps[r][jj].count : 0u;
        // mode 5: marks collected per 64-voxel word pair (kWide), what ships
        if (mode == 5) {
          if (P.free_filter) walk_slot_dilated<true, 0, false, true>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, NearGrid(), n_inner);
          else walk_slot_dilated<false, 0, false, true>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, NearGrid(), n_inner);
        } else if (P.free_filter) walk_slot_dilated<true, 0>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, NearGrid(), n_inner);
        else walk_slot_dilated<false, 0>(P, ms, dw, &cnt, cps[r][jj], recs[r], jj, origin[0], origin[1], origin[2], near_limit, &emitted, &longest, NearGrid(), n_inner);
      } else if (P.free_filter) { if (mode == 1) WC_SLOT(true, 1); else WC_SLOT(true, 0); }
      else { if (mode == 1) WC_SLOT(false, 1); else WC_SLOT(false, 0); }
#undef WC_SLOT
      trav += emitted;
      slots++;
    }
  }
  stats[0] = trav; stats[1] = longest; stats[2] = cnt.overflow_window; stats[3] = slots;
  return 0;
}
}
