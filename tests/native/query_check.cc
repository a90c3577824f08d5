// Host build of the product's bounding-box status query body (vmap_query.cuh) for the CPU
// test-suite, driven by tests/test_query_host_emulation.py.  The map is built in host memory in
// the product's own layout (block hash + pool, through the product's find_or_insert) from leaves
// the oracle exported; each box is then scanned by n_lanes emulated lanes, one after the other.
#include "../../volumetric_mapping_b200/csrc/vmap_query.cuh"

#include <vector>

using namespace vmapdev;

namespace {
struct HostMap {
  std::vector<unsigned long long> keys;
  std::vector<uint32_t> slot;
  std::vector<float> pool;
  Persist persist;
  Counters cnt;
  MapView view;
};
}  // namespace

extern "C" {

void* qc_map_create(const uint16_t* k3, const float* logodds, size_t n) {
  HostMap* hm = new HostMap();
  size_t cap = 1024;
  while (cap < 4 * (n / 8 + 16)) cap <<= 1;     // generous: at most n distinct blocks
  while (cap < 2 * n && cap < (1u << 22)) cap <<= 1;
  hm->keys.assign(cap, kEmpty);
  hm->slot.assign(cap, kUnassigned);
  memset(&hm->persist, 0, sizeof(hm->persist));
  memset(&hm->cnt, 0, sizeof(hm->cnt));
  MapView& m = hm->view;
  memset(&m, 0, sizeof(m));
  m.keys = hm->keys.data();
  m.slot = hm->slot.data();
  m.persist = &hm->persist;
  m.cap_mask = (uint32_t)cap - 1u;
  m.epoch = 1;
  // pass 1: blocks
  for (size_t i = 0; i < n; i++) {
    const uint32_t h = find_or_insert(m, &hm->cnt, block_key(k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]));
    if (h == kNotFound) { delete hm; return nullptr; }
    if (m.slot[h] == kUnassigned) m.slot[h] = hm->persist.n_slots++;
  }
  hm->pool.resize((size_t)hm->persist.n_slots * kBlockVoxels);
  uint32_t* bits = reinterpret_cast<uint32_t*>(hm->pool.data());
  for (size_t i = 0; i < hm->pool.size(); i++) bits[i] = kUnknownBits;
  m.pool = hm->pool.data();
  m.pool_cap = hm->persist.n_slots;
  for (size_t i = 0; i < n; i++) {
    const uint32_t h = find_block(m, block_key(k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]));
    hm->pool[(size_t)m.slot[h] * kBlockVoxels + local_index(k3[3 * i], k3[3 * i + 1], k3[3 * i + 2])] = logodds[i];
  }
  return hm;
}
void qc_map_destroy(void* p) { delete static_cast<HostMap*>(p); }

// getCellStatusBoundingBox for n points with one box, n_lanes emulated lanes per box
void qc_bbox_status(void* p, double res, float occ_thres, int treat_unknown_as_occupied, const double* xyz, size_t n,
                    const double* size3, uint32_t n_lanes, uint8_t* status) {
  HostMap* hm = static_cast<HostMap*>(p);
  DevParams P;
  memset(&P, 0, sizeof(P));
  P.res = res;
  P.res_factor = 1.0 / res;
  P.occ_thres = occ_thres;
  P.treat_unknown_as_occupied = treat_unknown_as_occupied;
  BoxQuery q;
  for (int a = 0; a < 3; a++) q.size[a] = size3[a];
  for (size_t i = 0; i < n; i++) {
    bool occ = false, unk = false;
    uint8_t decided = 0xFF;
    for (uint32_t lane = 0; lane < n_lanes; lane++) {
      const BoxScan r = bbox_scan(P, hm->view, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], q, lane, n_lanes);
      decided = r.decided;
      occ = occ || r.occupied;
      unk = unk || r.unknown;
    }
    status[i] = bbox_combine(P, decided, occ, unk);
  }
}

static DevParams qc_params(double res, float occ_thres, int treat_unknown_as_occupied) {
  DevParams P;
  memset(&P, 0, sizeof(P));
  P.res = res;
  P.res_factor = 1.0 / res;
  P.occ_thres = occ_thres;
  P.treat_unknown_as_occupied = treat_unknown_as_occupied;
  return P;
}

// getCellStatusPoint (true_status 0) / getCellTrueStatusPoint (1); logodds may be NULL
void qc_point_status(void* p, double res, float occ_thres, int unk, const double* xyz, size_t n, int true_status,
                     uint8_t* status, float* logodds) {
  HostMap* hm = static_cast<HostMap*>(p);
  const DevParams P = qc_params(res, occ_thres, unk);
  for (size_t i = 0; i < n; i++) {
    uint32_t bits;
    status[i] = point_status(P, hm->view, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], true_status, &bits);
    if (logodds) logodds[i] = __uint_as_float(bits == kUnknownBits ? 0x7FC00000u : bits);
  }
}

// getLineStatus
void qc_line_status(void* p, double res, float occ_thres, int unk, const double* ab, size_t n, uint8_t* status) {
  HostMap* hm = static_cast<HostMap*>(p);
  const DevParams P = qc_params(res, occ_thres, unk);
  for (size_t i = 0; i < n; i++)
    status[i] = line_status(P, hm->view, (float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2], (float)ab[6 * i + 3],
                            (float)ab[6 * i + 4], (float)ab[6 * i + 5]);
}

// getVisibility
void qc_visibility(void* p, double res, float occ_thres, int unk, const double* ab, size_t n, int stop_at_unknown,
                   uint8_t* status) {
  HostMap* hm = static_cast<HostMap*>(p);
  const DevParams P = qc_params(res, occ_thres, unk);
  for (size_t i = 0; i < n; i++)
    status[i] = visibility_status(P, hm->view, (float)ab[6 * i], (float)ab[6 * i + 1], (float)ab[6 * i + 2],
                                  (float)ab[6 * i + 3], (float)ab[6 * i + 4], (float)ab[6 * i + 5], stop_at_unknown);
}
}
