// vmap.cu -- host side of libvmap_b200.so: the C ABI of include/vmap/vmap.h, device memory
// management for the voxel-block map, and the launch sequences of the scan-insertion and query
// kernels (vmap_kernels.cuh).  There is no CPU fallback anywhere in this file: without a CUDA
// device vmap_create fails with VMAP_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/vmap/vmap.h"
#include "vmap_kernels.cuh"
#include "vmap_sort.cuh"
#include "vmap_shard.cuh"

namespace {

using namespace vmapdev;

thread_local std::string g_create_error;

struct DevBuf {  // grow-only device buffer
  void* p = nullptr;
  size_t cap = 0;
};

// dense mark window (vmap_device.cuh MarkSet): the per-scan key sets of ray-generated scans
struct Window {
  uint32_t* free_bits = nullptr;
  uint32_t* occ_bits = nullptr;
  uint32_t* touched_bits = nullptr;
  uint32_t* list = nullptr;
  uint32_t* locate = nullptr;
  uint32_t n_side = 0;        // blocks per axis needed (cube)
  uint32_t stride = 0;        // x / y extent as allocated: n_side, or the next power of two (dilated addressing)
  int reach_vox = -1;         // >= 0: dilated layout, the window starts at block (sensor voxel - reach_vox) >> 3
  int radius = 0;             // blocks from the sensor block to the window face
  uint64_t n_blocks = 0;
  uint32_t n_words = 0;       // touched bitmap words
  uint32_t list_cap = 0;
  bool disabled = false;      // window would exceed the budget / unlimited range
};

// What a pipelined scan needs to be run again from its staged input (recovery after a refusal).
struct ScanMeta {
  size_t n = 0, stride = 0;
  int is_dense = 1;
  double Tm[16] = {0};
};

// Everything ONE scan in flight owns: its stream, scratch, counters, staged input and mark window.
// `vmap` derives from it (the members below are "the current scan context" everywhere in this
// file) and keeps a second one in `alt`; pipelined insertion alternates between the two by swapping
// them, so consecutive scans overlap on the device while the map itself is updated in scan order.
struct ScanCtx {
  int ctx_id = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evk[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // per-kernel timing marks
  bool evk_set[5] = {false, false, false, false, false};
  cudaEvent_t evw[2] = {nullptr, nullptr};   // around the segment walker's launch alone (synchronous scans with stage marks)
  bool evw_set = false;
  float walk_kernel_ms = 0.f;                // ... its duration in the last such scan
  ScanBuffers sb{};
  size_t scan_cap = 0;
  size_t first_cap = 0;
  size_t ckpt_want = 0;       // checkpoint capacity to allocate with the scan scratch
  Counters* d_cnt = nullptr;
  Counters* h_cnt = nullptr;          // pinned
  Persist* h_persist_snap = nullptr;  // pinned: Persist as this context's last pipelined scan left it
  DevBuf in;                          // staged input (points / image / keys / query coordinates)
  Window win;
  MarkSet ms{};                       // marks of the scan in flight (table or dense mode)
  // pipelined insertion
  cudaEvent_t ev_done = nullptr;      // the scan in flight is complete, counters read back
  cudaEvent_t ev_h2d = nullptr;       // its input has left the caller's buffer
  bool in_flight = false;
  uint64_t seq = 0;                   // number of the scan in flight (order of insertion)
  ScanMeta meta;
};

}  // namespace

struct vmap : ScanCtx {
  vmap_params params;
  vmap_config cfg;
  DevParams P;
  int device = 0;
  cudaEvent_t evr0 = nullptr, evr1 = nullptr;  // timed region (vmap_region_begin / _end)
  ScanCtx alt;                   // the other scan context (created on the first pipelined insert)
  bool alt_ready = false;
  ScanCtx alt2;                  // a third one (single-GPU pipelined insertion, VMAP_PIPE_CTX=3): the next scan's front half
  bool alt2_ready = false;       // can be enqueued while two scans are still in flight, so the device never waits for the host
  int pipe_ctx = 3;              // scan contexts the single-GPU pipeline rotates through (2 or 3)
  int pipe_depth = 2;            // scans that may stay in flight when an insert call returns (0 = synchronous)
  int pipe_n = 0;                // scans in flight
  uint64_t pipe_seq = 0;
  bool pipe_grow = false;        // grow the map at the next quiescent point (headroom is getting thin)
  uint32_t pipe_touch_max = 0;   // most blocks one scan has touched so far
  uint64_t pipe_stalls = 0;      // scans whose update was refused on the device and replayed
  vmap_scan_stats totals{};      // sums over every scan inserted since creation / reset
  uint64_t total_scans = 0;
  cudaEvent_t ev_wb = nullptr;   // front-end kernel done (write-back of the world-frame cloud)
  DevBuf wb;                     // write-back image for strides != 12
  uint8_t* wb_host = nullptr;    // pinned staging for the NaN-compacting write-back
  size_t wb_host_cap = 0;

  // voxel-block map
  MapView m{};
  uint64_t table_cap = 0;
  uint64_t pool_cap = 0;
  uint32_t epoch = 0;

  // per-scan scratch
  Persist* h_persist = nullptr;  // pinned
  DevBuf out;                    // staged output (statuses, exported leaves)
  DevBuf aux;                    // second staged input / output
  SortBuffers sort{};            // VMAP_DEDUPE_SORT scratch

  // multi-GPU sharding (vmap_shard.cuh): this handle holds the blocks with owner_of(block) ==
  // shard_rank; scans generated here are packed into per-owner record streams
  int shard_rank = 0, shard_world = 1;
  DevBuf pack, recv;
  uint32_t* d_pack_counts_base = nullptr;     // [2 slots][kMaxBatch sub-scans][2 * kMaxRanks] counts + cursors
  uint32_t* d_pack_counts = nullptr;          // current (slot, sub-scan)
  uint32_t h_pack_counts_slot[2][kMaxBatch][2 * kMaxRanks] = {};
  uint32_t* h_pack_counts = h_pack_counts_slot[0][0];
  int pack_slot = 0;
  int pack_sub = 0;                           // sub-scans generated into the current batch
  size_t pack_off = 0;                        // records already in `pack` (earlier sub-scans)
  int batch_m = 1;                            // scans per rank per exchange (vmap_dist_set_batch)
  int round_subs = 0;                         // sub-scans of the pending round
  bool in_collective = false;                 // inside a vmap_dist_* call (may run NCCL)
  bool round_pending = false;                 // a generated + packed round awaits its exchange
  bool round_packed = false;                  // ... and its counts are on the device
  int round_slot = 0;
  vmap_scan_stats round_stats{};
  void* nccl_comm = nullptr;
  double dbg_ms[4] = {0, 0, 0, 0};            // VMAP_TIMING: host wall time per phase of a round
  uint64_t dbg_n = 0;
  bool dbg_reset = false;
  // overlapped rounds: exchange + apply of round k run on app_stream while round k+1 is generated
  cudaStream_t app_stream = nullptr;
  Counters* d_cnt_app = nullptr;
  Counters* h_cnt_app = nullptr;              // pinned
  Persist* h_persist_app = nullptr;           // pinned
  DevBuf pack_inflight;                       // records being sent (generate writes into `pack`)
  bool app_pending = false;
  cudaEvent_t ev_app0 = nullptr, ev_app1 = nullptr;
  DevBuf l2_flush;                            // vmap_set_debug_l2_flush
  size_t l2_flush_bytes = 0;
  uint64_t app_records = 0, app_sent = 0, app_received = 0;
  bool packed = false;                        // the last generate already packed its records
  uint32_t* d_count_matrix = nullptr;         // [world][kMaxBatch][2 * kMaxRanks]
  std::vector<uint32_t> h_count_matrix;

  // peer-memory transport of the sharded build (vmap_peer_host.inl)
  bool peer_on = false;
  void* mailbox = nullptr;
  size_t mailbox_bytes = 0;
  uint32_t peer_capacity = 0;                 // records per round buffer
  PeerTable pt{};
  void* peer_ipc_base[kMaxRanks] = {};        // mappings opened with cudaIpcOpenMemHandle
  uint32_t* d_round_info = nullptr;           // [2 parities][kRoundInfoWords]
  uint32_t* h_round_info = nullptr;           // pinned mirror
  uint32_t* d_sub_counts = nullptr;           // [2 scan contexts][counts + cursors]
  RecordLists* d_lists = nullptr;             // the round's record lists as the apply kernels see them
  RecordLists* h_lists = nullptr;             // pinned staging of the same
  cudaEvent_t ev_split[4] = {nullptr, nullptr, nullptr, nullptr};   // VMAP_PEER_SERIAL + VMAP_TIMING: per-kernel times of the apply
  uint32_t* d_located = nullptr;              // table index of every record of the round (k_locate_records)
  size_t located_cap = 0;
  cudaEvent_t ev_collect[2] = {nullptr, nullptr};
  uint64_t round_open = 0;                    // number of the round being packed
  int sub_open = 0;                           // scans packed into it so far
  uint64_t rounds_published = 0, rounds_applied = 0;
  bool peer_local = false;                    // vmap_dist_attach_local: the "peers" are handles of this process
  std::vector<std::pair<void*, size_t>> graveyard;   // device buffers retired while peers may be spinning (see dev_free)

  uint64_t dense_fallbacks = 0;
  bool traced = false;        // VMAP_TRACE: the walker selection has been printed
  uint32_t* chg0_store = nullptr;   // change-detection planes (MapView.chg0/1 point here while enabled)
  uint32_t* chg1_store = nullptr;
  int walk_variant = 48;  // near-field limit of the walker's read-before-RED filter (VMAP_WALK_VARIANT)
  int walk_ctas = 0;      // CTAs per SM of the segment walker's grid (VMAP_WALK_CTAS); 0 = 5 (all that fit at 48 registers) for a
                          // scan that runs alone, 4 for a pipelined scan: the fifth slot per SM is what lets the other scan's
                          // short kernels run beside the walk (pipelined 0.263 -> 0.251 ms per scan, alone 0.378 -> 0.393)
  bool walk_overlaps = false;   // the scan being enqueued shares the device with another scan in flight
  // vmap_stage_pointcloud / vmap_insert_staged: two device staging buffers filled on their own stream
  DevBuf stage[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  size_t stage_n[2] = {0, 0}, stage_stride[2] = {0, 0};
  bool staged[2] = {false, false};
  int stage_next = 0;         // slot the next vmap_stage_pointcloud fills
  int walk_order = 2;     // bit 0: golden-ratio chunk order in the dilated segment walker (VMAP_WALK_ORDER, experimental);
                          // bit 1: counted loops in non-last segments (default; VMAP_WALK_COUNTED=0 turns it off)
  int walk_addr = 9;      // segment walker: 9 dilated addressing + queued 64-voxel marks (default), 7 unqueued, 1 32-voxel words, 0 block addressing, 5 / 8 experiments (VMAP_WALK_ADDR)
  int update_variant = 2; // K4: 2 k_update_mlp (default; rows in flight together, blocks pipelined), 0 k_update<0> (row by row; also what change detection uses), 1 experiment (VMAP_UPDATE_VARIANT)
  int pdl = 123;          // VMAP_PDL: programmatic dependent launch along the scan's kernel chain (bit mask kPdl*; default: all but the walker)
  int stage_marks = -1;   // VMAP_STAGE_MARKS: 1 always, 0 never, unset: synchronous scans only
  int flush_at = 1;       // VMAP_FLUSH_AT: where a pipelined scan runs the benchmark's L2 flush (0 scan start, 1 before its update)
  int helpers = 1;        // VMAP_HELPERS: 1 folded helper chain (k_resolve_fold, k_checkpoints_pts), 0 round 1's
  int no_segments = 0;    // experiments (VMAP_NO_SEGMENTS): one thread per whole ray

  // parity hook
  bool capture = false;
  bool captured = false;
  DevBuf cap_free, cap_occ;
  uint32_t* d_cap_n = nullptr;  // [2]
  uint32_t h_cap_n[2] = {0, 0};

  vmap_scan_stats stats{};
  mutable std::string err;
  uint64_t launches = 0;
  uint64_t dev_bytes = 0;
};

namespace {

#define VM_CUDA(h, expr)                                                                    \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                       \
      return (e__ == cudaErrorMemoryAllocation) ? VMAP_ERR_CAPACITY : VMAP_ERR_CUDA;        \
    }                                                                                       \
  } while (0)

#define VM_TRY(expr)                 \
  do {                               \
    int s__ = (expr);                \
    if (s__ != VMAP_OK) return s__;  \
  } while (0)

// No C++ exception may cross the C ABI: every entry point is a function-try-block that ends here.
static int vmap_on_exception(vmap* h) noexcept {
  int code = VMAP_ERR_INVALID_ARGUMENT;
  try {
    std::string what = "unknown C++ exception";
    try {
      throw;
    } catch (const std::bad_alloc&) {
      what = "host memory allocation failed";
      code = VMAP_ERR_CAPACITY;
    } catch (const std::exception& e) {
      what = std::string("C++ exception: ") + e.what();
    } catch (...) {
    }
    if (h) h->err = what;
    else g_create_error = what;
  } catch (...) {
  }
  return code;
}

int fail(const vmap* h, int code, const char* msg) {
  h->err = msg;
  return code;
}

int flush_rounds(vmap* h); // vmap_shard_host.inl: wait for an overlapped exchange/apply round
int pipe_flush(vmap* h);   // below: drain the pipelined scans
// Every entry point that reads or edits the map outside the scan pipeline starts here: nothing is
// in flight on the handle afterwards.  Deferred failures of pipelined scans surface here.
int flush_app(vmap* h) {
  const int st = pipe_flush(h);
  const int st2 = flush_rounds(h);
  return st != VMAP_OK ? st : st2;
}

int dev_alloc(vmap* h, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) return VMAP_OK;
  VM_CUDA(h, cudaMalloc(p, bytes));
  h->dev_bytes += bytes;
  return VMAP_OK;
}
// cudaFree waits for EVERYTHING on the device.  With the peer-memory transport a one-warp kernel of
// this handle may be spinning on a flag a peer sets only after this host thread has made progress
// (and, with vmap_dist_attach_local, the peers' spinners live in this very context), so buffers
// retired while the transport is on are parked and released at the next quiescent point.
void dev_free(vmap* h, void* p, size_t bytes) {
  if (!p) return;
  if (h->peer_on) {
    h->graveyard.emplace_back(p, bytes);
    return;
  }
  cudaFree(p);
  h->dev_bytes -= bytes;
}
void empty_graveyard(vmap* h) {
  for (auto& g : h->graveyard) {
    cudaFree(g.first);
    h->dev_bytes -= g.second;
  }
  h->graveyard.clear();
}
int ensure(vmap* h, DevBuf* b, size_t bytes) {
  if (bytes <= b->cap) return VMAP_OK;
  // in-flight work may still reference the old buffer
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  dev_free(h, b->p, b->cap);
  b->p = nullptr;
  b->cap = 0;
  size_t want = std::max(bytes, (size_t)4096);
  want = (want + (want >> 2) + 255) & ~(size_t)255;
  VM_TRY(dev_alloc(h, &b->p, want));
  b->cap = want;
  return VMAP_OK;
}

uint64_t next_pow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

float logodds(double p) { return (float)std::log(p / (1 - p)); }  // octomap_utils.h

void derive_params(vmap* h) {
  const vmap_params& p = h->params;
  DevParams& P = h->P;
  P.res = p.resolution;
  P.res_factor = 1.0 / p.resolution;
  P.max_range = p.sensor_max_range;
  P.max_free_space = p.max_free_space;
  P.min_height_free_space = p.min_height_free_space;
  // setProbHit/Miss, setClampingThresMin/Max, setOccupancyThres (octomap_world.cc:94-98)
  P.hit = logodds(p.probability_hit);
  P.miss = logodds(p.probability_miss);
  P.cmin = logodds(p.threshold_min);
  P.cmax = logodds(p.threshold_max);
  P.occ_thres = logodds(p.threshold_occupancy);
  P.treat_unknown_as_occupied = p.treat_unknown_as_occupied ? 1 : 0;
  P.free_filter = (p.max_free_space != 0.0) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// map storage
// ---------------------------------------------------------------------------------------------
int alloc_table(vmap* h, uint64_t cap, MapView* out) {
  MapView v = h->m;
  VM_TRY(dev_alloc(h, (void**)&v.keys, cap * sizeof(unsigned long long)));
  VM_TRY(dev_alloc(h, (void**)&v.slot, cap * sizeof(uint32_t)));
  VM_TRY(dev_alloc(h, (void**)&v.free_mask, cap * kMaskWords * sizeof(uint32_t)));
  VM_TRY(dev_alloc(h, (void**)&v.occ_mask, cap * kMaskWords * sizeof(uint32_t)));
  VM_TRY(dev_alloc(h, (void**)&v.touched, cap * sizeof(uint32_t)));
  VM_CUDA(h, cudaMemsetAsync(v.keys, 0xFF, cap * sizeof(unsigned long long), h->stream));
  VM_CUDA(h, cudaMemsetAsync(v.slot, 0xFF, cap * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(v.free_mask, 0, cap * kMaskWords * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(v.occ_mask, 0, cap * kMaskWords * sizeof(uint32_t), h->stream));
  v.cap_mask = (uint32_t)(cap - 1);
  *out = v;
  return VMAP_OK;
}
void free_table(vmap* h, const MapView& v, uint64_t cap) {
  dev_free(h, v.keys, cap * sizeof(unsigned long long));
  dev_free(h, v.slot, cap * sizeof(uint32_t));
  dev_free(h, v.free_mask, cap * kMaskWords * sizeof(uint32_t));
  dev_free(h, v.occ_mask, cap * kMaskWords * sizeof(uint32_t));
  dev_free(h, v.touched, cap * sizeof(uint32_t));
}

// Double the hash table.  Only legal when no scan marks are pending (masks all zero).
int grow_table(vmap* h) {
  const uint64_t old_cap = h->table_cap;
  const uint64_t new_cap = old_cap * 2;
  if (new_cap > (1ull << 31)) return fail(h, VMAP_ERR_CAPACITY, "voxel-block table limit (2^31 entries) reached");
  MapView nv;
  VM_TRY(alloc_table(h, new_cap, &nv));
  k_rehash<<<(unsigned)((old_cap + 255) / 256), 256, 0, h->stream>>>(
      h->m.keys, h->m.slot, (uint32_t)old_cap, nv.keys, nv.slot, nv.cap_mask);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  free_table(h, h->m, old_cap);
  h->m = nv;
  h->table_cap = new_cap;
  return VMAP_OK;
}

int grow_pool(vmap* h, uint64_t need_blocks) {
  uint64_t new_cap = h->pool_cap;
  while (new_cap < need_blocks) new_cap *= 2;
  if (h->cfg.max_blocks && new_cap > h->cfg.max_blocks) {
    if (need_blocks > h->cfg.max_blocks)
      return fail(h, VMAP_ERR_CAPACITY, "voxel-block pool would exceed vmap_config.max_blocks");
    new_cap = h->cfg.max_blocks;
  }
  if (new_cap >= 0xFFFFFFFFull) return fail(h, VMAP_ERR_CAPACITY, "voxel-block pool limit reached");
  float* np_ = nullptr;
  const size_t blk = (size_t)kBlockVoxels * sizeof(float);
  VM_TRY(dev_alloc(h, (void**)&np_, new_cap * blk));
  VM_CUDA(h, cudaMemcpyAsync(np_, h->m.pool, h->pool_cap * blk, cudaMemcpyDeviceToDevice, h->stream));
  VM_CUDA(h, cudaMemsetAsync((char*)np_ + h->pool_cap * blk, 0xFF, (new_cap - h->pool_cap) * blk, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->chg0_store) {   // change-detection planes grow with the pool
    const size_t pl = (size_t)kMaskWords * sizeof(uint32_t);
    uint32_t* n0 = nullptr;
    uint32_t* n1 = nullptr;
    VM_TRY(dev_alloc(h, (void**)&n0, new_cap * pl));
    VM_TRY(dev_alloc(h, (void**)&n1, new_cap * pl));
    VM_CUDA(h, cudaMemcpy(n0, h->chg0_store, h->pool_cap * pl, cudaMemcpyDeviceToDevice));
    VM_CUDA(h, cudaMemcpy(n1, h->chg1_store, h->pool_cap * pl, cudaMemcpyDeviceToDevice));
    VM_CUDA(h, cudaMemset((char*)n0 + h->pool_cap * pl, 0, (new_cap - h->pool_cap) * pl));
    VM_CUDA(h, cudaMemset((char*)n1 + h->pool_cap * pl, 0, (new_cap - h->pool_cap) * pl));
    dev_free(h, h->chg0_store, h->pool_cap * pl);
    dev_free(h, h->chg1_store, h->pool_cap * pl);
    h->chg0_store = n0;
    h->chg1_store = n1;
    if (h->m.chg0) {
      h->m.chg0 = n0;
      h->m.chg1 = n1;
    }
  }
  dev_free(h, h->m.pool, h->pool_cap * blk);
  h->m.pool = np_;
  h->pool_cap = new_cap;
  h->m.pool_cap = (uint32_t)new_cap;
  return VMAP_OK;
}

// octree_->enableChangeDetection(on) (octomap_world.cc:99): the planes are allocated on first use
// and kept (disabling only stops the tracking, like octomap keeps its changed_keys).
int set_change_detection(vmap* h, bool on) {
  if (on && !h->chg0_store) {
    const size_t pl = (size_t)kMaskWords * sizeof(uint32_t);
    VM_TRY(dev_alloc(h, (void**)&h->chg0_store, h->pool_cap * pl));
    VM_TRY(dev_alloc(h, (void**)&h->chg1_store, h->pool_cap * pl));
    VM_CUDA(h, cudaMemset(h->chg0_store, 0, h->pool_cap * pl));
    VM_CUDA(h, cudaMemset(h->chg1_store, 0, h->pool_cap * pl));
  }
  h->m.chg0 = on ? h->chg0_store : nullptr;
  h->m.chg1 = on ? h->chg1_store : nullptr;
  return VMAP_OK;
}

int init_map(vmap* h) {
  const uint64_t blocks = h->cfg.initial_blocks ? h->cfg.initial_blocks : (1ull << 17);
  h->pool_cap = std::max<uint64_t>(blocks, 64);
  if (h->cfg.max_blocks) h->pool_cap = std::min(h->pool_cap, h->cfg.max_blocks);
  h->table_cap = next_pow2(h->pool_cap * 2);
  h->m = MapView{};
  VM_TRY(alloc_table(h, h->table_cap, &h->m));
  const size_t blk = (size_t)kBlockVoxels * sizeof(float);
  VM_TRY(dev_alloc(h, (void**)&h->m.pool, h->pool_cap * blk));
  VM_CUDA(h, cudaMemsetAsync(h->m.pool, 0xFF, h->pool_cap * blk, h->stream));
  VM_TRY(dev_alloc(h, (void**)&h->m.persist, sizeof(Persist)));
  VM_CUDA(h, cudaMemsetAsync(h->m.persist, 0, sizeof(Persist), h->stream));
  h->m.pool_cap = (uint32_t)h->pool_cap;
  h->epoch = 0;
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
}

void free_map(vmap* h) {
  if (h->table_cap) free_table(h, h->m, h->table_cap);
  dev_free(h, h->m.pool, h->pool_cap * kBlockVoxels * sizeof(float));
  dev_free(h, h->m.persist, sizeof(Persist));
  h->m = MapView{};
  h->table_cap = h->pool_cap = 0;
}

// Advance the scan epoch (entries whose stored epoch differs are "not yet touched this scan").
int next_epoch(vmap* h) {
  if (h->epoch >= kEpochMax) {
    k_reset_epochs<<<(unsigned)((h->table_cap + 255) / 256), 256, 0, h->stream>>>(h->m.keys, (uint32_t)h->table_cap);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
    h->epoch = 0;
  }
  h->epoch++;
  h->m.epoch = h->epoch;
  return VMAP_OK;
}

// ---------------------------------------------------------------------------------------------
// per-scan scratch
// ---------------------------------------------------------------------------------------------
int ensure_scan(vmap* h, size_t n) {
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "more than 2^31-1 points in one scan");
  if (n <= h->scan_cap && h->scan_cap) return VMAP_OK;
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  ScanBuffers& sb = h->sb;
  const size_t oc = h->scan_cap, ofc = h->first_cap;
  dev_free(h, sb.world, oc * 12);
  dev_free(h, sb.ray_end, oc * 12);
  dev_free(h, sb.key, oc * 8);
  dev_free(h, sb.flags, oc);
  dev_free(h, sb.rays, oc * 4);
  dev_free(h, sb.len, oc * 4);
  dev_free(h, sb.rank, oc * 4);
  dev_free(h, sb.seg_base, (kLenBins + 1) * 4);
  dev_free(h, sb.bin_off, kFineBins * 4);
  dev_free(h, sb.ray_rec, oc * sizeof(RayRec));
  dev_free(h, sb.ckpt, (size_t)sb.ckpt_cap * sizeof(Checkpoint));
  dev_free(h, sb.first_keys, ofc * 12);
  sb = ScanBuffers{};
  h->scan_cap = h->first_cap = 0;
  size_t cap = std::max<size_t>(n + (n >> 2), std::max<uint64_t>(h->cfg.initial_scan_points, 16384));
  const size_t fc = (size_t)next_pow2(2 * cap);
  VM_TRY(dev_alloc(h, (void**)&sb.world, cap * 12));
  VM_TRY(dev_alloc(h, (void**)&sb.ray_end, cap * 12));
  VM_TRY(dev_alloc(h, (void**)&sb.key, cap * 8));
  VM_TRY(dev_alloc(h, (void**)&sb.flags, cap));
  VM_TRY(dev_alloc(h, (void**)&sb.rays, cap * 4));
  VM_TRY(dev_alloc(h, (void**)&sb.len, cap * 4));
  VM_TRY(dev_alloc(h, (void**)&sb.rank, cap * 4));
  sb.len_hist = reinterpret_cast<uint32_t*>(h->d_cnt + 1);
  VM_TRY(dev_alloc(h, (void**)&sb.seg_base, (kLenBins + 1) * 4));
  VM_TRY(dev_alloc(h, (void**)&sb.bin_off, kFineBins * 4));
  VM_TRY(dev_alloc(h, (void**)&sb.ray_rec, cap * sizeof(RayRec)));
  {
    const size_t cc = std::max<size_t>(cap * 8, h->ckpt_want);
    VM_TRY(dev_alloc(h, (void**)&sb.ckpt, cc * sizeof(Checkpoint)));
    sb.ckpt_cap = (uint32_t)std::min<size_t>(cc, 0xFFFFFFFFu);
  }
  VM_TRY(dev_alloc(h, (void**)&sb.first_keys, fc * 12));   // keys, then the indices of the table in use
  sb.first_idx = reinterpret_cast<uint32_t*>(sb.first_keys + fc);
  sb.first_mask = (uint32_t)(fc - 1);
  h->scan_cap = cap;
  h->first_cap = fc;
  return VMAP_OK;
}

int begin_scan_scratch(vmap* h, size_t n) {
  // size the first-index table to the scan (power of two >= 2n) so clearing stays proportional
  const size_t fc = std::min(h->first_cap, (size_t)next_pow2(std::max<size_t>(2 * n, 1024)));
  h->sb.first_mask = (uint32_t)(fc - 1);
  // two memsets per scan: counters + histogram (zero), first-index table of fc entries (0xFF:
  // keys [0, fc) followed directly by their indices)
  h->sb.first_idx = reinterpret_cast<uint32_t*>(h->sb.first_keys + fc);
  VM_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters) + kFineBins * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->sb.first_keys, 0xFF, fc * 12, h->stream));
  return VMAP_OK;
}

int read_counters(vmap* h) {
  VM_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(h->h_persist, h->m.persist, sizeof(Persist), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
}

// Stage marks for vmap_scan_stats.{front,resolve,walk,update}_ms: 0 = scan start, 1 = after the
// front-end kernel, 2 = after resolve, 3 = after the walk (key-set complete), 4 = after update.
int mark_stage(vmap* h, int i) {
  // (an event between two kernels keeps the second from launching programmatically: pipelined scans
  // record no stage marks unless VMAP_STAGE_MARKS=1 -- their stage times read 0)
  if (h->stage_marks == 0 || (h->stage_marks < 0 && h->walk_overlaps)) return VMAP_OK;
  VM_CUDA(h, cudaEventRecord(h->evk[i], h->stream));
  h->evk_set[i] = true;
  return VMAP_OK;
}

// Launch with the programmatic-serialization attribute (vmap_device.cuh: pdl_wait / pdl_launch): ONLY for kernels
// that call pdl_wait() before they touch global memory.  h->pdl == 0 (VMAP_PDL=0): an ordinary launch.
enum { kPdlResolve = 1, kPdlCkpt = 2, kPdlWalk = 4, kPdlList = 8, kPdlAdmit = 16, kPdlLocate = 32, kPdlUpdate = 64 };
template <class... KArgs, class... Args>
cudaError_t launch_chain(const vmap* h, int which, void (*kernel)(KArgs...), unsigned grid, unsigned block, Args&&... args) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (h->pdl & which) ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

unsigned grid_for(size_t n, unsigned block) { return (unsigned)std::max<size_t>(1, (n + block - 1) / block); }

int g_sm_count(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) device = 0;
  if (!cached[device]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    cached[device] = v > 0 ? v : 148;
  }
  return cached[device];
}

int grow_checkpoints(vmap* h) {
  ScanBuffers& sb = h->sb;
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  const size_t want = std::min<size_t>((size_t)sb.ckpt_cap * 2 + 1024, 0xFFFFFFFFu);
  if (want <= sb.ckpt_cap) return fail(h, VMAP_ERR_CAPACITY, "segment checkpoint buffer limit reached");
  dev_free(h, sb.ckpt, (size_t)sb.ckpt_cap * sizeof(Checkpoint));
  sb.ckpt = nullptr;
  sb.ckpt_cap = 0;
  VM_TRY(dev_alloc(h, (void**)&sb.ckpt, want * sizeof(Checkpoint)));
  sb.ckpt_cap = (uint32_t)want;
  h->ckpt_want = want;
  return VMAP_OK;
}

// ---------------------------------------------------------------------------------------------
// mark sets
// ---------------------------------------------------------------------------------------------
void use_table_marks(vmap* h) {
  MarkSet& ms = h->ms;
  ms = MarkSet{};
  ms.free_mask = h->m.free_mask;
  ms.occ_mask = h->m.occ_mask;
  ms.list = h->m.touched;
  ms.list_cap = (uint32_t)h->table_cap;
  ms.dense = 0;
}

void free_window(vmap* h) {
  Window& w = h->win;
  dev_free(h, w.free_bits, w.n_blocks * kMaskWords * 4);
  dev_free(h, w.occ_bits, w.n_blocks * kMaskWords * 4);
  dev_free(h, w.touched_bits, (size_t)w.n_words * 4);
  dev_free(h, w.list, (size_t)w.list_cap * 4);
  dev_free(h, w.locate, (size_t)w.list_cap * 4);
  const bool dis = w.disabled;
  w = Window{};
  w.disabled = dis;
}

int grow_window_list(vmap* h, uint32_t want) {
  Window& w = h->win;
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  dev_free(h, w.list, (size_t)w.list_cap * 4);
  dev_free(h, w.locate, (size_t)w.list_cap * 4);
  w.list = w.locate = nullptr;
  w.list_cap = 0;
  const uint64_t cap = std::min<uint64_t>(w.n_blocks, std::max<uint64_t>(want, 1u << 16));
  VM_TRY(dev_alloc(h, (void**)&w.list, cap * 4));
  VM_TRY(dev_alloc(h, (void**)&w.locate, cap * 4));
  w.list_cap = (uint32_t)cap;
  return VMAP_OK;
}

// A cube of 8^3 blocks centred on the sensor block that holds every voxel a ray of this scan can
// visit: |point - origin| <= sensor_max_range after clipping (octomap_world.cc:184-185,210-212).
// Returns false (table mode) for an unlimited range or when the cube exceeds the memory budget.
bool setup_dense_marks(vmap* h, const float origin[3]) {
  Window& w = h->win;
  if (w.disabled || h->cfg.dedupe_mode != VMAP_DEDUPE_BITMASK) return false;
  if (!(h->P.max_range >= 0.0) || !std::isfinite(h->P.max_range)) return false;
  const double r_vox = h->P.max_range * h->P.res_factor;
  const double r_blk = std::ceil(r_vox / 8.0) + 2.0;
  const uint64_t budget = h->cfg.dense_window_bytes ? h->cfg.dense_window_bytes : (6ull << 30);
  double side = 2.0 * r_blk + 1.0;
  uint32_t stride = (uint32_t)std::min(side, 4294967295.0);
  int reach_vox = -1;
  if (h->walk_addr >= 1 && r_vox < 262144.0) {
    // Dilated addressing wants power-of-two x / y extents and a 31-bit bit address.  Placing the
    // window by the sensor's VOXEL (not its block) needs only floor(2 * reach / 8) + 2 blocks per
    // axis, reach = the farthest voxel a clipped ray can visit + the walker's slack.
    const int reach = (int)std::ceil(r_vox) + 3;
    const uint32_t need = (uint32_t)(2 * reach) / 8u + 2u;
    uint32_t p2 = 1;
    while (p2 < need) p2 <<= 1;
    if (dilated_window_fits(p2, p2, need)) {
      reach_vox = reach;
      side = (double)need;
      stride = p2;
    }
  }
  const double blocks = (double)stride * (double)stride * side;
  if (blocks * (2.0 * kMaskWords * 4 + 0.125) > (double)budget || blocks >= (double)(1u << 28)) return false;
  const int radius = (int)r_blk;
  if (w.radius != radius || w.stride != stride || w.reach_vox != reach_vox || !w.free_bits) {
    cudaStreamSynchronize(h->stream);
    free_window(h);
    w.radius = radius;
    w.n_side = (uint32_t)side;
    w.stride = stride;
    w.reach_vox = reach_vox;
    w.n_blocks = (uint64_t)w.stride * w.stride * w.n_side;
    w.n_words = (uint32_t)((w.n_blocks + 31) / 32);
    bool ok = dev_alloc(h, (void**)&w.free_bits, w.n_blocks * kMaskWords * 4) == VMAP_OK &&
              dev_alloc(h, (void**)&w.occ_bits, w.n_blocks * kMaskWords * 4) == VMAP_OK &&
              dev_alloc(h, (void**)&w.touched_bits, (size_t)w.n_words * 4) == VMAP_OK;
    if (ok) ok = cudaMemsetAsync(w.free_bits, 0, w.n_blocks * kMaskWords * 4, h->stream) == cudaSuccess &&
                 cudaMemsetAsync(w.occ_bits, 0, w.n_blocks * kMaskWords * 4, h->stream) == cudaSuccess &&
                 cudaMemsetAsync(w.touched_bits, 0, (size_t)w.n_words * 4, h->stream) == cudaSuccess;
    if (ok) ok = grow_window_list(h, 1u << 18) == VMAP_OK;
    if (!ok) {   // not enough device memory for the window: stay in table mode for good
      cudaGetLastError();
      free_window(h);
      w.disabled = true;
      return false;
    }
  }
  int ok[3];
  for (int i = 0; i < 3; i++) {
    const double f = std::floor(h->P.res_factor * (double)origin[i]);
    // sensor outside the map: its rays are never walked; park the window at the map centre
    ok[i] = (f >= -32768.0 && f < 32768.0) ? (((int)f + 32768) >> 3) : 4096;
  }
  MarkSet& ms = h->ms;
  ms = MarkSet{};
  ms.free_mask = w.free_bits;
  ms.occ_mask = w.occ_bits;
  ms.touched_bits = w.touched_bits;
  ms.list = w.list;
  ms.locate = w.locate;
  ms.list_cap = w.list_cap;
  ms.x0 = ok[0] - radius; ms.y0 = ok[1] - radius; ms.z0 = ok[2] - radius;
  if (w.reach_vox >= 0) {
    int first[3];
    for (int i = 0; i < 3; i++) {
      const double f = std::floor(h->P.res_factor * (double)origin[i]);
      first[i] = (f >= -32768.0 && f < 32768.0) ? (((int)f + 32768 - w.reach_vox) >> 3) : 4096;
    }
    ms.x0 = first[0]; ms.y0 = first[1]; ms.z0 = first[2];
  }
  ms.nx = ms.ny = w.stride;
  ms.nz = w.n_side;
  ms.dense = 1;
  return true;
}

int launch_list(vmap* h) {
  const unsigned grid = (unsigned)std::min<uint64_t>((h->win.n_words + 255) / 256, (uint64_t)g_sm_count(h->device) * 8);
  VM_CUDA(h, launch_chain(h, kPdlList, k_list_touched, std::max(grid, 1u), 256, h->ms, h->d_cnt, h->win.n_words));
  h->launches++;
  return VMAP_OK;
}

int clear_window_marks(vmap* h) {
  k_clear_window<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->ms, h->win.n_words);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  return VMAP_OK;
}

// Parity hook: count the scan's key sets, size the capture buffers, then expand the masks.
int launch_capture(vmap* h) {
  if (!h->capture) return VMAP_OK;
  for (int pass = 0; pass < 2; pass++) {
    VM_CUDA(h, cudaMemsetAsync(h->d_cap_n, 0, 2 * sizeof(uint32_t), h->stream));
    k_capture_keys<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(
        h->m, h->ms, h->d_cnt, (uint16_t*)h->cap_free.p, h->d_cap_n, pass ? (uint32_t)(h->cap_free.cap / 6) : 0u,
        (uint16_t*)h->cap_occ.p, h->d_cap_n + 1, pass ? (uint32_t)(h->cap_occ.cap / 6) : 0u);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
    if (pass == 0) {
      VM_CUDA(h, cudaMemcpyAsync(h->h_cap_n, h->d_cap_n, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
      VM_CUDA(h, cudaStreamSynchronize(h->stream));
      VM_TRY(ensure(h, &h->cap_free, (size_t)h->h_cap_n[0] * 6 + 6));
      VM_TRY(ensure(h, &h->cap_occ, (size_t)h->h_cap_n[1] * 6 + 6));
    }
  }
  return VMAP_OK;
}

int launch_update(vmap* h) {
  if (h->update_variant == 2 && h->m.chg0 == nullptr) {
    static int per_sm = 0;      // CTAs of k_update_mlp resident per SM
    if (!per_sm) {
      int v = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_update_mlp, 256, 0) != cudaSuccess || v < 1) { cudaGetLastError(); v = 4; }
      per_sm = v;
    }
    VM_CUDA(h, launch_chain(h, kPdlUpdate, k_update_mlp, g_sm_count(h->device) * (unsigned)per_sm, 256, h->P, h->m, h->ms, h->d_cnt));
  } else if (h->update_variant == 1) VM_CUDA(h, launch_chain(h, kPdlUpdate, k_update<1>, g_sm_count(h->device) * 8, 256, h->P, h->m, h->ms, h->d_cnt));
  else VM_CUDA(h, launch_chain(h, kPdlUpdate, k_update<0>, g_sm_count(h->device) * 8, 256, h->P, h->m, h->ms, h->d_cnt));
  h->launches++;
  return VMAP_OK;
}

// The ordered half of a scan ("tail"): everything that touches the MAP.  Dense mode: admission
// (one thread), table entries of the listed window blocks (+ the pool verdict by the last CTA),
// update.  Table mode: the pool verdict (one thread), update.  A refused tail leaves the scan's
// marks, list and touched bitmap as they are, so it can be run again after the map has grown.
// poison != 0 (pipelined insertion): a refusal also stalls every later tail (Persist::stalled).
int launch_tail(vmap* h, int poison) {
  if (h->ms.dense) {
    VM_CUDA(h, launch_chain(h, kPdlAdmit, k_admit, 1, 32, h->ms, h->m, h->d_cnt, poison));
    VM_CUDA(h, launch_chain(h, kPdlLocate, k_locate_blocks, g_sm_count(h->device) * 4, 256, h->ms, h->m, h->d_cnt, poison));
    h->launches += 2;
  } else {
    k_admit_pool<<<1, 32, 0, h->stream>>>(h->m, h->d_cnt);
    h->launches++;
  }
  VM_CUDA(h, cudaGetLastError());
  VM_TRY(mark_stage(h, 3));
  VM_TRY(launch_capture(h));
  if (h->capture) VM_TRY(mark_stage(h, 3));
  VM_TRY(launch_update(h));
  VM_TRY(mark_stage(h, 4));
  return VMAP_OK;
}

int clear_marks(vmap* h) {
  k_clear_masks<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->m, h->ms, h->d_cnt);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  return VMAP_OK;
}

// Dense mode, marks complete in the window and listed: run the tail, growing the table / the pool
// and running it again for as long as the device refuses it.  Synchronous (the caller has made
// sure nothing else is in flight on this handle).  On a capacity error the scan's marks are
// dropped, so the next scan does not inherit them.
int run_tail_sync(vmap* h, uint64_t* retries, bool first_attempt_done = false) {
  for (int guard = 0;; guard++) {
    if (!(first_attempt_done && guard == 0)) {
      VM_TRY(launch_tail(h, 0));
      VM_TRY(read_counters(h));
    }
    const Counters& c = *h->h_cnt;
    if (!(c.stall || c.overflow_table || c.overflow_pool)) return VMAP_OK;
    int gs = VMAP_OK;
    if (guard >= 40) gs = fail(h, VMAP_ERR_CAPACITY, "scan could not be applied after repeated growth");
    else if (c.stall & kStallMarks) gs = fail(h, VMAP_ERR_CUDA, "internal: tail run on incomplete marks");
    else if ((c.stall & kStallTable) || c.overflow_table) gs = grow_table(h);
    else if ((c.stall & kStallPool) || c.overflow_pool) gs = grow_pool(h, (uint64_t)h->h_persist->n_slots + c.need_slots);
    // (kStallEarlier alone: the stalled flag was left set by a pipelined scan; cleared below)
    if (gs != VMAP_OK) {
      const std::string msg = h->err;
      clear_window_marks(h);
      cudaStreamSynchronize(h->stream);
      h->err = msg;
      return gs;
    }
    ++*retries;
    VM_CUDA(h, cudaMemsetAsync(&h->m.persist->stalled, 0, sizeof(uint32_t), h->stream));
    VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->overflow_table, 0, 2 * sizeof(uint32_t), h->stream));   // overflow_table, overflow_pool
    VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->stall, 0, 3 * sizeof(uint32_t), h->stream));            // stall, need_slots, done_ctas
  }
}

void fill_stats(vmap* h, uint64_t retries, float ms_total) {
  const Counters& c = *h->h_cnt;
  vmap_scan_stats& s = h->stats;
  s.points_valid = c.points_valid;
  s.points_in_range = c.points_in_range;
  s.rays_cast = c.rays_cast;
  s.traversals = c.traversals;
  s.occupied_emitted = c.occupied_emitted;
  s.unique_free = c.unique_free;
  s.unique_occupied = c.unique_occupied;
  s.longest_ray = c.longest_ray;
  s.blocks_touched = c.n_touched;
  s.blocks_total = h->h_persist->n_slots;
  s.retries = retries;
  s.gpu_ms = ms_total;
  auto span = [&](int a, int b) -> double {
    float v = 0.f;
    if (!h->evk_set[a] || !h->evk_set[b]) return 0.0;
    if (cudaEventElapsedTime(&v, h->evk[a], h->evk[b]) != cudaSuccess) { cudaGetLastError(); return 0.0; }
    return (double)v;
  };
  h->walk_kernel_ms = 0.f;
  if (h->evw_set && cudaEventElapsedTime(&h->walk_kernel_ms, h->evw[0], h->evw[1]) != cudaSuccess) { cudaGetLastError(); h->walk_kernel_ms = 0.f; }
  h->evw_set = false;
  s.front_ms = span(0, 1);
  s.resolve_ms = span(1, 2);
  s.walk_ms = h->evk_set[2] ? span(2, 3) : span(0, 3);
  s.update_ms = span(3, 4);
  h->total_scans++;
  vmap_scan_stats& t = h->totals;
  t.points_in += s.points_in; t.points_valid += s.points_valid; t.points_in_range += s.points_in_range;
  t.rays_cast += s.rays_cast; t.traversals += s.traversals; t.occupied_emitted += s.occupied_emitted;
  t.unique_free += s.unique_free; t.unique_occupied += s.unique_occupied;
  t.longest_ray = std::max(t.longest_ray, s.longest_ray);
  t.blocks_touched += s.blocks_touched; t.blocks_total = s.blocks_total; t.retries += s.retries;
  t.gpu_ms += s.gpu_ms; t.front_ms += s.front_ms; t.resolve_ms += s.resolve_ms; t.walk_ms += s.walk_ms;
  t.update_ms += s.update_ms;
}

// Runs  mark() -> [capture] -> update  with the grow-and-retry protocol, synchronously.
//   mark() enqueues everything up to and including the kernels that set block masks, into the
//   mark set h->ms (dense window when `origin` is given and a window is available, else the
//   table's mask columns).
template <class Mark>
int run_scan(vmap* h, size_t n_points, const float* origin, Mark&& mark) {
  h->walk_overlaps = false;
  VM_TRY(next_epoch(h));
  uint64_t retries = 0;
  float ms_total = 0.f;
  bool dense = origin && setup_dense_marks(h, origin);
  if (!dense) use_table_marks(h);
  for (;;) {
    VM_TRY(begin_scan_scratch(h, n_points));
    for (bool& b : h->evk_set) b = false;
    VM_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    VM_TRY(mark_stage(h, 0));
    VM_TRY(mark());
    bool redo = false;
    if (dense) {
      // (one host synchronisation in the common case: the tail is enqueued blind; on the device it
      // refuses itself when the marks or the list turn out to be incomplete)
      VM_TRY(launch_list(h));
      VM_TRY(launch_tail(h, 0));
      VM_TRY(read_counters(h));
      if (h->h_cnt->overflow_ckpt) {
        // not every segment got a checkpoint: drop the partial marks, enlarge, run the scan again
        VM_TRY(clear_window_marks(h));
        VM_TRY(grow_checkpoints(h));
        redo = true;
      } else if (h->h_cnt->overflow_window) {
        // a voxel outside the window (cannot happen for clipped rays; kept as a safety net):
        // drop the marks and run this scan through the table's mask columns instead
        VM_TRY(clear_window_marks(h));
        h->dense_fallbacks++;
        dense = false;
        use_table_marks(h);
        redo = true;
      } else if (h->h_cnt->overflow_keys) {
        // block list too small: the marks are complete and still in the window, only list again
        VM_TRY(grow_window_list(h, std::max<uint32_t>(h->h_cnt->n_touched + (h->h_cnt->n_touched >> 2), 2 * h->win.list_cap)));
        h->ms.list = h->win.list; h->ms.locate = h->win.locate; h->ms.list_cap = h->win.list_cap;
        VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->n_touched, 0, sizeof(uint32_t), h->stream));
        VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->overflow_keys, 0, sizeof(uint32_t), h->stream));
        VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->stall, 0, 3 * sizeof(uint32_t), h->stream));
        VM_TRY(launch_list(h));
        VM_TRY(launch_tail(h, 0));
        VM_TRY(read_counters(h));
        if (h->h_cnt->overflow_keys) return fail(h, VMAP_ERR_CAPACITY, "window block list could not be sized");
      }
      if (!redo) {
        VM_TRY(run_tail_sync(h, &retries, true));
        VM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        VM_CUDA(h, cudaEventSynchronize(h->ev1));
      }
    } else {
      VM_TRY(launch_tail(h, 0));
      VM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
      VM_TRY(read_counters(h));
      if (h->h_cnt->overflow_table) {
        // the table filled up mid-scan: drop this scan's marks, double the table, run it again
        VM_TRY(clear_marks(h));
        VM_TRY(grow_table(h));  // resets every entry's epoch to 0
        use_table_marks(h);
        redo = true;
      } else if (h->h_cnt->overflow_pool) {
        VM_TRY(clear_marks(h));   // (table mode: the marks are dropped before growing, so a failed grow leaves none)
        VM_TRY(grow_pool(h, std::min<uint64_t>((uint64_t)h->h_persist->n_entries, (uint64_t)h->h_persist->n_slots + h->h_cnt->n_touched)));
        // epochs are still the current one: give the re-run a fresh epoch so blocks are re-listed
        VM_TRY(next_epoch(h));
        redo = true;
      }
    }
    if (!redo) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->ev0, h->ev1);
      ms_total += ms;
      break;
    }
    if (++retries > 40) return fail(h, VMAP_ERR_CAPACITY, "scan could not be applied after repeated growth");
  }
  // keep the load factor <= 1/2 so probes stay short
  while ((uint64_t)h->h_persist->n_entries * 2 > h->table_cap) VM_TRY(grow_table(h));
  fill_stats(h, retries, ms_total);
  h->captured = h->capture;
  return VMAP_OK;
}

void fill_transform_matrix(Transform* T, const double M[16]) {
  std::memset(T, 0, sizeof(*T));
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) T->m[4 * r + c] = M[4 * r + c];
  // pointEigenToOctomap(T_G_sensor.getPosition())  (octomap_world.cc:120-121)
  T->origin[0] = (float)M[3];
  T->origin[1] = (float)M[7];
  T->origin[2] = (float)M[11];
}

void fill_transform_quat(Transform* T, const double q[4], const double t[3]) {
  std::memset(T, 0, sizeof(*T));
  for (int i = 0; i < 4; i++) T->q[i] = q[i];
  for (int i = 0; i < 3; i++) T->t[i] = t[i];
  // sensor_origin = T_G_sensor * Vector3d::Zero()  (octomap_world.cc:146-148): Eigen's
  // _transformVector on the zero vector, evaluated with the same operation order.
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double v[3] = {0.0, 0.0, 0.0};
  double uv[3] = {y * v[2] - z * v[1], z * v[0] - x * v[2], x * v[1] - y * v[0]};
  for (int i = 0; i < 3; i++) uv[i] += uv[i];
  const double c[3] = {y * uv[2] - z * uv[1], z * uv[0] - x * uv[2], x * uv[1] - y * uv[0]};
  for (int i = 0; i < 3; i++) T->origin[i] = (float)(((v[i] + w * uv[i]) + c[i]) + t[i]);
}

// K1b + K2 of the bitmask mode (everything after the front-end kernel).
int launch_resolve_and_walk(vmap* h, size_t n, const Transform& T) {
  VM_TRY(mark_stage(h, 1));
  // VMAP_HELPERS=1 (default): k_resolve_fold -> k_checkpoints_pts (no k_order_rays launch, no ordered ray
  // list); 0: round 1's chain k_resolve -> k_order_rays -> k_checkpoints
  const bool fold = h->helpers >= 1 && h->ms.dense && !h->no_segments;
  if (fold) {
    VM_CUDA(h, launch_chain(h, kPdlResolve, k_resolve_fold, grid_for(n, 256), 256, h->sb, h->m, h->ms, h->d_cnt, (uint32_t)n, 1));
    h->launches++;
  } else {
    k_resolve<<<grid_for(n, 256), 256, 0, h->stream>>>(h->sb, h->m, h->ms, h->d_cnt, (uint32_t)n, 1);
    k_order_rays<<<grid_for(n, 256), 256, 0, h->stream>>>(h->sb, (uint32_t)n);
    h->launches += 2;
  }
  VM_CUDA(h, cudaGetLastError());
  VM_TRY(mark_stage(h, 2));
  const float ox = T.origin[0], oy = T.origin[1], oz = T.origin[2];
  if (h->ms.dense && !h->no_segments) {
    // K2a + K2b: exact checkpoints, then one thread per ray segment
    if (fold) VM_CUDA(h, launch_chain(h, kPdlCkpt, k_checkpoints_pts, grid_for(n, 256), 256, h->P, h->sb, h->d_cnt, ox, oy, oz));
    else k_checkpoints<<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->d_cnt, ox, oy, oz);
    h->launches++;
    const unsigned grid = g_sm_count(h->device) * (unsigned)(h->walk_ctas > 0 ? h->walk_ctas : (h->walk_overlaps ? 4 : 5));
    // VMAP_WALK_ADDR: 0 block addressing, 1 dilated addressing with 32-voxel mask words (round 1), 5 the same
    // + near-field combining in shared memory, 7 dilated addressing with 64-voxel words, 9 (default) the same
    // with queued marks (walk_slot_queued), 8 queued marks built for 4 CTAs / SM.  (Measured and dropped: deferred touched test, 6 CTAs / SM, 64 registers:
    // profiles/r01_ab_walk_variants.json, r02_ab_walk2.json.)
    const bool dilated = h->walk_addr >= 1 && dilated_window_fits(h->ms.nx, h->ms.ny, h->ms.nz);
    h->evw_set = h->evk_set[2];      // stage marks are being recorded for this scan: time the walker's launch alone as well
    if (h->evw_set) VM_CUDA(h, cudaEventRecord(h->evw[0], h->stream));
    if (!h->traced && getenv("VMAP_TRACE")) {
      h->traced = true;
      fprintf(stderr, "[vmap] segment walker: %s addressing (mode %d), window %u x %u x %u blocks at (%d, %d, %d), fallbacks so far %llu\n",
              dilated ? "dilated" : "block", h->walk_addr, h->ms.nx, h->ms.ny, h->ms.nz, h->ms.x0, h->ms.y0, h->ms.z0,
              (unsigned long long)h->dense_fallbacks);
    }
#define VMAP_WALK_DIL(F, M, C, W) k_walk_segments_dilated<F, M, C, W><<<grid, 256, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant, h->walk_order, dilated_window(h->ms))
    // (with the free-space filter the keys are needed at every step anyway and the block-addressed
    // kernel measured faster: profiles/r01_ab_walk_dilated.json)
    if (dilated && !h->P.free_filter) {
      if (h->walk_addr == 5) k_walk_segments_near<<<grid, 256, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant, h->walk_order);
      else if (h->walk_addr == 1) VMAP_WALK_DIL(false, 0, 5, false);
      else if (h->walk_addr == 7) VMAP_WALK_DIL(false, 0, 5, true);
      else if (h->walk_addr == 8) k_walk_segments_queued<false, 4, false><<<grid, 256, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant, dilated_window(h->ms));
      else if (h->walk_addr == 10) k_walk_segments_queued<false, 5, true><<<grid, 256, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant, dilated_window(h->ms));
      else VM_CUDA(h, launch_chain(h, kPdlWalk, k_walk_segments_queued<false, 5, false>, grid, 256, h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant, dilated_window(h->ms)));
    }
#undef VMAP_WALK_DIL
    else if (h->P.free_filter) k_walk_segments<true><<<grid, 256, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant);
    else k_walk_segments<false><<<grid, 256, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant);
  } else if (h->ms.dense) {
    if (h->P.free_filter) k_walk_dense<true><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant);
    else k_walk_dense<false><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->ms, h->d_cnt, ox, oy, oz, h->walk_variant);
  } else {
    if (h->P.free_filter) k_walk_mask<true><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->m, h->d_cnt, ox, oy, oz);
    else k_walk_mask<false><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->m, h->d_cnt, ox, oy, oz);
  }
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  if (h->ms.dense && !h->no_segments && h->evw_set) VM_CUDA(h, cudaEventRecord(h->evw[1], h->stream));
  return VMAP_OK;
}

// VMAP_DEDUPE_SORT: K1b (no occupied marks) -> emit packed keys -> radix sort -> unique with
// occupied-wins -> set block masks from the unique keys (vmap_sort.cuh).
int launch_resolve_and_sort(vmap* h, size_t n, const Transform& T);

int launch_after_front(vmap* h, size_t n, const Transform& T) {
  if (h->cfg.dedupe_mode == VMAP_DEDUPE_SORT) return launch_resolve_and_sort(h, n, T);
  return launch_resolve_and_walk(h, n, T);
}

// ---------------------------------------------------------------------------------------------
// scan contexts and pipelined insertion
// ---------------------------------------------------------------------------------------------
void swap_ctx(vmap* h) { std::swap(static_cast<ScanCtx&>(*h), h->alt); }
void swap_ctx2(vmap* h) { std::swap(static_cast<ScanCtx&>(*h), h->alt2); }
// the context (other than the current one) whose scan in flight is the oldest / the newest; null: none in flight
ScanCtx* other_in_flight(vmap* h, bool newest) {
  ScanCtx* best = nullptr;
  for (ScanCtx* c : {&h->alt, &h->alt2})
    if (c->in_flight && (!best || (newest ? c->seq > best->seq : c->seq < best->seq))) best = c;
  return best;
}
void make_current(vmap* h, ScanCtx* c) {
  if (c == &h->alt) swap_ctx(h);
  else if (c == &h->alt2) swap_ctx2(h);
}

int init_ctx_resources(vmap* h) {   // for the CURRENT context
  int prio_lo = 0, prio_hi = 0;
  VM_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  if (!h->stream) VM_CUDA(h, cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi));
  if (!h->ev0) VM_CUDA(h, cudaEventCreate(&h->ev0));
  if (!h->ev1) VM_CUDA(h, cudaEventCreate(&h->ev1));
  for (auto& e : h->evk)
    if (!e) VM_CUDA(h, cudaEventCreate(&e));
  for (auto& e : h->evw)
    if (!e) VM_CUDA(h, cudaEventCreate(&e));
  if (!h->ev_done) VM_CUDA(h, cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  if (!h->ev_h2d) VM_CUDA(h, cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming));
  if (!h->h_cnt) {
    VM_CUDA(h, cudaMallocHost((void**)&h->h_cnt, sizeof(Counters)));
    std::memset(h->h_cnt, 0, sizeof(Counters));
  }
  if (!h->h_persist_snap) {
    VM_CUDA(h, cudaMallocHost((void**)&h->h_persist_snap, sizeof(Persist)));
    std::memset(h->h_persist_snap, 0, sizeof(Persist));
  }
  // per-scan counters and the segment-count histogram share one allocation (one memset per scan)
  if (!h->d_cnt) VM_TRY(dev_alloc(h, (void**)&h->d_cnt, sizeof(Counters) + kFineBins * sizeof(uint32_t)));
  return VMAP_OK;
}

void free_ctx_resources(vmap* h) {  // of the CURRENT context
  ScanBuffers& sb = h->sb;
  cudaFree(sb.world); cudaFree(sb.ray_end); cudaFree(sb.key); cudaFree(sb.flags); cudaFree(sb.rays);
  cudaFree(sb.len); cudaFree(sb.rank);
  cudaFree(sb.seg_base); cudaFree(sb.bin_off); cudaFree(sb.ray_rec); cudaFree(sb.ckpt);
  cudaFree(sb.first_keys);
  sb = ScanBuffers{};
  cudaFree(h->in.p);
  h->in = DevBuf{};
  cudaFree(h->d_cnt);
  h->d_cnt = nullptr;
  free_window(h);
  if (h->h_cnt) cudaFreeHost(h->h_cnt);
  if (h->h_persist_snap) cudaFreeHost(h->h_persist_snap);
  h->h_cnt = nullptr;
  h->h_persist_snap = nullptr;
  cudaEvent_t* evs[] = {&h->ev0, &h->ev1, &h->ev_done, &h->ev_h2d, &h->evk[0], &h->evk[1], &h->evk[2], &h->evk[3], &h->evk[4], &h->evw[0], &h->evw[1]};
  for (cudaEvent_t* e : evs)
    if (*e) { cudaEventDestroy(*e); *e = nullptr; }
  if (h->stream) cudaStreamDestroy(h->stream);
  h->stream = nullptr;
}

// The classic, synchronous insertion of the cloud staged at d_in (device memory).
int insert_cloud_sync(vmap* h, const uint8_t* d_in, size_t n, size_t stride, int is_dense, const double Tm[16]) {
  Transform T;
  fill_transform_matrix(&T, Tm);
  VM_TRY(ensure_scan(h, n));
  h->stats = vmap_scan_stats{};
  h->stats.points_in = n;
  return run_scan(h, n, T.origin, [&]() -> int {
    k_front_cloud<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, d_in, (uint32_t)n,
                                                          (uint32_t)stride, is_dense);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
    return launch_after_front(h, n, T);
  });
}

// Upper bound of the segment checkpoints n rays can need: a clipped ray is at most max_range long,
// its DDA takes at most the Manhattan distance of its end keys (+3) steps, one segment per
// kSegSteps steps.
uint64_t checkpoint_bound(const vmap* h, size_t n) {
  const double r_vox = h->P.max_range * h->P.res_factor + 2.0;
  const double per_ray = std::floor((1.7320508075688772 * r_vox + 3.0) / (double)kSegSteps) + 2.0;
  const double total = per_ray * (double)std::max<size_t>(n, 1);
  return total < 1.8e19 ? (uint64_t)total : ~0ull;
}

int ensure_ckpt(vmap* h, uint64_t want) {
  ScanBuffers& sb = h->sb;
  if (want <= sb.ckpt_cap) return VMAP_OK;
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  dev_free(h, sb.ckpt, (size_t)sb.ckpt_cap * sizeof(Checkpoint));
  sb.ckpt = nullptr;
  sb.ckpt_cap = 0;
  VM_TRY(dev_alloc(h, (void**)&sb.ckpt, (size_t)want * sizeof(Checkpoint)));
  sb.ckpt_cap = (uint32_t)want;
  h->ckpt_want = (size_t)want;
  return VMAP_OK;
}

// May this scan go through the pipeline?  (Everything else -- key capture, sort mode, unlimited
// range, windows over the budget, sharded handles -- takes the synchronous path.)
bool pipe_eligible(const vmap* h, size_t n) {
  if (h->pipe_depth <= 0 || h->capture || h->cfg.dedupe_mode != VMAP_DEDUPE_BITMASK) return false;
  if (h->no_segments || h->shard_world > 1 || h->win.disabled || h->alt.win.disabled || h->alt2.win.disabled) return false;
  if (!(h->P.max_range >= 0.0) || !std::isfinite(h->P.max_range)) return false;
  const uint64_t ck = checkpoint_bound(h, n);
  return ck < 0xFFFFFFFFull && ck * sizeof(Checkpoint) <= (2ull << 30);
}

int pipe_retire(vmap* h);
int peer_retire(vmap* h);   // vmap_peer_host.inl

// Drain the pipeline.  Returns the first deferred failure (the failing scan was dropped, later
// scans were still applied).
int pipe_flush(vmap* h) {
  int first = VMAP_OK;
  std::string msg;
  while (h->pipe_n > 0) {
    const int st = pipe_retire(h);
    if (st != VMAP_OK && first == VMAP_OK) { first = st; msg = h->err; }
  }
  if (first != VMAP_OK) h->err = msg;
  return first;
}

// Quiescent-point work before a pipelined insert: the second context, and map growth ahead of need
// (a scan that finds too little room is refused on the device and replayed -- correct but slow).
int pipe_prepare(vmap* h, bool third = false) {
  if (!h->alt_ready) {
    swap_ctx(h);
    h->ctx_id = 1;
    const int st = init_ctx_resources(h);
    swap_ctx(h);
    VM_TRY(st);
    h->alt_ready = true;
  }
  if (third && !h->alt2_ready) {
    swap_ctx2(h);
    h->ctx_id = 2;
    const int st = init_ctx_resources(h);
    swap_ctx2(h);
    VM_TRY(st);
    h->alt2_ready = true;
  }
  if (!h->copy_stream) {
    VM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& e : h->ev_stage) VM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (!h->ev_wb) VM_CUDA(h, cudaEventCreateWithFlags(&h->ev_wb, cudaEventDisableTiming));
  if (h->pipe_grow) {
    VM_TRY(pipe_flush(h));
    h->pipe_grow = false;
    const uint64_t t = h->pipe_touch_max;
    while (((uint64_t)h->h_persist->n_entries + 8 * t) * 2 > h->table_cap) {
      if (grow_table(h) != VMAP_OK) { cudaGetLastError(); break; }       // (not fatal here: the refusal path reports real shortages)
    }
    uint64_t want = (uint64_t)h->h_persist->n_slots + 8 * t;
    if (h->cfg.max_blocks) want = std::min<uint64_t>(want, h->cfg.max_blocks);
    if (want > h->pool_cap && grow_pool(h, want) != VMAP_OK) cudaGetLastError();
  }
  return VMAP_OK;
}

// After a pipelined scan was applied: is the headroom for the next ones getting thin?
void pipe_watch_headroom(vmap* h, uint32_t n_touched) {
  h->pipe_touch_max = std::max(h->pipe_touch_max, n_touched);
  const uint64_t t = h->pipe_touch_max;
  // (up to three scans are in flight behind the counters this is decided on)
  if (((uint64_t)h->h_persist->n_entries + 4 * t) * 2 > h->table_cap) h->pipe_grow = true;
  const bool pool_can_grow = !h->cfg.max_blocks || h->pool_cap < h->cfg.max_blocks;
  if (pool_can_grow && (uint64_t)h->h_persist->n_slots + 4 * t > h->pool_cap) h->pipe_grow = true;
}

// The current context holds a scan whose tail the device refused; nothing else is in flight.
int pipe_replay(vmap* h, bool first) {
  const Counters c = *h->h_cnt;
  if (first) *h->h_persist = *h->h_persist_snap;   // (a younger scan's snapshot is older than what the replay before it left)
  VM_CUDA(h, cudaMemsetAsync(&h->m.persist->stalled, 0, sizeof(uint32_t), h->stream));
  if (c.overflow_window || c.overflow_ckpt || c.overflow_keys) {
    // the marks are incomplete: drop them and run the whole scan again from its staged input
    VM_TRY(clear_window_marks(h));
    VM_CUDA(h, cudaStreamSynchronize(h->stream));
    return insert_cloud_sync(h, (const uint8_t*)h->in.p, h->meta.n, h->meta.stride, h->meta.is_dense, h->meta.Tm);
  }
  // marks, list and touched bitmap are complete: make room, then run the tail until it is admitted
  while (((uint64_t)h->h_persist->n_entries + c.n_touched) * 2 > h->table_cap) {
    const int gs = grow_table(h);
    if (gs != VMAP_OK) {
      const std::string msg = h->err;
      clear_window_marks(h);
      cudaStreamSynchronize(h->stream);
      h->err = msg;
      return gs;
    }
  }
  VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->overflow_table, 0, 2 * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->stall, 0, 3 * sizeof(uint32_t), h->stream));
  uint64_t retries = 1;
  VM_TRY(run_tail_sync(h, &retries));
  VM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  VM_CUDA(h, cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->stats = vmap_scan_stats{};
  h->stats.points_in = h->meta.n;
  fill_stats(h, retries, ms);
  pipe_watch_headroom(h, h->h_cnt->n_touched);
  return VMAP_OK;
}

// Retire the OLDEST scan in flight: wait for it, take its counters, and -- when the device refused
// its update -- replay it (and the younger scan behind it, which refused itself too).
int pipe_retire(vmap* h) {
  if (h->pipe_n <= 0) return VMAP_OK;
  if (h->peer_on) return peer_retire(h);
  {
    ScanCtx* oldest = other_in_flight(h, false);
    if (oldest && (!h->in_flight || oldest->seq < h->seq)) make_current(h, oldest);
  }
  VM_CUDA(h, cudaEventSynchronize(h->ev_done));
  h->in_flight = false;
  h->pipe_n--;
  const Counters& c = *h->h_cnt;
  if (!(c.stall || c.overflow_table || c.overflow_pool || c.overflow_window || c.overflow_ckpt || c.overflow_keys)) {
    *h->h_persist = *h->h_persist_snap;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats = vmap_scan_stats{};
    h->stats.points_in = h->meta.n;
    fill_stats(h, 0, ms);
    pipe_watch_headroom(h, c.n_touched);
    return VMAP_OK;
  }
  // refused on the device: every younger scan in flight has refused itself too (Persist::stalled);
  // wait for them all, then replay in scan order
  h->pipe_stalls++;
  for (ScanCtx* y : {&h->alt, &h->alt2})
    if (y->in_flight) VM_CUDA(h, cudaEventSynchronize(y->ev_done));
  int st = pipe_replay(h, true);
  while (ScanCtx* y = other_in_flight(h, false)) {
    make_current(h, y);
    h->in_flight = false;
    h->pipe_n--;
    const std::string msg = h->err;
    const int st2 = pipe_replay(h, false);
    if (st != VMAP_OK) h->err = msg;
    else st = st2;
  }
  return st;
}

// Benchmark hook (vmap_set_debug_l2_flush): overwrite a buffer larger than the L2 so that what follows finds
// nothing of the map in it.  (A kernel of our own with one light CTA per SM, meant to run beside another scan's
// walk, measured SLOWER than the memset node: 0.224 vs 0.205 ms per pipelined scan, profiles/README.md.)
int enqueue_l2_flush(vmap* h, cudaStream_t st, uint64_t tag) {
  if (!h->l2_flush_bytes) return VMAP_OK;
  VM_CUDA(h, cudaMemsetAsync(h->l2_flush.p, (int)(tag & 0xFF), h->l2_flush_bytes, st));
  return VMAP_OK;
}

__global__ void __launch_bounds__(256)
k_writeback(const uint8_t* __restrict__ in, const float* __restrict__ world, uint8_t* __restrict__ out, uint32_t n_words,
            uint32_t stride) {
  // the caller's cloud with x, y, z replaced by the world-frame point (every other byte as it came)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) {
    const uint32_t byte = i * 4u, point = byte / stride, off = byte - point * stride;
    reinterpret_cast<uint32_t*>(out)[i] = off < 12u ? __float_as_uint(world[3u * point + (off >> 2)])
                                                   : reinterpret_cast<const uint32_t*>(in)[i];
  }
}

// The reference leaves the NaN-compacted, world-frame cloud in the caller's buffer
// (octomap_world.cc:115-119).  `st`: a stream on which the front-end kernel's outputs are ready;
// d_in: the staged input.  Blocks until the caller's buffer is rewritten.
int write_back_cloud(vmap* h, cudaStream_t st, const uint8_t* d_in, size_t n, size_t stride, int is_dense,
                     float* out, size_t* n_out) {
  if (n_out) *n_out = n;
  if (!n || (!out && (is_dense || !n_out))) return VMAP_OK;
  if (is_dense) {
    if (stride == 12) {
      VM_CUDA(h, cudaMemcpyAsync(out, h->sb.world, n * 12, cudaMemcpyDeviceToHost, st));
    } else {
      const size_t bytes = (n - 1) * stride + 12;
      if (bytes > h->wb.cap) {   // (only ever used from here, and every use is followed by a synchronisation)
        dev_free(h, h->wb.p, h->wb.cap);
        h->wb = DevBuf{};
        VM_TRY(dev_alloc(h, &h->wb.p, bytes + (bytes >> 2)));
        h->wb.cap = bytes + (bytes >> 2);
      }
      k_writeback<<<g_sm_count(h->device) * 4, 256, 0, st>>>(d_in, h->sb.world, (uint8_t*)h->wb.p, (uint32_t)(bytes / 4), (uint32_t)stride);
      h->launches++;
      VM_CUDA(h, cudaGetLastError());
      VM_CUDA(h, cudaMemcpyAsync(out, h->wb.p, bytes, cudaMemcpyDeviceToHost, st));
    }
    VM_CUDA(h, cudaStreamSynchronize(st));
    return VMAP_OK;
  }
  // removeNaNFromPointCloud: order-preserving compaction, done on the host from pinned staging
  const size_t need = n * 13;
  if (need > h->wb_host_cap) {
    if (h->wb_host) cudaFreeHost(h->wb_host);
    h->wb_host = nullptr;
    h->wb_host_cap = 0;
    VM_CUDA(h, cudaMallocHost((void**)&h->wb_host, need + (need >> 2)));
    h->wb_host_cap = need + (need >> 2);
  }
  float* w = reinterpret_cast<float*>(h->wb_host);
  uint8_t* f = h->wb_host + n * 12;
  VM_CUDA(h, cudaMemcpyAsync(w, h->sb.world, n * 12, cudaMemcpyDeviceToHost, st));
  VM_CUDA(h, cudaMemcpyAsync(f, h->sb.flags, n, cudaMemcpyDeviceToHost, st));
  VM_CUDA(h, cudaStreamSynchronize(st));
  size_t j = 0;
  for (size_t i = 0; i < n; i++) {
    if (!(f[i] & kFlagValid)) continue;
    if (out) {
      float* o = reinterpret_cast<float*>(reinterpret_cast<char*>(out) + j * stride);
      o[0] = w[3 * i]; o[1] = w[3 * i + 1]; o[2] = w[3 * i + 2];
    }
    j++;
  }
  if (n_out) *n_out = j;
  return VMAP_OK;
}

// insertPointcloudIntoMapImpl.  src: the caller's points, in host memory (src_on_host) or on the
// device.  Pipelined (default): the scan is staged into the free scan context, all of its kernels
// are enqueued on that context's stream, and the call returns once the caller's buffer has been
// consumed; up to pipe_depth scans stay in flight, their map updates ordered by events.
int insert_cloud_common(vmap* h, const void* src, bool src_on_host, size_t n, size_t stride, int is_dense,
                        const double Tm[16], float* wb_out, size_t* wb_n_out) {
  VM_TRY(flush_rounds(h));
  if (wb_n_out) *wb_n_out = n;
  if (n == 0) {
    // updateOccupancy on two empty sets: nothing changes
    VM_TRY(pipe_flush(h));
    h->stats = vmap_scan_stats{};
    h->h_cap_n[0] = h->h_cap_n[1] = 0;
    h->captured = h->capture;
    return VMAP_OK;
  }
  const size_t bytes = (n - 1) * stride + 12;
  bool piped = pipe_eligible(h, n);
  Transform T;
  fill_transform_matrix(&T, Tm);
  if (piped) {
    const int n_ctx = h->pipe_depth >= 2 && h->pipe_ctx >= 3 ? 3 : 2;
    VM_TRY(pipe_prepare(h, n_ctx == 3));
    while (h->pipe_n >= n_ctx) VM_TRY(pipe_retire(h));
    if (h->in_flight) {                       // the current context is a free one from here on
      if (!h->alt.in_flight) swap_ctx(h);
      else swap_ctx2(h);
    }
    if (h->epoch >= kEpochMax) VM_TRY(pipe_flush(h));
    VM_TRY(ensure_scan(h, n));
    VM_TRY(ensure_ckpt(h, checkpoint_bound(h, n)));
    piped = setup_dense_marks(h, T.origin);
    if (piped && h->win.list_cap < h->win.n_blocks) {
      // a list that can hold every block of the window: the list can never overflow
      piped = h->win.n_blocks * 8 <= (1ull << 30) && grow_window_list(h, (uint32_t)h->win.n_blocks) == VMAP_OK &&
              h->win.list_cap >= h->win.n_blocks;
      h->ms.list = h->win.list; h->ms.locate = h->win.locate; h->ms.list_cap = h->win.list_cap;
    }
  }
  if (!piped) {
    VM_TRY(pipe_flush(h));
    const uint8_t* d_in = reinterpret_cast<const uint8_t*>(src);
    if (src_on_host) {
      VM_TRY(ensure(h, &h->in, bytes));
      VM_CUDA(h, cudaMemcpyAsync(h->in.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
      d_in = (const uint8_t*)h->in.p;
    }
    VM_TRY(enqueue_l2_flush(h, h->stream, h->total_scans));
    VM_TRY(insert_cloud_sync(h, d_in, n, stride, is_dense, Tm));
    return write_back_cloud(h, h->stream, d_in, n, stride, is_dense, wb_out, wb_n_out);
  }
  // ---- pipelined ----
  h->walk_overlaps = true;
  VM_TRY(ensure(h, &h->in, bytes));
  // benchmark hook: no scan finds its map blocks in L2.  VMAP_FLUSH_AT=0: at the start of the scan (round 2's
  // placement -- the update of the scan before may still run after it); 1 (default): in the ordered half, between
  // the previous scan's update and this one's
  if (h->flush_at == 0) VM_TRY(enqueue_l2_flush(h, h->stream, h->pipe_seq));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, src, bytes, src_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
  VM_CUDA(h, cudaEventRecord(h->ev_h2d, h->stream));
  VM_TRY(next_epoch(h));
  VM_TRY(begin_scan_scratch(h, n));
  for (bool& b : h->evk_set) b = false;
  VM_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  VM_TRY(mark_stage(h, 0));
  k_front_cloud<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, (const uint8_t*)h->in.p, (uint32_t)n,
                                                        (uint32_t)stride, is_dense);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  const bool wb = wb_out || (wb_n_out && !is_dense);
  if (wb) VM_CUDA(h, cudaEventRecord(h->ev_wb, h->stream));
  VM_TRY(launch_after_front(h, n, T));
  VM_TRY(launch_list(h));
  // the ordered half: after the update of the scan before this one
  if (ScanCtx* prev = other_in_flight(h, true)) VM_CUDA(h, cudaStreamWaitEvent(h->stream, prev->ev_done, 0));
  if (h->flush_at != 0) VM_TRY(enqueue_l2_flush(h, h->stream, h->pipe_seq));
  VM_TRY(launch_tail(h, 1));
  VM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(h->h_persist_snap, h->m.persist, sizeof(Persist), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaEventRecord(h->ev_done, h->stream));
  h->in_flight = true;
  h->seq = ++h->pipe_seq;
  h->pipe_n++;
  h->meta.n = n;
  h->meta.stride = stride;
  h->meta.is_dense = is_dense;
  std::memcpy(h->meta.Tm, Tm, sizeof(h->meta.Tm));
  // the caller may reuse its buffer as soon as this call returns
  VM_CUDA(h, cudaEventSynchronize(h->ev_h2d));
  if (wb) {
    VM_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_wb, 0));
    VM_TRY(write_back_cloud(h, h->copy_stream, (const uint8_t*)h->in.p, n, stride, is_dense, wb_out, wb_n_out));
  }
  while (h->pipe_n > h->pipe_depth) VM_TRY(pipe_retire(h));
  return VMAP_OK;
}

}  // namespace

#include "vmap_sort_host.inl"
#include "vmap_shard_host.inl"
#include "vmap_peer_host.inl"

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int vmap_dist_shutdown(vmap* h);

const char* vmap_version(void) { return "volumetric_mapping_b200 0.1 (sm_100a)"; }

void vmap_default_params(vmap_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->resolution = 0.15;
  p->probability_hit = 0.65;
  p->probability_miss = 0.4;
  p->threshold_min = 0.12;
  p->threshold_max = 0.97;
  p->threshold_occupancy = 0.7;
  p->filter_speckles = 1;
  p->max_free_space = 0.0;
  p->min_height_free_space = 0.0;
  p->sensor_max_range = 5.0;
  p->visualize_min_z = -1.7976931348623157e308;  // -std::numeric_limits<double>::max()
  p->visualize_max_z = 1.7976931348623157e308;
  p->treat_unknown_as_occupied = 1;
  p->change_detection_enabled = 0;
}

void vmap_default_config(vmap_config* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->device = -1;
  c->dedupe_mode = VMAP_DEDUPE_BITMASK;
}

const char* vmap_last_error(const vmap* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vmap_create(const vmap_params* params, const vmap_config* config, vmap** out) try {
  if (!out) {
    g_create_error = "vmap_create: out is NULL";
    return VMAP_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  vmap_params p;
  if (params) p = *params; else vmap_default_params(&p);
  vmap_config c;
  if (config) c = *config; else vmap_default_config(&c);
  if (!(p.resolution > 0.0) || !std::isfinite(p.resolution)) {
    g_create_error = "vmap_create: resolution must be positive";
    return VMAP_ERR_INVALID_ARGUMENT;
  }
  if (c.dedupe_mode != VMAP_DEDUPE_BITMASK && c.dedupe_mode != VMAP_DEDUPE_SORT) {
    g_create_error = "vmap_create: unknown dedupe_mode";
    return VMAP_ERR_INVALID_ARGUMENT;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    g_create_error = std::string("vmap_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); this library has no CPU fallback";
    cudaGetLastError();
    return VMAP_ERR_NO_DEVICE;
  }
  int dev = c.device;
  if (dev < 0) {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  }
  if (dev >= n_dev) {
    g_create_error = "vmap_create: device ordinal out of range";
    return VMAP_ERR_INVALID_ARGUMENT;
  }
  vmap* h = new (std::nothrow) vmap();
  if (!h) {
    g_create_error = "vmap_create: out of host memory";
    return VMAP_ERR_CAPACITY;
  }
  h->params = p;
  h->cfg = c;
  h->device = dev;
  derive_params(h);
  if (const char* e = getenv("VMAP_WALK_VARIANT")) h->walk_variant = atoi(e);
  if (const char* e = getenv("VMAP_WALK_ADDR")) h->walk_addr = atoi(e);
  if (const char* e = getenv("VMAP_WALK_CTAS")) h->walk_ctas = std::max(1, atoi(e));
  if (const char* e = getenv("VMAP_WALK_ORDER")) h->walk_order = (h->walk_order & ~1) | (atoi(e) & 1);
  if (const char* e = getenv("VMAP_WALK_COUNTED")) h->walk_order = (h->walk_order & ~2) | (atoi(e) ? 2 : 0);
  if (const char* e = getenv("VMAP_WALK_FAR_BLIND")) h->walk_order = (h->walk_order & ~4) | (atoi(e) ? 4 : 0);
  if (const char* e = getenv("VMAP_NO_SEGMENTS")) h->no_segments = atoi(e);
  if (const char* e = getenv("VMAP_UPDATE_VARIANT")) h->update_variant = atoi(e);
  if (const char* e = getenv("VMAP_HELPERS")) h->helpers = atoi(e);
  if (const char* e = getenv("VMAP_FLUSH_AT")) h->flush_at = atoi(e);
  if (const char* e = getenv("VMAP_PDL")) h->pdl = atoi(e);
  if (const char* e = getenv("VMAP_STAGE_MARKS")) h->stage_marks = atoi(e) != 0;
  auto bail = [&](int code) {
    g_create_error = h->err;
    vmap_destroy(h);
    return code;
  };
  auto cuda_ok = [&](cudaError_t ce, const char* what) {
    if (ce == cudaSuccess) return true;
    h->err = std::string(what) + ": " + cudaGetErrorString(ce);
    return false;
  };
  if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return bail(VMAP_ERR_CUDA);
  if (const char* e = getenv("VMAP_PIPE_DEPTH")) h->pipe_depth = std::min(2, std::max(0, atoi(e)));
  h->alt.ctx_id = 1;
  h->alt2.ctx_id = 2;
  if (const char* e = getenv("VMAP_PIPE_CTX")) h->pipe_ctx = std::min(3, std::max(2, atoi(e)));
  if (init_ctx_resources(h) != VMAP_OK) return bail(VMAP_ERR_CUDA);
  if (!cuda_ok(cudaEventCreate(&h->evr0), "cudaEventCreate")) return bail(VMAP_ERR_CUDA);
  if (!cuda_ok(cudaEventCreate(&h->evr1), "cudaEventCreate")) return bail(VMAP_ERR_CUDA);
  if (!cuda_ok(cudaMallocHost((void**)&h->h_persist, sizeof(Persist)), "cudaMallocHost")) return bail(VMAP_ERR_CUDA);
  std::memset(h->h_persist, 0, sizeof(Persist));
  int s = VMAP_OK;
  if (s == VMAP_OK) s = dev_alloc(h, (void**)&h->d_cap_n, 2 * sizeof(uint32_t));
  if (s == VMAP_OK) s = dev_alloc(h, (void**)&h->d_pack_counts_base, (size_t)2 * kMaxBatch * 2 * kMaxRanks * sizeof(uint32_t));
  h->d_pack_counts = h->d_pack_counts_base;
  if (s == VMAP_OK) s = init_map(h);
  if (s == VMAP_OK) s = set_change_detection(h, p.change_detection_enabled != 0);
  if (s != VMAP_OK) return bail(s);
  *out = h;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(nullptr); }

int vmap_destroy(vmap* h) try {
  if (!h) return VMAP_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();      // pipelined scans, apply rounds: nothing may be in flight
  h->pipe_n = 0;
  h->in_flight = h->alt.in_flight = h->alt2.in_flight = false;
  free_map(h);
  cudaFree(h->out.p); cudaFree(h->aux.p);
  cudaFree(h->cap_free.p); cudaFree(h->cap_occ.p);
  cudaFree(h->d_cap_n);
  vmap_dist_shutdown(h);
  cudaFree(h->pack_inflight.p); cudaFree(h->l2_flush.p); cudaFree(h->d_cnt_app);
  if (h->h_cnt_app) cudaFreeHost(h->h_cnt_app);
  if (h->h_persist_app) cudaFreeHost(h->h_persist_app);
  if (h->ev_app0) cudaEventDestroy(h->ev_app0);
  if (h->ev_app1) cudaEventDestroy(h->ev_app1);
  if (h->app_stream) cudaStreamDestroy(h->app_stream);
  cudaFree(h->pack.p); cudaFree(h->recv.p); cudaFree(h->d_pack_counts_base); cudaFree(h->d_count_matrix);
  sort_free(h);
  cudaFree(h->chg0_store); cudaFree(h->chg1_store);
  cudaFree(h->stage[0].p); cudaFree(h->stage[1].p);
  cudaFree(h->wb.p);
  if (h->wb_host) cudaFreeHost(h->wb_host);
  for (auto& e : h->ev_stage)
    if (e) cudaEventDestroy(e);
  if (h->ev_wb) cudaEventDestroy(h->ev_wb);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->h_persist) cudaFreeHost(h->h_persist);
  if (h->evr0) cudaEventDestroy(h->evr0);
  if (h->evr1) cudaEventDestroy(h->evr1);
  free_ctx_resources(h);
  swap_ctx(h);
  free_ctx_resources(h);
  swap_ctx2(h);
  free_ctx_resources(h);
  delete h;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_reset(vmap* h) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  // octree_->clear(): keep the allocations, forget every block
  VM_CUDA(h, cudaMemsetAsync(h->m.keys, 0xFF, h->table_cap * sizeof(unsigned long long), h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->m.slot, 0xFF, h->table_cap * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->m.free_mask, 0, h->table_cap * kMaskWords * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->m.occ_mask, 0, h->table_cap * kMaskWords * sizeof(uint32_t), h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->m.pool, 0xFF, h->pool_cap * kBlockVoxels * sizeof(float), h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->m.persist, 0, sizeof(Persist), h->stream));
  if (h->chg0_store) {
    VM_CUDA(h, cudaMemsetAsync(h->chg0_store, 0, h->pool_cap * kMaskWords * sizeof(uint32_t), h->stream));
    VM_CUDA(h, cudaMemsetAsync(h->chg1_store, 0, h->pool_cap * kMaskWords * sizeof(uint32_t), h->stream));
  }
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  h->epoch = 0;
  h->stats = vmap_scan_stats{};
  h->totals = vmap_scan_stats{};
  h->total_scans = 0;
  h->pipe_touch_max = 0;
  h->pipe_grow = false;
  std::memset(h->h_persist, 0, sizeof(Persist));
  h->captured = false;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_set_params(vmap* h, const vmap_params* params) try {
  if (!h || !params) return VMAP_ERR_INVALID_ARGUMENT;
  if (!(params->resolution > 0.0) || !std::isfinite(params->resolution))
    return fail(h, VMAP_ERR_INVALID_ARGUMENT, "resolution must be positive");
  const bool new_res = params->resolution != h->params.resolution;  // octomap_world.cc:85
  h->params = *params;
  derive_params(h);
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(set_change_detection(h, params->change_detection_enabled != 0));   // octomap_world.cc:99
  if (new_res) return vmap_reset(h);
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_get_params(const vmap* h, vmap_params* params) try {
  if (!h || !params) return VMAP_ERR_INVALID_ARGUMENT;
  *params = h->params;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(const_cast<vmap*>(h)); }

// ---------------------------------------------------------------------------------------------
// insertion
// ---------------------------------------------------------------------------------------------
int vmap_insert_pointcloud_dev(vmap* h, const float* d_xyz, size_t n, size_t stride_bytes,
                               const double T_rowmajor[16]) try {
  if (!h || (!d_xyz && n) || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  VM_CUDA(h, cudaSetDevice(h->device));
  return insert_cloud_common(h, d_xyz, false, n, stride_bytes, 1, T_rowmajor, nullptr, nullptr);
} catch (...) { return vmap_on_exception(h); }

int vmap_insert_pointcloud(vmap* h, const float* xyz, size_t n, size_t stride_bytes, int is_dense,
                           const double T_rowmajor[16], float* xyz_world_out, size_t* n_out) try {
  if (!h || (!xyz && n) || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  VM_CUDA(h, cudaSetDevice(h->device));
  return insert_cloud_common(h, xyz, true, n, stride_bytes, is_dense ? 1 : 0, T_rowmajor, xyz_world_out, n_out);
} catch (...) { return vmap_on_exception(h); }

// The scan pipeline from the outside.
int vmap_flush(vmap* h) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  return flush_app(h);
} catch (...) { return vmap_on_exception(h); }

int vmap_set_pipeline_depth(vmap* h, int depth) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (depth < 0 || depth > 2) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "pipeline depth must be 0 (synchronous), 1 or 2");
  VM_CUDA(h, cudaSetDevice(h->device));
  const int st = flush_app(h);
  h->pipe_depth = depth;
  return st;
} catch (...) { return vmap_on_exception(h); }

int vmap_totals(vmap* h, vmap_scan_stats* sums, uint64_t* n_scans, uint64_t* n_replayed) try {
  if (!h || !sums) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  const int st = flush_app(h);
  *sums = h->totals;
  if (n_scans) *n_scans = h->total_scans;
  if (n_replayed) *n_replayed = h->pipe_stalls;
  return st;
} catch (...) { return vmap_on_exception(h); }

// Timed region on the device: begin drains everything and records an event; end drains again,
// records a second one and returns the time between them.
int vmap_region_begin(vmap* h) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  const int st = flush_app(h);
  VM_CUDA(h, cudaDeviceSynchronize());
  VM_CUDA(h, cudaEventRecord(h->evr0, h->stream));
  return st;
} catch (...) { return vmap_on_exception(h); }

int vmap_region_end(vmap* h, double* ms) try {
  if (!h || !ms) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  const int st = flush_app(h);
  VM_CUDA(h, cudaDeviceSynchronize());
  VM_CUDA(h, cudaEventRecord(h->evr1, h->stream));
  VM_CUDA(h, cudaEventSynchronize(h->evr1));
  float v = 0.f;
  VM_CUDA(h, cudaEventElapsedTime(&v, h->evr0, h->evr1));
  *ms = (double)v;
  return st;
} catch (...) { return vmap_on_exception(h); }

static int insert_image_common(vmap* h, const uint8_t* d_img, int rows, int cols, size_t pitch,
                               const double* Q /* null: image is CV_32FC3 */, const double q[4],
                               const double t[3]) {
  VM_TRY(flush_app(h));
  Transform T;
  fill_transform_quat(&T, q, t);
  const size_t n = (size_t)rows * (size_t)cols;
  VM_TRY(ensure_scan(h, n));
  h->stats = vmap_scan_stats{};
  h->stats.points_in = n;
  if (n == 0) {
    h->h_cap_n[0] = h->h_cap_n[1] = 0;
    h->captured = h->capture;
    return VMAP_OK;
  }
  QMat Qm;
  if (Q) std::memcpy(Qm.q, Q, sizeof(Qm.q)); else std::memset(Qm.q, 0, sizeof(Qm.q));
  return run_scan(h, n, T.origin, [&]() -> int {
    if (Q) {
      VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->min_disparity_ord, 0xFF, sizeof(uint32_t), h->stream));
      k_min_disparity<<<g_sm_count(h->device) * 4, 256, 0, h->stream>>>(h->d_cnt, d_img, rows, cols, pitch);
      h->launches++;
      VM_CUDA(h, cudaGetLastError());
      k_front_disparity<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, Qm, d_img, rows, cols, pitch, nullptr);
    } else {
      k_front_projected<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, d_img, rows, cols, pitch);
    }
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
    return launch_after_front(h, n, T);
  });
}

static int check_image_args(vmap* h, const void* img, int rows, int cols, size_t pitch, size_t px_bytes,
                            const double* q, const double* t) {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (rows < 0 || cols < 0 || !q || !t) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bad image arguments");
  if ((size_t)rows * (size_t)cols && !img) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "null image");
  if (pitch < (size_t)cols * px_bytes || (pitch & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "row pitch too small or not a multiple of 4");
  return VMAP_OK;
}

// Pipelined form of vmap_insert_pointcloud for callers that stream scans from host memory: stage
// scan k+1 (asynchronous host->device copy on a stream of its own; the host buffer -- ideally
// pinned -- must stay untouched until the matching vmap_insert_staged returns) while scan k is
// being inserted.  Scans are inserted in the order they were staged; at most two may be staged.
int vmap_stage_pointcloud(vmap* h, const float* xyz, size_t n, size_t stride_bytes) try {
  if (!h || (!xyz && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many points");
  const int k = h->stage_next;
  if (h->staged[k]) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "two scans are already staged: call vmap_insert_staged first");
  VM_CUDA(h, cudaSetDevice(h->device));
  if (!h->copy_stream) {
    VM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& e : h->ev_stage) VM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const size_t bytes = n ? (n - 1) * stride_bytes + 12 : 0;
  if (bytes > h->stage[k].cap) {
    // (the slot is free: nothing in flight reads it; ensure() also drains the main stream)
    VM_CUDA(h, cudaStreamSynchronize(h->copy_stream));
    VM_TRY(ensure(h, &h->stage[k], bytes));
  }
  if (bytes) VM_CUDA(h, cudaMemcpyAsync(h->stage[k].p, xyz, bytes, cudaMemcpyHostToDevice, h->copy_stream));
  VM_CUDA(h, cudaEventRecord(h->ev_stage[k], h->copy_stream));
  h->stage_n[k] = n;
  h->stage_stride[k] = stride_bytes;
  h->staged[k] = true;
  h->stage_next = k ^ 1;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_insert_staged(vmap* h, int is_dense, const double T_rowmajor[16]) try {
  if (!h || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  // oldest staged slot: the one the next stage call would NOT fill, if it is staged; else the other
  int k = h->stage_next;
  if (!h->staged[k]) k ^= 1;
  if (!h->staged[k]) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "no staged scan");
  // (the slot is released on every exit path: a failed call must not leave it staged for good)
  struct Release { vmap* h; int k; ~Release() { h->staged[k] = false; } } release{h, k};
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_CUDA(h, cudaEventSynchronize(h->ev_stage[k]));   // (the insert picks its own stream: the staged copy must be complete)
  return insert_cloud_common(h, h->stage[k].p, false, h->stage_n[k], h->stage_stride[k], is_dense ? 1 : 0, T_rowmajor, nullptr, nullptr);
} catch (...) { return vmap_on_exception(h); }

int vmap_insert_projected(vmap* h, const float* xyz_hw3, int rows, int cols, size_t row_pitch_bytes,
                          const double q_wxyz[4], const double t[3]) try {
  VM_TRY(check_image_args(h, xyz_hw3, rows, cols, row_pitch_bytes, 12, q_wxyz, t));
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  const size_t bytes = (size_t)rows * row_pitch_bytes;
  if (bytes) {
    VM_TRY(ensure(h, &h->in, bytes));
    VM_CUDA(h, cudaMemcpyAsync(h->in.p, xyz_hw3, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  return insert_image_common(h, (const uint8_t*)h->in.p, rows, cols, row_pitch_bytes, nullptr, q_wxyz, t);
} catch (...) { return vmap_on_exception(h); }

// world_base.cc:55-69
static void fixup_q(const double Q_full[16], double full_image_width, int cols, double Q[16]) {
  std::memcpy(Q, Q_full, 16 * sizeof(double));
  const double downsampling_factor = full_image_width / cols;
  if (std::fabs(downsampling_factor - 1.0) > 1e-6) {
    const double downsampling_squared = downsampling_factor * downsampling_factor;
    Q[0] /= downsampling_factor;
    Q[3] /= downsampling_squared;
    Q[5] /= downsampling_factor;
    Q[7] /= downsampling_squared;
    Q[11] /= downsampling_squared;
    Q[14] /= downsampling_factor;
    Q[15] /= downsampling_factor;
  }
}

int vmap_insert_disparity_dev(vmap* h, const float* d_disparity, int rows, int cols,
                              size_t row_pitch_bytes, const double Q_full_rowmajor[16],
                              double full_image_width, const double q_wxyz[4], const double t[3]) try {
  VM_TRY(check_image_args(h, d_disparity, rows, cols, row_pitch_bytes, 4, q_wxyz, t));
  if (!Q_full_rowmajor) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "null Q");
  VM_CUDA(h, cudaSetDevice(h->device));
  double Q[16];
  fixup_q(Q_full_rowmajor, full_image_width, cols > 0 ? cols : 1, Q);
  return insert_image_common(h, (const uint8_t*)d_disparity, rows, cols, row_pitch_bytes, Q, q_wxyz, t);
} catch (...) { return vmap_on_exception(h); }

int vmap_insert_disparity(vmap* h, const float* disparity, int rows, int cols, size_t row_pitch_bytes,
                          const double Q_full_rowmajor[16], double full_image_width,
                          const double q_wxyz[4], const double t[3]) try {
  VM_TRY(check_image_args(h, disparity, rows, cols, row_pitch_bytes, 4, q_wxyz, t));
  if (!Q_full_rowmajor) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "null Q");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  const size_t bytes = (size_t)rows * row_pitch_bytes;
  if (bytes) {
    VM_TRY(ensure(h, &h->in, bytes));
    VM_CUDA(h, cudaMemcpyAsync(h->in.p, disparity, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  double Q[16];
  fixup_q(Q_full_rowmajor, full_image_width, cols > 0 ? cols : 1, Q);
  return insert_image_common(h, (const uint8_t*)h->in.p, rows, cols, row_pitch_bytes, Q, q_wxyz, t);
} catch (...) { return vmap_on_exception(h); }

int vmap_reproject(vmap* h, const float* disparity, int rows, int cols, size_t row_pitch_bytes,
                   const double Q_rowmajor[16], float* xyz_hw3_out) try {
  const double qi[4] = {1, 0, 0, 0}, ti[3] = {0, 0, 0};
  VM_TRY(check_image_args(h, disparity, rows, cols, row_pitch_bytes, 4, qi, ti));
  if (!Q_rowmajor || !xyz_hw3_out) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  const size_t n = (size_t)rows * cols;
  if (!n) return VMAP_OK;
  const size_t bytes = (size_t)rows * row_pitch_bytes;
  VM_TRY(ensure(h, &h->in, bytes));
  VM_TRY(ensure(h, &h->out, n * 12));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, disparity, bytes, cudaMemcpyHostToDevice, h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->min_disparity_ord, 0xFF, sizeof(uint32_t), h->stream));
  k_min_disparity<<<g_sm_count(h->device) * 4, 256, 0, h->stream>>>(h->d_cnt, (const uint8_t*)h->in.p, rows, cols, row_pitch_bytes);
  QMat Qm;
  std::memcpy(Qm.q, Q_rowmajor, sizeof(Qm.q));
  Transform T;
  std::memset(&T, 0, sizeof(T));
  k_front_disparity<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, Qm, (const uint8_t*)h->in.p, rows, cols,
                                                            row_pitch_bytes, (float*)h->out.p);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(xyz_hw3_out, h->out.p, n * 12, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_apply_keys(vmap* h, const uint16_t* free_k3, size_t n_free, const uint16_t* occ_k3, size_t n_occ) try {
  if (!h || (!free_k3 && n_free) || (!occ_k3 && n_occ)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n_free > 0x7FFFFFFFull || n_occ > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many keys");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  h->stats = vmap_scan_stats{};
  if (n_free + n_occ == 0) return VMAP_OK;
  VM_TRY(ensure_scan(h, 1));
  VM_TRY(ensure(h, &h->in, n_free * 6));
  VM_TRY(ensure(h, &h->aux, n_occ * 6));
  if (n_free) VM_CUDA(h, cudaMemcpyAsync(h->in.p, free_k3, n_free * 6, cudaMemcpyHostToDevice, h->stream));
  if (n_occ) VM_CUDA(h, cudaMemcpyAsync(h->aux.p, occ_k3, n_occ * 6, cudaMemcpyHostToDevice, h->stream));
  return run_scan(h, 1, nullptr, [&]() -> int {
    if (n_occ) {
      k_mark_keys<<<grid_for(n_occ, 256), 256, 0, h->stream>>>(h->m, h->d_cnt, (const uint16_t*)h->aux.p, (uint32_t)n_occ, 1);
      h->launches++;
    }
    if (n_free) {
      k_mark_keys<<<grid_for(n_free, 256), 256, 0, h->stream>>>(h->m, h->d_cnt, (const uint16_t*)h->in.p, (uint32_t)n_free, 0);
      h->launches++;
    }
    VM_CUDA(h, cudaGetLastError());
    return VMAP_OK;
  });
} catch (...) { return vmap_on_exception(h); }

// ---------------------------------------------------------------------------------------------
// queries
// ---------------------------------------------------------------------------------------------
int vmap_query_points_dev(vmap* h, const double* d_xyz, size_t n, uint8_t* d_status) try {
  if (!h || ((!d_xyz || !d_status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many query points");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  k_query_points<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->m, d_xyz, (uint32_t)n, d_status, 0, nullptr);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_query_points(vmap* h, const double* xyz, size_t n, uint8_t* status) try {
  if (!h || ((!xyz || !status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->in, n * 24));
  VM_TRY(ensure(h, &h->out, n));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, xyz, n * 24, cudaMemcpyHostToDevice, h->stream));
  VM_TRY(vmap_query_points_dev(h, (const double*)h->in.p, n, (uint8_t*)h->out.p));
  VM_CUDA(h, cudaMemcpyAsync(status, h->out.p, n, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_query_lines_dev(vmap* h, const double* d_ab, size_t n, uint8_t* d_status) try {
  if (!h || ((!d_ab || !d_status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many query segments");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  k_query_lines<<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->m, d_ab, (uint32_t)n, d_status);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_query_lines(vmap* h, const double* ab, size_t n, uint8_t* status) try {
  if (!h || ((!ab || !status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->in, n * 48));
  VM_TRY(ensure(h, &h->out, n));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, ab, n * 48, cudaMemcpyHostToDevice, h->stream));
  VM_TRY(vmap_query_lines_dev(h, (const double*)h->in.p, n, (uint8_t*)h->out.p));
  VM_CUDA(h, cudaMemcpyAsync(status, h->out.p, n, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// vmap_query_lines_dev plus the number of voxels the reference's loop would probe (measurement).
int vmap_query_lines_probes_dev(vmap* h, const double* d_ab, size_t n, uint8_t* d_status, uint64_t* probes) try {
  if (!h || ((!d_ab || !d_status) && n) || !probes) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many query segments");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  *probes = 0;
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->aux, 16));
  VM_CUDA(h, cudaMemsetAsync(h->aux.p, 0, 16, h->stream));
  k_query_lines_probes<<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->m, d_ab, (uint32_t)n, d_status, (unsigned long long*)h->aux.p);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  unsigned long long v = 0;
  VM_CUDA(h, cudaMemcpyAsync(&v, h->aux.p, 8, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  *probes = v;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// getCellTrueStatusPoint / getCellProbabilityPoint (octomap_world.cc:363-393), batched.
// probability (may be NULL): OcTreeNode::getOccupancy() = 1 - 1 / (1 + exp(log-odds)) evaluated on
// the HOST in double like the reference does (libm exp), -1.0 where there is no node.
int vmap_query_points_probability(vmap* h, const double* xyz, size_t n, uint8_t* status, double* probability) try {
  if (!h || ((!xyz || !status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many query points");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->in, n * 24));
  VM_TRY(ensure(h, &h->out, n));
  VM_TRY(ensure(h, &h->aux, n * 4));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, xyz, n * 24, cudaMemcpyHostToDevice, h->stream));
  k_query_points<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->m, (const double*)h->in.p, (uint32_t)n,
                                                         (uint8_t*)h->out.p, 1, (float*)h->aux.p);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(status, h->out.p, n, cudaMemcpyDeviceToHost, h->stream));
  std::vector<float> lo;
  if (probability) {
    lo.resize(n);
    VM_CUDA(h, cudaMemcpyAsync(lo.data(), h->aux.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  if (probability)
    for (size_t i = 0; i < n; i++)
      probability[i] = (status[i] == VMAP_CELL_UNKNOWN) ? -1.0 : 1. - (1. / (1. + std::exp((double)lo[i])));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// getVisibility (octomap_world.cc:420-449), batched: n x (view point xyz, voxel to test xyz).
int vmap_query_visibility(vmap* h, const double* view_voxel_xyz6, size_t n, int stop_at_unknown_cell, uint8_t* status) try {
  if (!h || ((!view_voxel_xyz6 || !status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many queries");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->in, n * 48));
  VM_TRY(ensure(h, &h->out, n));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, view_voxel_xyz6, n * 48, cudaMemcpyHostToDevice, h->stream));
  k_query_visibility<<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->m, (const double*)h->in.p, (uint32_t)n,
                                                             stop_at_unknown_cell ? 1 : 0, (uint8_t*)h->out.p);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(status, h->out.p, n, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// getCellStatusBoundingBox (octomap_world.cc:274-344), batched: every point with the same box.
static int query_bbox_device(vmap* h, const double* d_xyz, size_t n, const double bounding_box_size[3], uint8_t* d_status) {
  if (h->m.sharded) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bounding-box status needs the whole map on one device");
  BoxQuery q;
  for (int a = 0; a < 3; a++) {
    if (!std::isfinite(bounding_box_size[a])) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bounding_box_size must be finite");
    q.size[a] = bounding_box_size[a];
  }
  const unsigned grid = (unsigned)std::min<size_t>((n + 7) / 8, (size_t)g_sm_count(h->device) * 8);
  k_query_bbox<<<std::max(grid, 1u), 256, 0, h->stream>>>(h->P, h->m, d_xyz, (uint32_t)n, q, d_status);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  return VMAP_OK;
}

int vmap_query_bbox_status_dev(vmap* h, const double* d_xyz, size_t n, const double bounding_box_size[3], uint8_t* d_status) try {
  if (!h || ((!d_xyz || !d_status) && n) || !bounding_box_size) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many queries");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  VM_TRY(query_bbox_device(h, d_xyz, n, bounding_box_size, d_status));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_query_bbox_status(vmap* h, const double* xyz, size_t n, const double bounding_box_size[3], uint8_t* status) try {
  if (!h || ((!xyz || !status) && n) || !bounding_box_size) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many queries");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->in, n * 24));
  VM_TRY(ensure(h, &h->out, n));
  VM_CUDA(h, cudaMemcpyAsync(h->in.p, xyz, n * 24, cudaMemcpyHostToDevice, h->stream));
  VM_TRY(query_bbox_device(h, (const double*)h->in.p, n, bounding_box_size, (uint8_t*)h->out.p));
  VM_CUDA(h, cudaMemcpyAsync(status, h->out.p, n, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// checkPathForCollisionsWithRobot / checkSinglePoseCollision (octomap_world.cc:1384-1412): every
// pose is tested in one launch; the earliest colliding index is picked on the host.
int vmap_check_path_collisions(vmap* h, const double* xyz, size_t n, const double robot_size[3], int* collides,
                               size_t* collision_index) try {
  if (!h || (!xyz && n) || !robot_size || !collides) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  *collides = 0;
  if (!n) return VMAP_OK;
  std::vector<uint8_t> st(n);
  VM_TRY(vmap_query_bbox_status(h, xyz, n, robot_size, st.data()));
  for (size_t i = 0; i < n; i++) {
    const bool hit = h->params.treat_unknown_as_occupied ? (st[i] != VMAP_CELL_FREE) : (st[i] == VMAP_CELL_OCCUPIED);
    if (hit) {
      *collides = 1;
      if (collision_index) *collision_index = i;
      break;
    }
  }
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// getLineStatusBoundingBox (octomap_world.cc:451-493), batched: every segment is swept with the
// same box.  The offset grid is produced by the reference's own loops (accumulating doubles).
int vmap_query_lines_bbox(vmap* h, const double* start_end_xyz6, size_t n, const double bounding_box_size[3], uint8_t* status) try {
  if (!h || ((!start_end_xyz6 || !status) && n) || !bounding_box_size) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  const double epsilon = 0.001;
  const double resolution = h->params.resolution;
  double disc[3];
  for (int a = 0; a < 3; a++) {
    disc[a] = bounding_box_size[a] / std::ceil((bounding_box_size[a] + epsilon) / resolution);
    if (disc[a] <= 0.0) disc[a] = 1.0;
  }
  const double half[3] = {bounding_box_size[0] * 0.5, bounding_box_size[1] * 0.5, bounding_box_size[2] * 0.5};
  std::vector<double> off;
  for (double x = -half[0]; x <= half[0]; x += disc[0])
    for (double y = -half[1]; y <= half[1]; y += disc[1])
      for (double z = -half[2]; z <= half[2]; z += disc[2]) {
        off.push_back(x); off.push_back(y); off.push_back(z);
        if (off.size() > 3u * (1u << 22)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bounding box too large for the resolution");
      }
  const size_t n_off = off.size() / 3;
  if (n_off == 0) {   // (negative box size: the reference's loops never run -> free)
    std::memset(status, 0, n);
    return VMAP_OK;
  }
  if (n > 0x7FFFFFFFull || n_off >= (1u << 30)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many queries");
  VM_TRY(ensure(h, &h->in, n * 48 + off.size() * 8));
  VM_TRY(ensure(h, &h->aux, n * 4));
  VM_TRY(ensure(h, &h->out, n));
  double* d_ab = (double*)h->in.p;
  double* d_off = d_ab + n * 6;
  VM_CUDA(h, cudaMemcpyAsync(d_ab, start_end_xyz6, n * 48, cudaMemcpyHostToDevice, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, h->stream));
  VM_CUDA(h, cudaMemsetAsync(h->aux.p, 0xFF, n * 4, h->stream));
  const unsigned long long work = (unsigned long long)n * n_off;
  if (work > 0x7FFFFFFFull * 128ull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many (segment, offset) pairs");
  k_query_lines_bbox<<<(unsigned)((work + 127) / 128), 128, 0, h->stream>>>(h->P, h->m, d_ab, (uint32_t)n, d_off, (uint32_t)n_off, (uint32_t*)h->aux.p);
  k_unpack_line_status<<<grid_for(n, 256), 256, 0, h->stream>>>((const uint32_t*)h->aux.p, (uint32_t)n, (uint8_t*)h->out.p);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(status, h->out.p, n, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// ---------------------------------------------------------------------------------------------
// leaves in / out
// ---------------------------------------------------------------------------------------------
static int count_leaves(vmap* h, uint16_t* d_k3, float* d_val, uint32_t cap, size_t* n) {
  VM_TRY(flush_app(h));
  VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->n_leaves, 0, sizeof(uint32_t), h->stream));
  k_export<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->m, h->d_cnt, d_k3, d_val, cap);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  uint32_t v = 0;
  VM_CUDA(h, cudaMemcpyAsync(&v, &h->d_cnt->n_leaves, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  *n = v;
  return VMAP_OK;
}

int vmap_leaf_count(vmap* h, size_t* n) try {
  if (!h || !n) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  return count_leaves(h, nullptr, nullptr, 0, n);
} catch (...) { return vmap_on_exception(h); }

int vmap_export_leaves(vmap* h, uint16_t* keys_k3, float* logodds_out, size_t capacity, size_t* n) try {
  if (!h || !n) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  size_t total = 0;
  VM_TRY(count_leaves(h, nullptr, nullptr, 0, &total));
  *n = total;
  if (total == 0) return VMAP_OK;
  if (capacity < total || !keys_k3 || !logodds_out) return fail(h, VMAP_ERR_BUFFER_TOO_SMALL, "export buffers too small; *n holds the leaf count");
  VM_TRY(ensure(h, &h->out, total * 6));
  VM_TRY(ensure(h, &h->aux, total * 4));
  size_t again = 0;
  VM_TRY(count_leaves(h, (uint16_t*)h->out.p, (float*)h->aux.p, (uint32_t)total, &again));
  VM_CUDA(h, cudaMemcpyAsync(keys_k3, h->out.p, total * 6, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(logodds_out, h->aux.p, total * 4, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// The same leaves straight into the caller's DEVICE buffers (no host round trip): the feed of a
// read-only replica on another GPU (volumetric_mapping_b200.sharded.ShardedOctomapWorld.replicate).
int vmap_export_leaves_dev(vmap* h, uint16_t* d_keys_k3, float* d_logodds, size_t capacity, size_t* n) try {
  if (!h || !n) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  size_t total = 0;
  VM_TRY(count_leaves(h, nullptr, nullptr, 0, &total));
  *n = total;
  if (total == 0) return VMAP_OK;
  if (capacity < total || !d_keys_k3 || !d_logodds) return fail(h, VMAP_ERR_BUFFER_TOO_SMALL, "export buffers too small; *n holds the leaf count");
  if (total > 0xFFFFFFFFull) return fail(h, VMAP_ERR_CAPACITY, "too many leaves for one export");
  size_t again = 0;
  return count_leaves(h, d_keys_k3, d_logodds, (uint32_t)total, &again);
} catch (...) { return vmap_on_exception(h); }

int vmap_leaf_digest(vmap* h, uint64_t* n_leaves, uint64_t* digest) try {
  if (!h || !n_leaves || !digest) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  VM_TRY(ensure(h, &h->aux, 16));
  VM_CUDA(h, cudaMemsetAsync(h->aux.p, 0, 16, h->stream));
  k_digest<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->m, (unsigned long long*)h->aux.p);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  unsigned long long v[2] = {0, 0};
  VM_CUDA(h, cudaMemcpyAsync(v, h->aux.p, 16, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_leaves = v[0];
  *digest = v[1];
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

static int import_leaves_common(vmap* h, const uint16_t* keys_k3, const float* logodds_in, size_t n, bool on_device) {
  if (!h || ((!keys_k3 || !logodds_in) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many leaves");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  const uint16_t* d_k3 = keys_k3;
  const float* d_val = logodds_in;
  if (!on_device) {
    VM_TRY(ensure(h, &h->in, n * 6));
    VM_TRY(ensure(h, &h->aux, n * 4));
    VM_CUDA(h, cudaMemcpyAsync(h->in.p, keys_k3, n * 6, cudaMemcpyHostToDevice, h->stream));
    VM_CUDA(h, cudaMemcpyAsync(h->aux.p, logodds_in, n * 4, cudaMemcpyHostToDevice, h->stream));
    d_k3 = (const uint16_t*)h->in.p;
    d_val = (const float*)h->aux.p;
  }
  for (int attempt = 0; attempt < 40; attempt++) {
    VM_TRY(next_epoch(h));
    VM_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
    k_touch_keys<<<grid_for(n, 256), 256, 0, h->stream>>>(h->m, h->d_cnt, d_k3, (uint32_t)n);
    k_assign_slots<<<grid_for(h->table_cap, 256), 256, 0, h->stream>>>(h->m, h->d_cnt);
    k_write_leaves<<<grid_for(n, 256), 256, 0, h->stream>>>(h->m, h->d_cnt, d_k3, d_val, (uint32_t)n);
    h->launches += 3;
    VM_CUDA(h, cudaGetLastError());
    VM_TRY(read_counters(h));
    if (h->h_cnt->overflow_table) {
      VM_TRY(grow_table(h));
      continue;
    }
    if (h->h_cnt->overflow_pool) {
      VM_TRY(grow_pool(h, (uint64_t)h->h_persist->n_entries));
      continue;
    }
    while ((uint64_t)h->h_persist->n_entries * 2 > h->table_cap) VM_TRY(grow_table(h));
    return VMAP_OK;
  }
  return fail(h, VMAP_ERR_CAPACITY, "import could not be applied after repeated growth");
}

int vmap_import_leaves(vmap* h, const uint16_t* keys_k3, const float* logodds_in, size_t n) try {
  return import_leaves_common(h, keys_k3, logodds_in, n, false);
} catch (...) { return vmap_on_exception(h); }

// Leaves that already sit in device memory (of this GPU): read in place.
int vmap_import_leaves_dev(vmap* h, const uint16_t* d_keys_k3, const float* d_logodds, size_t n) try {
  return import_leaves_common(h, d_keys_k3, d_logodds, n, true);
} catch (...) { return vmap_on_exception(h); }

// OctomapWorld::getChangedPoints (octomap_world.cc:1413-1440): the keys octomap's change detection
// collected since the last call (created leaves and leaves whose occupancy flipped), each with
// the CURRENT occupancy of its node, then resetChangeDetection().  Needs change_detection_enabled.
// When `capacity` is too small nothing is reset and *n holds the count.
int vmap_get_changed(vmap* h, uint16_t* keys_k3, uint8_t* occupied, size_t capacity, size_t* n) try {
  if (!h || !n) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  *n = 0;
  if (!h->m.chg0) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "change detection is not enabled (OctomapParameters::change_detection_enabled)");
  const unsigned grid = g_sm_count(h->device) * 8;
  VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->n_leaves, 0, sizeof(uint32_t), h->stream));
  k_changed<<<grid, 256, 0, h->stream>>>(h->P, h->m, h->d_cnt, nullptr, nullptr, 0, 0);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  uint32_t total = 0;
  VM_CUDA(h, cudaMemcpyAsync(&total, &h->d_cnt->n_leaves, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  *n = total;
  if (total == 0) return VMAP_OK;
  if (!keys_k3 || !occupied || capacity < total) return fail(h, VMAP_ERR_BUFFER_TOO_SMALL, "buffers too small; *n holds the number of changed keys");
  VM_TRY(ensure(h, &h->out, (size_t)total * 6));
  VM_TRY(ensure(h, &h->aux, total));
  VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->n_leaves, 0, sizeof(uint32_t), h->stream));
  k_changed<<<grid, 256, 0, h->stream>>>(h->P, h->m, h->d_cnt, (uint16_t*)h->out.p, (uint8_t*)h->aux.p, total, 1);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(keys_k3, h->out.p, (size_t)total * 6, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(occupied, h->aux.p, total, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
int vmap_last_scan_stats(const vmap* h, vmap_scan_stats* stats) try {
  if (!h || !stats) return VMAP_ERR_INVALID_ARGUMENT;
  // (an overlapped multi-GPU round still in flight: its apply-side counters arrive with
  // vmap_dist_flush; the generation-side counters are already final.)  Pipelined scans are
  // retired first, so these are the counters of the scan inserted last.
  vmap* hm = const_cast<vmap*>(h);
  cudaSetDevice(hm->device);
  const int st = pipe_flush(hm);
  *stats = h->stats;
  return st;
} catch (...) { return vmap_on_exception(const_cast<vmap*>(h)); }

int vmap_set_debug_capture(vmap* h, int enabled) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  h->capture = enabled != 0;
  if (!h->capture) h->captured = false;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_debug_last_scan_keys(vmap* h, uint16_t* free_k3, size_t free_capacity, size_t* n_free,
                              uint16_t* occ_k3, size_t occ_capacity, size_t* n_occ) try {
  if (!h || !n_free || !n_occ) return VMAP_ERR_INVALID_ARGUMENT;
  if (!h->captured) return fail(h, VMAP_ERR_NOT_CAPTURED, "debug capture was not enabled for the last scan");
  VM_CUDA(h, cudaSetDevice(h->device));
  *n_free = h->h_cap_n[0];
  *n_occ = h->h_cap_n[1];
  if (!free_k3 && !occ_k3) return VMAP_OK;  // size query
  if (free_capacity < *n_free || occ_capacity < *n_occ) return fail(h, VMAP_ERR_BUFFER_TOO_SMALL, "key buffers too small");
  if (*n_free) VM_CUDA(h, cudaMemcpyAsync(free_k3, h->cap_free.p, *n_free * 6, cudaMemcpyDeviceToHost, h->stream));
  if (*n_occ) VM_CUDA(h, cudaMemcpyAsync(occ_k3, h->cap_occ.p, *n_occ * 6, cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_logodds_constants(const vmap* h, float out5[5]) try {
  if (!h || !out5) return VMAP_ERR_INVALID_ARGUMENT;
  out5[0] = h->P.hit; out5[1] = h->P.miss; out5[2] = h->P.cmin; out5[3] = h->P.cmax; out5[4] = h->P.occ_thres;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(const_cast<vmap*>(h)); }

void* vmap_stream(vmap* h) { return h ? (void*)h->stream : nullptr; }
uint64_t vmap_kernel_launches(const vmap* h) { return h ? h->launches : 0; }
uint64_t vmap_device_bytes(const vmap* h) { return h ? h->dev_bytes : 0; }
double vmap_last_walk_kernel_ms(const vmap* h) { return h ? (double)h->walk_kernel_ms : 0.0; }

}  // extern "C"

#include "vmap_export_host.inl"
#include "vmap_cloud_io_host.inl"
