// vmap_shard_host.inl -- host side of the multi-GPU sharding (included by vmap.cu).
//
//   vmap_shard_generate*   K1..K2 of one scan on this rank + packing into per-owner records
//   vmap_shard_apply_dev   owner side: OR records of one scan after another into the table, K4
//   vmap_dist_*            the same two steps with the exchange done here over NCCL
//                          (libnccl.so.2 is dlopen'ed on first use; single-GPU users never need it)
#include <dlfcn.h>

#include <chrono>

namespace {

// ---- minimal NCCL surface (nccl.h 2.27: types are ABI-stable since 2.x) ----------------------
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclUint8 = 1, kNcclUint64 = 5, kNcclUint32 = 3, kNcclMin = 3 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    g_nccl.error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return false;
  }
  bool ok = true;
  auto sym = [&](const char* s) {
    void* p = dlsym(g_nccl.lib, s);
    if (!p) { ok = false; g_nccl.error = std::string("missing NCCL symbol ") + s; }
    return p;
  };
  g_nccl.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  if (!ok) {
    dlclose(g_nccl.lib);
    g_nccl.lib = nullptr;
  }
  return ok;
}

#define VM_NCCL(h, expr)                                                               \
  do {                                                                                 \
    int r__ = (expr);                                                                  \
    if (r__ != 0) {                                                                    \
      (h)->err = std::string(#expr) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
      return VMAP_ERR_DIST;                                                            \
    }                                                                                  \
  } while (0)

PackView pack_view(vmap* h) {
  PackView pv;
  pv.records = (BlockRecord*)h->pack.p + h->pack_off;   // earlier sub-scans of the batch come first
  pv.counts = h->d_pack_counts;
  pv.cursors = h->d_pack_counts + kMaxRanks;
  pv.capacity = (uint32_t)std::min<size_t>(h->pack.cap / sizeof(BlockRecord) - std::min(h->pack_off, h->pack.cap / sizeof(BlockRecord)), 0xFFFFFFFFu);
  pv.world = (uint32_t)h->shard_world;
  return pv;
}

PackView pack_view(vmap* h);

// Enlarge the record buffer keeping the records of the batch's earlier sub-scans.
int grow_pack(vmap* h, size_t want_records) {
  const size_t have = h->pack.cap / sizeof(BlockRecord);
  if (want_records <= have) return VMAP_OK;
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  size_t bytes = (want_records + (want_records >> 2) + 1024) * sizeof(BlockRecord);
  void* np_ = nullptr;
  VM_TRY(dev_alloc(h, &np_, bytes));
  if (h->pack_off) VM_CUDA(h, cudaMemcpy(np_, h->pack.p, h->pack_off * sizeof(BlockRecord), cudaMemcpyDeviceToDevice));
  dev_free(h, h->pack.p, h->pack.cap);
  h->pack.p = np_;
  h->pack.cap = bytes;
  return VMAP_OK;
}

// Enqueue the two pack kernels and the copy of the per-destination counts (no synchronisation).
int launch_pack(vmap* h) {
  if (h->pack.cap / sizeof(BlockRecord) < h->pack_off + 1024) VM_TRY(grow_pack(h, h->pack_off + (1u << 16)));
  VM_CUDA(h, cudaMemsetAsync(h->d_pack_counts, 0, 2 * kMaxRanks * sizeof(uint32_t), h->stream));
  PackView pv = pack_view(h);
  const unsigned grid = g_sm_count(h->device) * 8;
  k_pack_count<<<grid, 256, 0, h->stream>>>(h->m, h->ms, h->d_cnt, pv);
  k_pack_write<<<grid, 256, 0, h->stream>>>(h->m, h->ms, h->d_cnt, pv);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(h->h_pack_counts, h->d_pack_counts, kMaxRanks * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, h->stream));
  return VMAP_OK;
}

// K1..K2 for one scan with the grow-and-retry protocol; the marks are left in h->ms (dense
// window or the table's mask columns) together with the list of touched blocks.
// `between` runs once, after the first attempt's kernels are enqueued and before the host waits
// for them: the multi-GPU path uses it to exchange the PREVIOUS round while this one is generated.
template <class Mark, class Between>
int generate_marks(vmap* h, size_t n_points, const float* origin, Mark&& mark, Between&& between) {
  bool between_done = false;
  VM_TRY(next_epoch(h));
  uint64_t retries = 0;
  float ms_total = 0.f;
  bool dense = origin && setup_dense_marks(h, origin);
  if (!dense) {
    // a table-mode generate touches the hash table an in-flight apply is updating: no overlap
    between_done = true;
    VM_TRY(between());
    VM_TRY(flush_rounds(h));
    use_table_marks(h);
  }
  h->packed = false;
  auto launch_list = [&]() -> int {
    const unsigned grid = (unsigned)std::min<uint64_t>((h->win.n_words + 255) / 256, (uint64_t)g_sm_count(h->device) * 8);
    k_list_touched<<<std::max(grid, 1u), 256, 0, h->stream>>>(h->ms, h->d_cnt, h->win.n_words);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
    return VMAP_OK;
  };
  for (;;) {
    VM_TRY(begin_scan_scratch(h, n_points));
    for (bool& b : h->evk_set) b = false;
    VM_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    VM_TRY(mark_stage(h, 0));
    VM_TRY(mark());
    if (dense) VM_TRY(launch_list());
    VM_TRY(mark_stage(h, 3));
    if (dense) VM_TRY(launch_pack(h));     // same synchronisation reads the scan's and the pack's counters
    VM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    if (!between_done) {
      between_done = true;
      VM_TRY(between());
    }
    VM_TRY(read_counters(h));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    ms_total += ms;
    if (dense) {
      if (h->h_cnt->overflow_ckpt) {
        VM_TRY(clear_window_marks(h));
        VM_TRY(grow_checkpoints(h));
      } else if (h->h_cnt->overflow_window) {
        VM_TRY(clear_window_marks(h));
        h->dense_fallbacks++;
        dense = false;
        VM_TRY(flush_rounds(h));
        use_table_marks(h);
      } else if (h->h_cnt->overflow_keys) {
        VM_TRY(grow_window_list(h, std::max<uint32_t>(h->h_cnt->n_touched + (h->h_cnt->n_touched >> 2), 2 * h->win.list_cap)));
        h->ms.list = h->win.list; h->ms.locate = h->win.locate; h->ms.list_cap = h->win.list_cap;
        VM_TRY(clear_window_marks(h));   // simplest: regenerate the scan with the larger list
      } else {
        while (h->h_cnt->overflow_pack) {
          // marks and counts are intact: enlarge the record buffer and write again
          const size_t need = h->h_cnt->overflow_pack;
          VM_TRY(grow_pack(h, h->pack_off + need));
          VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->overflow_pack, 0, sizeof(uint32_t), h->stream));
          VM_CUDA(h, cudaMemsetAsync(h->d_pack_counts + kMaxRanks, 0, kMaxRanks * sizeof(uint32_t), h->stream));
          k_pack_write<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->m, h->ms, h->d_cnt, pack_view(h));
          h->launches++;
          VM_CUDA(h, cudaGetLastError());
          VM_TRY(read_counters(h));
        }
        VM_CUDA(h, cudaMemsetAsync(h->win.touched_bits, 0, (size_t)h->win.n_words * 4, h->stream));
        h->packed = true;
        break;
      }
    } else {
      if (!h->h_cnt->overflow_table) break;
      VM_TRY(clear_marks(h));
      VM_TRY(grow_table(h));
      use_table_marks(h);
    }
    if (++retries > 40) return fail(h, VMAP_ERR_CAPACITY, "scan could not be generated after repeated growth");
  }
  const Counters& c = *h->h_cnt;
  vmap_scan_stats& s = h->stats;
  s.points_valid = c.points_valid;
  s.points_in_range = c.points_in_range;
  s.rays_cast = c.rays_cast;
  s.traversals = c.traversals;
  s.occupied_emitted = c.occupied_emitted;
  s.longest_ray = c.longest_ray;
  s.blocks_touched = c.n_touched;
  s.retries = retries;
  s.gpu_ms = ms_total;
  auto span = [&](int a, int b) -> double {
    float v = 0.f;
    if (!h->evk_set[a] || !h->evk_set[b]) return 0.0;
    if (cudaEventElapsedTime(&v, h->evk[a], h->evk[b]) != cudaSuccess) { cudaGetLastError(); return 0.0; }
    return (double)v;
  };
  s.front_ms = span(0, 1);
  s.resolve_ms = span(1, 2);
  s.walk_ms = h->evk_set[2] ? span(2, 3) : span(0, 3);
  return VMAP_OK;
}

// Pack every touched block of the generated scan into per-destination records.
int pack_marks(vmap* h) {
  if (h->packed) {
    while ((uint64_t)h->h_persist->n_entries * 2 > h->table_cap) VM_TRY(grow_table(h));
    return VMAP_OK;
  }
  const size_t n_touched = h->h_cnt->n_touched;
  VM_TRY(grow_pack(h, h->pack_off + std::max<size_t>(n_touched, 1)));
  VM_CUDA(h, cudaMemsetAsync(h->d_pack_counts, 0, 2 * kMaxRanks * sizeof(uint32_t), h->stream));
  PackView pv = pack_view(h);
  const unsigned grid = g_sm_count(h->device) * 8;
  k_pack_count<<<grid, 256, 0, h->stream>>>(h->m, h->ms, h->d_cnt, pv);
  k_pack_write<<<grid, 256, 0, h->stream>>>(h->m, h->ms, h->d_cnt, pv);
  if (h->ms.dense) VM_CUDA(h, cudaMemsetAsync(h->win.touched_bits, 0, (size_t)h->win.n_words * 4, h->stream));
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(h->h_pack_counts, h->d_pack_counts, kMaxRanks * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  // the table may now be re-hashed safely (no pending marks)
  while ((uint64_t)h->h_persist->n_entries * 2 > h->table_cap) VM_TRY(grow_table(h));
  return VMAP_OK;
}

template <class Between>
int shard_generate_common(vmap* h, const uint8_t* d_in, size_t n, size_t stride, int is_dense,
                          const double Tm[16], uint64_t* counts_out, Between&& between) {
  VM_TRY(pipe_flush(h));
  Transform T;
  fill_transform_matrix(&T, Tm);
  VM_TRY(ensure_scan(h, std::max<size_t>(n, 1)));
  h->stats = vmap_scan_stats{};
  h->stats.points_in = n;
  std::memset(h->h_pack_counts, 0, kMaxRanks * sizeof(uint32_t));
  h->packed = false;
  if (!n) VM_TRY(between());
  if (n) {
    VM_TRY(generate_marks(h, n, T.origin, [&]() -> int {
      k_front_cloud<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, d_in, (uint32_t)n,
                                                            (uint32_t)stride, is_dense);
      h->launches++;
      VM_CUDA(h, cudaGetLastError());
      return launch_resolve_and_walk(h, n, T);
    }, between));
    VM_TRY(pack_marks(h));
  }
  if (counts_out)
    for (int d = 0; d < h->shard_world; d++) counts_out[d] = h->h_pack_counts[d];
  return VMAP_OK;
}

// Owner side: apply `n_lists` scans' record lists in order (list i = the i-th scan of the round).
int shard_apply_lists(vmap* h, const BlockRecord* const* lists, const size_t* counts, size_t n_lists) {
  VM_TRY(flush_app(h));
  size_t total = 0;
  for (size_t i = 0; i < n_lists; i++) total += counts[i];
  if (total > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many block records");
  // make room up front so that neither the table nor the pool can overflow inside the round:
  // every record adds at most one table entry and one pool slot
  while (((uint64_t)h->h_persist->n_entries + total) * 2 > h->table_cap) VM_TRY(grow_table(h));
  if ((uint64_t)h->h_persist->n_slots + total > h->pool_cap) VM_TRY(grow_pool(h, (uint64_t)h->h_persist->n_slots + total));
  use_table_marks(h);   // received records are OR-ed into the table's mask columns
  VM_CUDA(h, cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  VM_CUDA(h, cudaEventRecord(h->evk[3], h->stream));
  uint64_t touched = 0;
  for (size_t i = 0; i < n_lists; i++) {
    if (!counts[i]) continue;
    const unsigned grid = (unsigned)std::min<size_t>((counts[i] + 7) / 8, (size_t)g_sm_count(h->device) * 8);
    if (!h->capture) {
      // fused: records -> voxels in one launch per scan
      k_apply_update<<<std::max(grid, 1u), 256, 0, h->stream>>>(h->P, h->m, h->d_cnt, lists[i], (uint32_t)counts[i]);
      h->launches++;
      VM_CUDA(h, cudaGetLastError());
    } else {
      // parity hook: go through the table's mask columns so the key sets can be captured
      VM_TRY(next_epoch(h));
      VM_CUDA(h, cudaMemsetAsync(&h->d_cnt->n_touched, 0, sizeof(uint32_t), h->stream));
      k_apply_records<<<std::max(grid, 1u), 256, 0, h->stream>>>(h->m, h->d_cnt, lists[i], (uint32_t)counts[i]);
      h->launches++;
      VM_CUDA(h, cudaGetLastError());
      k_admit_pool<<<1, 32, 0, h->stream>>>(h->m, h->d_cnt);
      h->launches++;
      VM_TRY(launch_capture(h));
      VM_TRY(launch_update(h));
    }
    touched += counts[i];
  }
  VM_CUDA(h, cudaEventRecord(h->evk[4], h->stream));
  VM_TRY(read_counters(h));
  if (h->h_cnt->overflow_table || h->h_cnt->overflow_pool)
    return fail(h, VMAP_ERR_CAPACITY, "internal: map overflow inside a sharded apply round");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->evk[3], h->evk[4]);
  h->stats.update_ms = ms;
  h->stats.gpu_ms += ms;
  h->stats.unique_free = h->h_cnt->unique_free;
  h->stats.unique_occupied = h->h_cnt->unique_occupied;
  h->stats.blocks_total = h->h_persist->n_slots;
  h->stats.blocks_applied = touched;
  h->captured = h->capture;
  return VMAP_OK;
}

}  // namespace

extern "C" {

uint32_t vmap_shard_owner(uint16_t kx, uint16_t ky, uint16_t kz, int world) {
  if (world <= 1) return 0;
  const unsigned long long b = (unsigned long long)(kx >> 3) | ((unsigned long long)(ky >> 3) << 13) |
                               ((unsigned long long)(kz >> 3) << 26);
  return owner_of(b, (uint32_t)world);
}

size_t vmap_shard_record_bytes(void) { return sizeof(BlockRecord); }

int vmap_set_shard(vmap* h, int rank, int world) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
    return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bad shard rank / world");
  h->shard_rank = rank;
  h->shard_world = world;
  h->m.sharded = world > 1 ? 1u : 0u;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_shard_generate_dev(vmap* h, const float* d_xyz, size_t n, size_t stride_bytes,
                            const double T_rowmajor[16], uint64_t* counts_out) try {
  if (!h || (!d_xyz && n) || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  VM_CUDA(h, cudaSetDevice(h->device));
  return shard_generate_common(h, (const uint8_t*)d_xyz, n, stride_bytes, 1, T_rowmajor, counts_out, []() -> int { return VMAP_OK; });
} catch (...) { return vmap_on_exception(h); }

int vmap_shard_generate(vmap* h, const float* xyz, size_t n, size_t stride_bytes, int is_dense,
                        const double T_rowmajor[16], uint64_t* counts_out) try {
  if (!h || (!xyz && n) || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(pipe_flush(h));
  if (n) {
    const size_t bytes = (n - 1) * stride_bytes + 12;
    VM_TRY(ensure(h, &h->in, bytes));
    VM_CUDA(h, cudaMemcpyAsync(h->in.p, xyz, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  return shard_generate_common(h, (const uint8_t*)h->in.p, n, stride_bytes, is_dense ? 1 : 0, T_rowmajor, counts_out, []() -> int { return VMAP_OK; });
} catch (...) { return vmap_on_exception(h); }

int vmap_shard_records(vmap* h, int dest, const void** d_records, size_t* n_records) try {
  if (!h || !d_records || !n_records) return VMAP_ERR_INVALID_ARGUMENT;
  if (dest < 0 || dest >= h->shard_world) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "bad destination rank");
  size_t off = 0;
  for (int d = 0; d < dest; d++) off += h->h_pack_counts[d];
  *d_records = (const BlockRecord*)h->pack.p + off;
  *n_records = h->h_pack_counts[dest];
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_shard_apply_dev(vmap* h, const void* const* d_record_lists, const size_t* n_records, size_t n_lists) try {
  if (!h || ((!d_record_lists || !n_records) && n_lists)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  return shard_apply_lists(h, (const BlockRecord* const*)d_record_lists, n_records, n_lists);
} catch (...) { return vmap_on_exception(h); }

}  // extern "C"
namespace {
int peer_alloc_mailbox(vmap* h, uint64_t records_per_round);
void peer_set_table(vmap* h, int s, void* base);
uint64_t peer_default_records();
void peer_release(vmap* h);
int peer_insert(vmap* h, const void* src, bool src_on_host, size_t n, size_t stride, int is_dense, const double Tm[16]);
int peer_flush_begin(vmap* h);
int peer_flush_end(vmap* h);
int init_app_stream(vmap* h);
}  // namespace
extern "C" {

// ---- NCCL-backed collective insertion --------------------------------------------------------
int vmap_dist_unique_id(uint8_t out128[128]) try {
  if (!out128) return VMAP_ERR_INVALID_ARGUMENT;
  if (!nccl_load()) {
    g_create_error = g_nccl.error;
    return VMAP_ERR_DIST;
  }
  NcclUniqueId id;
  const int r = g_nccl.GetUniqueId(&id);
  if (r != 0) {
    g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
    return VMAP_ERR_DIST;
  }
  std::memcpy(out128, id.internal, 128);
  return VMAP_OK;
} catch (...) { return vmap_on_exception(nullptr); }

int vmap_dist_init(vmap* h, const uint8_t id128[128], int rank, int world) try {
  if (!h || !id128) return VMAP_ERR_INVALID_ARGUMENT;
  VM_TRY(vmap_set_shard(h, rank, world));
  if (!nccl_load()) return fail(h, VMAP_ERR_DIST, g_nccl.error.c_str());
  VM_CUDA(h, cudaSetDevice(h->device));
  NcclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  NcclComm comm = nullptr;
  VM_NCCL(h, g_nccl.CommInitRank(&comm, world, id, rank));
  h->nccl_comm = comm;
  // Peer-memory transport (default): every rank allocates its mailbox, the CUDA IPC handles travel
  // in one all-gather, every rank maps every mailbox; a second all-gather makes the choice
  // unanimous (one rank that cannot map a peer sends everybody back to the NCCL transport).
  const char* tr = getenv("VMAP_DIST_TRANSPORT");
  if (world > 1 && !(tr && std::strcmp(tr, "nccl") == 0)) {
    int ok = 1;
    if (init_app_stream(h) != VMAP_OK || peer_alloc_mailbox(h, peer_default_records()) != VMAP_OK) ok = 0;
    constexpr size_t kSlot = 128;   // sizeof(cudaIpcMemHandle_t) == 64, plus a flag byte
    static_assert(sizeof(cudaIpcMemHandle_t) <= kSlot - 8, "IPC handle size");
    uint8_t mine[kSlot] = {0};
    if (ok) {
      cudaIpcMemHandle_t hd;
      if (cudaIpcGetMemHandle(&hd, h->mailbox) != cudaSuccess) { cudaGetLastError(); ok = 0; }
      else std::memcpy(mine, &hd, sizeof(hd));
    }
    mine[kSlot - 1] = (uint8_t)ok;
    uint8_t* d_x = nullptr;
    VM_CUDA(h, cudaMalloc((void**)&d_x, (size_t)(world + 1) * kSlot));
    std::vector<uint8_t> all((size_t)world * kSlot);
    auto gather = [&]() -> int {
      VM_CUDA(h, cudaMemcpyAsync(d_x, mine, kSlot, cudaMemcpyHostToDevice, h->stream));
      VM_NCCL(h, g_nccl.AllGather(d_x, d_x + kSlot, kSlot, kNcclUint8, comm, h->stream));
      VM_CUDA(h, cudaMemcpyAsync(all.data(), d_x + kSlot, (size_t)world * kSlot, cudaMemcpyDeviceToHost, h->stream));
      VM_CUDA(h, cudaStreamSynchronize(h->stream));
      return VMAP_OK;
    };
    int st = gather();
    if (st == VMAP_OK) {
      for (int s = 0; s < world; s++) ok = ok && all[(size_t)s * kSlot + kSlot - 1];
      if (ok) {
        for (int s = 0; s < world && ok; s++) {
          if (s == rank) { peer_set_table(h, s, h->mailbox); continue; }
          cudaIpcMemHandle_t hd;
          std::memcpy(&hd, &all[(size_t)s * kSlot], sizeof(hd));
          void* base = nullptr;
          if (cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
          h->peer_ipc_base[s] = base;
          peer_set_table(h, s, base);
        }
      }
      mine[kSlot - 1] = (uint8_t)ok;
      st = gather();                       // unanimous?
      if (st == VMAP_OK)
        for (int s = 0; s < world; s++) ok = ok && all[(size_t)s * kSlot + kSlot - 1];
    }
    cudaFree(d_x);
    VM_TRY(st);
    if (ok) {
      h->pt.capacity = h->peer_capacity;
      h->pt.world = (uint32_t)world;
      h->pt.me = (uint32_t)rank;
      h->peer_on = true;
    } else {
      peer_release(h);
    }
    if (getenv("VMAP_TRACE")) fprintf(stderr, "[vmap rank %d] sharded transport: %s\n", rank, h->peer_on ? "peer memory (CUDA IPC)" : "NCCL send/recv");
  }
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// One process, several handles on one device standing for `world` ranks (tests and single-GPU
// development): the mailboxes are mapped by plain pointers instead of CUDA IPC, everything else is
// the code the multi-process build runs.  Call once, with every handle, before the first insert;
// then drive every handle from its OWN host thread (a rank's host blocks on its peers' progress, as
// separate processes do; one thread driving all handles would wait for work it has not enqueued yet).
int vmap_dist_attach_local(vmap** handles, int world, uint64_t records_per_round) try {
  if (!handles || world < 1 || world > kMaxRanks) return VMAP_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < world; r++) {
    vmap* h = handles[r];
    if (!h) return VMAP_ERR_INVALID_ARGUMENT;
    VM_CUDA(h, cudaSetDevice(h->device));
    VM_TRY(vmap_set_shard(h, r, world));
    VM_TRY(init_app_stream(h));
    VM_TRY(peer_alloc_mailbox(h, records_per_round ? records_per_round : peer_default_records()));
  }
  for (int r = 0; r < world; r++) {
    vmap* h = handles[r];
    for (int s = 0; s < world; s++) peer_set_table(h, s, handles[s]->mailbox);
    h->pt.capacity = h->peer_capacity;
    h->pt.world = (uint32_t)world;
    h->pt.me = (uint32_t)r;
    h->peer_on = true;
    h->peer_local = true;
  }
  return VMAP_OK;
} catch (...) { return vmap_on_exception(nullptr); }

int vmap_dist_flush_begin(vmap* h) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  if (!h->peer_on) return VMAP_OK;
  h->in_collective = true;
  const int st = peer_flush_begin(h);
  h->in_collective = false;
  return st;
} catch (...) { return vmap_on_exception(h); }

int vmap_dist_shutdown(vmap* h) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (h->nccl_comm && !h->peer_on && getenv("VMAP_TIMING") && h->dbg_n)
    fprintf(stderr, "[vmap rank %d] per round (host wall ms): generate+pack %.3f | count allgather + wait for previous apply %.3f | exchange/apply enqueue %.3f | (sync path) %.3f  (%llu rounds)\n",
            h->shard_rank, h->dbg_ms[0] / h->dbg_n, h->dbg_ms[1] / h->dbg_n, h->dbg_ms[2] / h->dbg_n, h->dbg_ms[3] / h->dbg_n,
            (unsigned long long)h->dbg_n);
  if (h->peer_on && getenv("VMAP_TIMING") && h->total_scans)
    fprintf(stderr, "[vmap rank %d] peer transport, device ms per scan (events, overlapped with the other scan context): front %.3f resolve %.3f walk+list %.3f pack %.3f | apply per round %.3f (locate %.3f bucket %.3f apply %.3f) (%llu rounds, %llu records) | scans %llu\n",
            h->shard_rank, h->totals.front_ms / h->total_scans, h->totals.resolve_ms / h->total_scans,
            h->totals.walk_ms / h->total_scans, h->totals.update_ms / h->total_scans,
            h->dbg_n ? h->dbg_ms[0] / h->dbg_n : 0.0, h->dbg_n ? h->dbg_ms[1] / h->dbg_n : 0.0, h->dbg_n ? h->dbg_ms[2] / h->dbg_n : 0.0,
            h->dbg_n ? h->dbg_ms[3] / h->dbg_n : 0.0, (unsigned long long)h->dbg_n, (unsigned long long)h->app_records,
            (unsigned long long)h->total_scans);
  if (h->peer_on || h->mailbox) {
    cudaSetDevice(h->device);
    if (h->peer_on && h->app_stream && h->rounds_published) {
      // nobody may still be reading my round buffers when they are freed: wait (bounded) until every
      // rank has applied my last round
      k_wait_consumed<<<1, 64, 0, h->app_stream>>>(h->pt, (uint32_t)h->rounds_published);
    }
    cudaDeviceSynchronize();
    cudaGetLastError();
    peer_release(h);
  }
  if (h->nccl_comm) {
    cudaSetDevice(h->device);
    h->round_pending = false;   // (a round never exchanged is dropped: vmap_dist_flush is the collective way out)
    flush_rounds(h);
    cudaStreamSynchronize(h->stream);
    g_nccl.CommDestroy((NcclComm)h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// Exchange + apply after a local generate: rank r's scan is the r-th scan of this round.
static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int dist_exchange_and_apply(vmap* h) {
  if (!h->nccl_comm) return fail(h, VMAP_ERR_DIST, "vmap_dist_init has not been called");
  const double t_a = now_ms();
  NcclComm comm = (NcclComm)h->nccl_comm;
  const int G = h->shard_world, me = h->shard_rank;
  // counts matrix: row s = records rank s sends to each destination
  // the per-destination counts are still on the device (k_pack_count): gather them from there
  if (!h->d_count_matrix) VM_TRY(dev_alloc(h, (void**)&h->d_count_matrix, (size_t)kMaxRanks * kMaxBatch * 2 * kMaxRanks * sizeof(uint32_t)));
  h->h_count_matrix.resize((size_t)G * kMaxRanks);
  uint32_t* d_matrix = h->d_count_matrix;
  uint32_t* h_matrix = h->h_count_matrix.data();
  if (!h->packed)   // table-mode generate or an empty scan: counts live on the host only
    VM_CUDA(h, cudaMemcpyAsync(h->d_pack_counts, h->h_pack_counts, kMaxRanks * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  VM_NCCL(h, g_nccl.AllGather(h->d_pack_counts, d_matrix, (size_t)kMaxRanks, kNcclUint32, comm, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(h_matrix, d_matrix, (size_t)G * kMaxRanks * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  const double t_b = now_ms();
  size_t recv_counts[kMaxRanks], recv_off[kMaxRanks], total = 0;
  for (int s = 0; s < G; s++) {
    recv_counts[s] = (size_t)h_matrix[(size_t)s * kMaxRanks + me];
    recv_off[s] = total;
    total += recv_counts[s];
  }
  VM_TRY(ensure(h, &h->recv, std::max<size_t>(total, 1) * sizeof(BlockRecord)));
  BlockRecord* recv = (BlockRecord*)h->recv.p;
  const BlockRecord* send = (const BlockRecord*)h->pack.p;
  size_t send_off[kMaxRanks], so = 0;
  for (int d = 0; d < G; d++) { send_off[d] = so; so += h->h_pack_counts[d]; }
  VM_CUDA(h, cudaEventRecord(h->evk[2], h->stream));
  VM_NCCL(h, g_nccl.GroupStart());
  for (int p = 0; p < G; p++) {
    if (p == me) continue;
    if (h->h_pack_counts[p])
      VM_NCCL(h, g_nccl.Send(send + send_off[p], (size_t)h->h_pack_counts[p] * sizeof(BlockRecord), kNcclUint8, p, comm, h->stream));
    if (recv_counts[p])
      VM_NCCL(h, g_nccl.Recv(recv + recv_off[p], recv_counts[p] * sizeof(BlockRecord), kNcclUint8, p, comm, h->stream));
  }
  VM_NCCL(h, g_nccl.GroupEnd());
  if (recv_counts[me])
    VM_CUDA(h, cudaMemcpyAsync(recv + recv_off[me], send + send_off[me], recv_counts[me] * sizeof(BlockRecord),
                               cudaMemcpyDeviceToDevice, h->stream));
  const BlockRecord* lists[kMaxRanks];
  for (int s = 0; s < G; s++) lists[s] = recv + recv_off[s];
  h->stats.records_sent = so - h->h_pack_counts[me];
  h->stats.records_received = total - recv_counts[me];
  const double t_c = now_ms();
  const int st = shard_apply_lists(h, lists, recv_counts, (size_t)G);
  const double t_d = now_ms();
  h->dbg_ms[1] += t_b - t_a; h->dbg_ms[2] += t_c - t_b; h->dbg_ms[3] += t_d - t_c; h->dbg_n++;
  if (h->dbg_n == 5 && !h->dbg_reset) { h->dbg_reset = true; h->dbg_n = 0; for (double& v : h->dbg_ms) v = 0.0; }   // skip warm-up rounds
  return st;
}

// Wait for the exchange + apply of the previous round (app_stream) and collect its counters.
}  // extern "C"
namespace {
int flush_rounds(vmap* h) {
  if (h->peer_on) {
    if ((h->sub_open || h->rounds_applied < h->rounds_published || h->pipe_n) && !h->in_collective)
      return fail(h, VMAP_ERR_DIST, "sharded scans are still in flight: call vmap_dist_flush on every rank first");
    if (h->app_pending && !h->in_collective) {   // (left by a flush that failed half-way)
      VM_CUDA(h, cudaStreamSynchronize(h->app_stream));
      h->app_pending = false;
      *h->h_persist = *h->h_persist_app;
    }
    return VMAP_OK;
  }
  if (h->round_pending && !h->in_collective)
    return fail(h, VMAP_ERR_DIST, "a sharded round awaits its exchange: call vmap_dist_flush on every rank first");
  if (!h->app_pending) return VMAP_OK;
  VM_CUDA(h, cudaStreamSynchronize(h->app_stream));
  h->app_pending = false;
  if (h->h_cnt_app->overflow_table || h->h_cnt_app->overflow_pool)
    return fail(h, VMAP_ERR_CAPACITY, "internal: map overflow inside a sharded apply round");
  *h->h_persist = *h->h_persist_app;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev_app0, h->ev_app1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
  h->stats.update_ms = ms;
  h->stats.unique_free = h->h_cnt_app->unique_free;
  h->stats.unique_occupied = h->h_cnt_app->unique_occupied;
  h->stats.blocks_total = h->h_persist->n_slots;
  h->stats.blocks_applied = h->app_records;
  h->stats.records_sent = h->app_sent;
  h->stats.records_received = h->app_received;
  return VMAP_OK;
}

// grow-only buffer used on `st` only (idle when this is called)
int ensure_on(vmap* h, DevBuf* b, size_t bytes, cudaStream_t st) {
  if (bytes <= b->cap) return VMAP_OK;
  VM_CUDA(h, cudaStreamSynchronize(st));
  dev_free(h, b->p, b->cap);
  b->p = nullptr;
  b->cap = 0;
  size_t want = std::max(bytes, (size_t)4096);
  want = (want + (want >> 1) + 255) & ~(size_t)255;
  VM_TRY(dev_alloc(h, &b->p, want));
  b->cap = want;
  return VMAP_OK;
}

int init_app_stream(vmap* h) {
  if (h->app_stream) return VMAP_OK;
  // lower priority than the main stream: the exchange / apply of the previous round fills the
  // gaps of the generation of the current one instead of competing with its walk kernel
  int prio_lo = 0, prio_hi = 0;
  VM_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  VM_CUDA(h, cudaStreamCreateWithPriority(&h->app_stream, cudaStreamNonBlocking, prio_lo));
  VM_CUDA(h, cudaEventCreate(&h->ev_app0));
  VM_CUDA(h, cudaEventCreate(&h->ev_app1));
  VM_CUDA(h, cudaMallocHost((void**)&h->h_cnt_app, sizeof(Counters)));
  VM_CUDA(h, cudaMallocHost((void**)&h->h_persist_app, sizeof(Persist)));
  std::memset(h->h_cnt_app, 0, sizeof(Counters));
  VM_TRY(dev_alloc(h, (void**)&h->d_cnt_app, sizeof(Counters)));
  return VMAP_OK;
}

// One round after the local generate + pack: count exchange, record all-to-all and the apply of
// the round's scans, all on app_stream so that they overlap with the generation of the next
// round on the main stream (a dense-window generate never touches the hash table or the pool).
// Returns without waiting; flush_app() (next round, or any call that reads the map) collects it.
int complete_round(vmap* h) {
  if (!h->round_pending) return VMAP_OK;
  if (!h->nccl_comm) return fail(h, VMAP_ERR_DIST, "vmap_dist_init has not been called");
  VM_TRY(init_app_stream(h));
  h->round_pending = false;
  NcclComm comm = (NcclComm)h->nccl_comm;
  const int G = h->shard_world, me = h->shard_rank, M = h->round_subs;
  const int slot = h->round_slot;
  cudaStream_t as = h->app_stream;
  const double t_a = now_ms();
  // (A) count matrix [rank][sub-scan][destination], queued behind the previous round's apply
  constexpr size_t kRow = (size_t)kMaxBatch * 2 * kMaxRanks;
  uint32_t* const d_counts = h->d_pack_counts_base + (size_t)slot * kRow;
  VM_CUDA(h, cudaMemcpyAsync(d_counts, h->h_pack_counts_slot[slot], kRow * sizeof(uint32_t), cudaMemcpyHostToDevice, as));
  if (!h->d_count_matrix) VM_TRY(dev_alloc(h, (void**)&h->d_count_matrix, (size_t)kMaxRanks * kRow * sizeof(uint32_t)));
  h->h_count_matrix.resize((size_t)G * kRow);
  VM_NCCL(h, g_nccl.AllGather(d_counts, h->d_count_matrix, kRow, kNcclUint32, comm, as));
  VM_CUDA(h, cudaMemcpyAsync(h->h_count_matrix.data(), h->d_count_matrix, (size_t)G * kRow * sizeof(uint32_t), cudaMemcpyDeviceToHost, as));
  VM_CUDA(h, cudaStreamSynchronize(as));
  VM_TRY(flush_rounds(h));                         // the round before this one is complete now
  const double t_b = now_ms();
  auto count = [&](int src, int sub, int dst) -> size_t {
    return h->h_count_matrix[(size_t)src * kRow + (size_t)sub * 2 * kMaxRanks + dst];
  };
  // receive layout: scan order = sub-scan major, source rank minor
  std::vector<size_t> recv_cnt((size_t)M * G), recv_off((size_t)M * G);
  size_t total = 0;
  for (int j = 0; j < M; j++)
    for (int s = 0; s < G; s++) {
      recv_cnt[(size_t)j * G + s] = count(s, j, me);
      recv_off[(size_t)j * G + s] = total;
      total += recv_cnt[(size_t)j * G + s];
    }
  if (total > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many block records");
  // send layout: sub-scan major, destination minor (as packed)
  std::vector<size_t> send_off((size_t)M * G);
  size_t so = 0, sent = 0;
  for (int j = 0; j < M; j++)
    for (int d = 0; d < G; d++) {
      send_off[(size_t)j * G + d] = so;
      so += count(me, j, d);
      if (d != me) sent += count(me, j, d);
    }
  // (B) the apply stream is idle here and a dense-window generate (possibly running on the main
  // stream right now) touches neither the map nor these buffers
  VM_TRY(ensure_on(h, &h->recv, std::max<size_t>(total, 1) * sizeof(BlockRecord), as));
  while (((uint64_t)h->h_persist->n_entries + total) * 2 > h->table_cap) VM_TRY(grow_table(h));
  if ((uint64_t)h->h_persist->n_slots + total > h->pool_cap) VM_TRY(grow_pool(h, (uint64_t)h->h_persist->n_slots + total));
  BlockRecord* recv = (BlockRecord*)h->recv.p;
  const BlockRecord* send = (const BlockRecord*)h->pack_inflight.p;
  // (C) exchange + apply, asynchronous
  if (h->l2_flush_bytes) VM_CUDA(h, cudaMemsetAsync(h->l2_flush.p, (int)(h->dbg_n & 0xFF), h->l2_flush_bytes, as));
  VM_CUDA(h, cudaEventRecord(h->ev_app0, as));
  VM_NCCL(h, g_nccl.GroupStart());
  for (int j = 0; j < M; j++)
    for (int p = 0; p < G; p++) {
      if (p == me) continue;
      const size_t ns = count(me, j, p), nr = recv_cnt[(size_t)j * G + p];
      if (ns) VM_NCCL(h, g_nccl.Send(send + send_off[(size_t)j * G + p], ns * sizeof(BlockRecord), kNcclUint8, p, comm, as));
      if (nr) VM_NCCL(h, g_nccl.Recv(recv + recv_off[(size_t)j * G + p], nr * sizeof(BlockRecord), kNcclUint8, p, comm, as));
    }
  VM_NCCL(h, g_nccl.GroupEnd());
  for (int j = 0; j < M; j++)
    if (recv_cnt[(size_t)j * G + me])
      VM_CUDA(h, cudaMemcpyAsync(recv + recv_off[(size_t)j * G + me], send + send_off[(size_t)j * G + me],
                                 recv_cnt[(size_t)j * G + me] * sizeof(BlockRecord), cudaMemcpyDeviceToDevice, as));
  VM_CUDA(h, cudaMemsetAsync(h->d_cnt_app, 0, sizeof(Counters), as));
  for (size_t i = 0; i < recv_cnt.size(); i++) {
    if (!recv_cnt[i]) continue;
    const unsigned grid = (unsigned)std::min<size_t>((recv_cnt[i] + 7) / 8, (size_t)g_sm_count(h->device) * 8);
    k_apply_update<<<std::max(grid, 1u), 256, 0, as>>>(h->P, h->m, h->d_cnt_app, recv + recv_off[i], (uint32_t)recv_cnt[i]);
    h->launches++;
  }
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaEventRecord(h->ev_app1, as));
  VM_CUDA(h, cudaMemcpyAsync(h->h_cnt_app, h->d_cnt_app, sizeof(Counters), cudaMemcpyDeviceToHost, as));
  VM_CUDA(h, cudaMemcpyAsync(h->h_persist_app, h->m.persist, sizeof(Persist), cudaMemcpyDeviceToHost, as));
  h->app_records = total;
  h->app_sent = sent;
  h->app_received = total;
  for (int j = 0; j < M; j++) h->app_received -= recv_cnt[(size_t)j * G + me];
  h->app_pending = true;
  const double t_c = now_ms();
  h->dbg_ms[1] += t_b - t_a; h->dbg_ms[2] += t_c - t_b; h->dbg_n++;
  if (h->dbg_n == 5 && !h->dbg_reset) { h->dbg_reset = true; h->dbg_n = 0; for (double& v : h->dbg_ms) v = 0.0; }   // skip warm-up rounds
  return VMAP_OK;
}

// Seal the batch under construction: it becomes the pending round.
void seal_batch(vmap* h) {
  if (h->pack_sub == 0) return;
  h->round_pending = true;
  h->round_slot = h->pack_slot;
  h->round_subs = h->pack_sub;
  std::swap(h->pack, h->pack_inflight);   // the sealed round's records are always in pack_inflight
  h->pack_slot ^= 1;
  h->pack_sub = 0;
  h->pack_off = 0;
}
}  // namespace
extern "C" {

// Waits for the round in flight (if any); afterwards vmap_last_scan_stats reports its counters.
int vmap_dist_flush(vmap* h) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  if (h->peer_on) {
    h->in_collective = true;
    int st = peer_flush_begin(h);
    const int st2 = peer_flush_end(h);
    h->in_collective = false;
    return st != VMAP_OK ? st : st2;
  }
  h->in_collective = true;
  int st = complete_round(h);          // a sealed round waiting for its exchange
  if (st == VMAP_OK) {
    seal_batch(h);                     // ... and a partial batch (every rank has the same one)
    st = complete_round(h);
  }
  if (st == VMAP_OK) st = flush_rounds(h);
  h->in_collective = false;
  return st;
} catch (...) { return vmap_on_exception(h); }

// Scans per rank that share one exchange (1..8).  Larger batches amortise the exchange over more
// scans; the map lags up to m collective calls behind (vmap_dist_flush completes it).  Collective:
// every rank must use the same value, set between rounds.
int vmap_dist_set_batch(vmap* h, int scans_per_rank) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (scans_per_rank < 1 || scans_per_rank > kMaxBatch) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "batch must be 1..8");
  if (h->pack_sub || h->sub_open) return fail(h, VMAP_ERR_DIST, "a batch is under construction: call vmap_dist_flush first");
  h->batch_m = scans_per_rank;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// Test / benchmark hook: overwrite `bytes` of scratch (larger than L2) on the apply stream before
// every round's exchange, so no round finds its map blocks in L2.  0 disables.
int vmap_set_debug_l2_flush(vmap* h, size_t bytes) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (bytes) VM_TRY(ensure(h, &h->l2_flush, bytes));
  h->l2_flush_bytes = bytes;
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

}  // extern "C"
namespace {
// One collective insert: the new scan is generated on the main stream; while its kernels run, the
// previous round (generated by the previous call) is exchanged and its apply is queued on the
// apply stream.  The new round stays pending until the next collective call or vmap_dist_flush.
int dist_insert(vmap* h, const uint8_t* d_in, size_t n, size_t stride, int is_dense, const double Tm[16]) {
  if (!h->nccl_comm) return fail(h, VMAP_ERR_DIST, "vmap_dist_init has not been called");
  VM_TRY(init_app_stream(h));
  h->in_collective = true;
  struct Leave { vmap* h; ~Leave() { h->in_collective = false; } } leave{h};
  if (h->capture) {   // parity hook: synchronous round through the table's mask columns
    seal_batch(h);
    VM_TRY(complete_round(h));
    VM_TRY(flush_rounds(h));
    h->pack_off = 0;
    h->d_pack_counts = h->d_pack_counts_base;
    h->h_pack_counts = h->h_pack_counts_slot[0][0];
    VM_TRY(shard_generate_common(h, d_in, n, stride, is_dense, Tm, nullptr, []() -> int { return VMAP_OK; }));
    return dist_exchange_and_apply(h);
  }
  // A batch of `batch_m` scans per rank shares one exchange.  Sealing a batch moves its records
  // to `pack_inflight` and flips the count slot (seal_batch), so this batch packs into the buffer
  // the pending round is not using; the sends that last read it were queued a whole batch ago --
  // the main stream waits for them anyway.
  if (h->pack_sub == 0) {
    h->pack_off = 0;
    std::memset(h->h_pack_counts_slot[h->pack_slot], 0, sizeof(h->h_pack_counts_slot[h->pack_slot]));
  }
  h->d_pack_counts = h->d_pack_counts_base + ((size_t)h->pack_slot * kMaxBatch + h->pack_sub) * 2 * kMaxRanks;
  h->h_pack_counts = h->h_pack_counts_slot[h->pack_slot][h->pack_sub];
  if (h->app_pending) VM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_app1, 0));
  const double t0 = now_ms();
  VM_TRY(shard_generate_common(h, d_in, n, stride, is_dense, Tm, nullptr, [&]() -> int { return complete_round(h); }));
  h->dbg_ms[0] += now_ms() - t0;
  for (int d = 0; d < h->shard_world; d++) h->pack_off += h->h_pack_counts[d];
  h->pack_sub++;
  if (h->pack_sub >= h->batch_m) seal_batch(h);
  return VMAP_OK;
}
}  // namespace
extern "C" {

int vmap_dist_insert_pointcloud_dev(vmap* h, const float* d_xyz, size_t n, size_t stride_bytes,
                                    const double T_rowmajor[16]) try {
  if (!h || (!d_xyz && n) || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  VM_CUDA(h, cudaSetDevice(h->device));
  if (h->peer_on) {
    h->in_collective = true;
    const int st = peer_insert(h, d_xyz, false, n, stride_bytes, 1, T_rowmajor);
    h->in_collective = false;
    return st;
  }
  return dist_insert(h, (const uint8_t*)d_xyz, n, stride_bytes, 1, T_rowmajor);
} catch (...) { return vmap_on_exception(h); }

int vmap_dist_insert_pointcloud(vmap* h, const float* xyz, size_t n, size_t stride_bytes, int is_dense,
                                const double T_rowmajor[16]) try {
  if (!h || (!xyz && n) || !T_rowmajor) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "stride_bytes must be a multiple of 4 and >= 12");
  VM_CUDA(h, cudaSetDevice(h->device));
  if (h->peer_on) {
    h->in_collective = true;
    const int st = peer_insert(h, xyz, true, n, stride_bytes, is_dense ? 1 : 0, T_rowmajor);
    h->in_collective = false;
    return st;
  }
  VM_TRY(pipe_flush(h));
  if (n) {
    const size_t bytes = (n - 1) * stride_bytes + 12;
    VM_TRY(ensure(h, &h->in, bytes));
    VM_CUDA(h, cudaMemcpyAsync(h->in.p, xyz, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  return dist_insert(h, (const uint8_t*)h->in.p, n, stride_bytes, is_dense ? 1 : 0, T_rowmajor);
} catch (...) { return vmap_on_exception(h); }

// Sharded queries: partial answers of this shard (combine = element-wise MIN over ranks).
int vmap_shard_query_points_dev(vmap* h, const double* d_xyz, size_t n, uint8_t* d_status) try {
  if (!h || ((!d_xyz || !d_status) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many query points");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  k_query_points_sharded<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->m, d_xyz, (uint32_t)n, d_status,
                                                                 (uint32_t)h->shard_rank, (uint32_t)h->shard_world);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_shard_query_lines_dev(vmap* h, const double* d_ab, size_t n, uint32_t* d_packed) try {
  if (!h || ((!d_ab || !d_packed) && n)) return h ? fail(h, VMAP_ERR_INVALID_ARGUMENT, "null argument") : VMAP_ERR_INVALID_ARGUMENT;
  if (n > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many query segments");
  VM_CUDA(h, cudaSetDevice(h->device));
  VM_TRY(flush_app(h));
  if (!n) return VMAP_OK;
  k_query_lines_sharded<<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->m, d_ab, (uint32_t)n, d_packed,
                                                                (uint32_t)h->shard_rank, (uint32_t)h->shard_world);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

// Collective versions: every rank passes the same queries and receives the global answer.
int vmap_dist_query_points_dev(vmap* h, const double* d_xyz, size_t n, uint8_t* d_status) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (!h->nccl_comm) return fail(h, VMAP_ERR_DIST, "vmap_dist_init has not been called");
  VM_TRY(vmap_dist_flush(h));
  VM_TRY(vmap_shard_query_points_dev(h, d_xyz, n, d_status));
  if (!n) return VMAP_OK;
  VM_NCCL(h, g_nccl.AllReduce(d_status, d_status, n, kNcclUint8, kNcclMin, (NcclComm)h->nccl_comm, h->stream));
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

int vmap_dist_query_lines_dev(vmap* h, const double* d_ab, size_t n, uint8_t* d_status) try {
  if (!h) return VMAP_ERR_INVALID_ARGUMENT;
  if (!h->nccl_comm) return fail(h, VMAP_ERR_DIST, "vmap_dist_init has not been called");
  VM_TRY(vmap_dist_flush(h));
  if (!n) return VMAP_OK;
  VM_TRY(ensure(h, &h->aux, n * sizeof(uint32_t)));
  VM_TRY(vmap_shard_query_lines_dev(h, d_ab, n, (uint32_t*)h->aux.p));
  VM_NCCL(h, g_nccl.AllReduce(h->aux.p, h->aux.p, n, kNcclUint32, kNcclMin, (NcclComm)h->nccl_comm, h->stream));
  k_unpack_line_status<<<grid_for(n, 256), 256, 0, h->stream>>>((const uint32_t*)h->aux.p, (uint32_t)n, d_status);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaStreamSynchronize(h->stream));
  return VMAP_OK;
} catch (...) { return vmap_on_exception(h); }

}  // extern "C"
