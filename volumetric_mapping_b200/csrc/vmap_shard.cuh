// vmap_shard.cuh -- multi-GPU sharding of the voxel-block map.
//
// Every 8x8x8 voxel block has exactly one owner rank (owner_of); a rank generates the key sets
// of its own scans locally (K1..K2 into the mask columns of its own table), packs every touched
// block into a per-destination record stream, the streams are exchanged (all-to-all over NVLink)
// and each owner ORs the received masks into its table and runs the ordinary K4 update -- one
// scan at a time, in scan order, so the non-commutative clamped updates stay ordered per voxel.
#pragma once

#include "vmap_kernels.cuh"

namespace vmapdev {

constexpr int kMaxRanks = 64;
constexpr int kMaxBatch = 8;     // scans per rank in one exchange round
constexpr uint32_t kSlotClaimed = 0xFFFFFFFEu;   // MapView::slot while k_locate_records reserves the block's pool slot

// One touched block of one scan: 8 + 64 + 64 bytes.
struct __align__(8) BlockRecord {
  unsigned long long block;          // 39-bit block key (vmap_device.cuh block_key)
  uint32_t free_mask[kMaskWords];
  uint32_t occ_mask[kMaskWords];
};
static_assert(sizeof(BlockRecord) == 136, "BlockRecord layout");

// Owner of a block: a mix of the block key's upper hash bits (the table index uses the low
// ones), so ownership is balanced wherever the sensor is and independent of the table size.
__host__ __device__ __forceinline__ uint32_t owner_of(unsigned long long b, uint32_t world) {
  b ^= b >> 30; b *= 0xbf58476d1ce4e5b9ull;
  b ^= b >> 27; b *= 0x94d049bb133111ebull;
  b ^= b >> 31;
  return (uint32_t)((b >> 40) % (unsigned long long)world);
}

struct PackView {
  BlockRecord* records;     // [capacity] bucketed by destination
  uint32_t* counts;         // [kMaxRanks] records per destination (pass 1)
  uint32_t* cursors;        // [kMaxRanks] running write position per destination (pass 2)
  uint32_t capacity;
  uint32_t world;
};

// pass 1: how many touched blocks with a non-empty mask go to each destination.  Counts are
// accumulated per CTA in shared memory: one global atomic per destination per CTA, not per block
// (tens of thousands of atomics on `world` addresses serialise in the L2).
__global__ void __launch_bounds__(256) k_pack_count(MapView m, MarkSet ms, Counters* cnt, PackView pv) {
  if (cnt->overflow_table || cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys) return;
  __shared__ uint32_t s_cnt[kMaxRanks];
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_touched = cnt->n_touched;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_touched; t += warps) {
    const uint32_t e = ms.list[t];
    const uint32_t wv = (lane < kMaskWords) ? ms.free_mask[(size_t)e * kMaskWords + lane]
                                            : ms.occ_mask[(size_t)e * kMaskWords + lane - kMaskWords];
    if (!__any_sync(0xFFFFFFFFu, wv != 0u)) continue;
    const unsigned long long b = ms.dense ? window_block_key(ms, e) : (m.keys[e] >> kEpochBits);
    if (lane == 0) atomicAdd(&s_cnt[owner_of(b, pv.world)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < pv.world && s_cnt[threadIdx.x]) atomicAdd(pv.counts + threadIdx.x, s_cnt[threadIdx.x]);
}

// pass 2: write the records (bucket offsets = exclusive prefix of counts) and clear the masks.
// A CTA works on rounds of 8 blocks (one per warp): the slots of a round are reserved with one
// global atomic per destination, the position inside the reservation comes from shared memory.
__global__ void __launch_bounds__(256) k_pack_write(MapView m, MarkSet ms, Counters* cnt, PackView pv) {
  if (cnt->overflow_table || cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys) return;
  __shared__ uint32_t s_base[kMaxRanks];   // first record of every destination's bucket
  __shared__ uint32_t s_cnt[kMaxRanks];    // blocks of this round per destination
  __shared__ uint32_t s_res[kMaxRanks];    // reservation of this round inside the bucket
  {
    // all-or-nothing: when the record buffer is too small nothing is written and no mask is
    // cleared, so the host can enlarge it and launch this kernel again
    uint32_t total = 0;
    for (uint32_t d = 0; d < pv.world; d++) total += pv.counts[d];
    if (total > pv.capacity) {
      if (blockIdx.x == 0 && threadIdx.x == 0) cnt->overflow_pack = total;
      return;
    }
  }
  if (threadIdx.x < pv.world) {
    uint32_t base = 0;
    for (uint32_t d = 0; d < threadIdx.x; d++) base += pv.counts[d];
    s_base[threadIdx.x] = base;
  }
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t n_touched = cnt->n_touched;
  const uint32_t stride = gridDim.x * warps_per_cta;
  for (uint32_t t0 = blockIdx.x * warps_per_cta; t0 < n_touched; t0 += stride) {   // uniform per CTA
    if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t t = t0 + warp;
    uint32_t wv = 0, dest = 0, local = 0;
    unsigned long long b = 0;
    uint32_t* src = nullptr;
    bool have = false;
    if (t < n_touched) {
      const uint32_t e = ms.list[t];
      src = (lane < kMaskWords) ? ms.free_mask + (size_t)e * kMaskWords + lane
                                : ms.occ_mask + (size_t)e * kMaskWords + lane - kMaskWords;
      wv = *src;
      have = __any_sync(0xFFFFFFFFu, wv != 0u);
      if (have) {
        b = ms.dense ? window_block_key(ms, e) : (m.keys[e] >> kEpochBits);
        dest = owner_of(b, pv.world);
        if (lane == 0) local = atomicAdd(&s_cnt[dest], 1u);
        local = __shfl_sync(0xFFFFFFFFu, local, 0);
      }
    }
    __syncthreads();
    if (threadIdx.x < pv.world && s_cnt[threadIdx.x]) s_res[threadIdx.x] = atomicAdd(pv.cursors + threadIdx.x, s_cnt[threadIdx.x]);
    __syncthreads();
    if (have) {
      if (wv) *src = 0u;
      const uint32_t pos = s_base[dest] + s_res[dest] + local;
      if (pos < pv.capacity) {   // always true: capacity >= sum(counts)
        BlockRecord* r = pv.records + pos;
        if (lane == 0) r->block = b;
        if (lane < kMaskWords) r->free_mask[lane] = wv;
        else r->occ_mask[lane - kMaskWords] = wv;
      }
    }
    __syncthreads();
  }
}

// Owner side: OR one scan's received records into the table's mask columns (one warp per
// record); K4 (k_update) then applies them.
__global__ void __launch_bounds__(256)
k_apply_records(MapView m, Counters* cnt, const BlockRecord* __restrict__ rec, uint32_t n) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const BlockRecord* r = rec + i;
    uint32_t h = kNotFound;
    if (lane == 0) h = find_or_insert_touch(m, cnt, r->block);
    h = __shfl_sync(0xFFFFFFFFu, h, 0);
    if (h == kNotFound) continue;
    const uint32_t wv = (lane < kMaskWords) ? r->free_mask[lane] : r->occ_mask[lane - kMaskWords];
    if (wv) {
      uint32_t* dst = (lane < kMaskWords) ? m.free_mask + (size_t)h * kMaskWords + lane
                                          : m.occ_mask + (size_t)h * kMaskWords + lane - kMaskWords;
      atomicOr(dst, wv);
    }
  }
}

// Owner side, fused: one warp per record looks the block up (inserting it and giving it a pool
// slot when new) and applies the record's masks straight to the voxels -- occupied wins, one
// clamped update per marked voxel, exactly K4's arithmetic.  A block appears at most once per
// list, so every voxel still has a single writer per scan.
__global__ void __launch_bounds__(256)
k_apply_update(DevParams P, MapView m, Counters* cnt, const BlockRecord* __restrict__ rec, uint32_t n) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t n_free = 0, n_occ = 0;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    // a record is 34 words: key (2), free mask (16), occupied mask (16); lane l holds mask word l
    const uint32_t mask = reinterpret_cast<const uint32_t*>(rec + i)[2 + lane];
    uint32_t slot = kUnassigned;
    if (lane == 0) {
      const uint32_t h = find_or_insert(m, cnt, rec[i].block);
      if (h != kNotFound) {
        slot = m.slot[h];
        if (slot == kUnassigned) {
          slot = atomicAdd(&m.persist->n_slots, 1u);
          if (slot < m.pool_cap) m.slot[h] = slot;
          else { cnt->overflow_pool = 1u; slot = kUnassigned; }
        }
      }
    }
    slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
    if (slot == kUnassigned) continue;
    // lanes 0..15: free words with the occupied bits removed (occupied wins); lanes 16..31: occupied words
    const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, mask, 16);
    const uint32_t mine = lane < kMaskWords ? (mask & ~other) : mask;
    uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    uint32_t val[kMaskWords];       // all rows are loaded before the first one is updated: one round trip per block
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t any = __shfl_sync(0xFFFFFFFFu, mine, j) | __shfl_sync(0xFFFFFFFFu, mine, j + kMaskWords);
      val[j] = ((any >> lane) & 1u) ? __ldcg(vox + j * 32 + lane) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t fw = __shfl_sync(0xFFFFFFFFu, mine, j);
      const uint32_t ow = __shfl_sync(0xFFFFFFFFu, mine, j + kMaskWords);
      const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
      if (is_occ || is_free) vox[j * 32 + lane] = update_leaf(val[j], is_occ ? P.hit : P.miss, P.cmin, P.cmax);
      if (lane == 0) { n_free += __popc(fw); n_occ += __popc(ow); }
    }
  }
  if (lane == 0) {
    if (n_free) atomicAdd(&cnt->unique_free, n_free);
    if (n_occ) atomicAdd(&cnt->unique_occupied, n_occ);
  }
}

// Owner side of the peer-memory transport, in two steps per round:
//   k_locate_records  ONE launch over every record of the round, one thread per record: looks the block
//                     up (inserting it when new) and gives it a pool slot; located[] keeps the table
//                     index.  Order-free: inserting blocks and reserving slots commute.
//   k_apply_located   one launch per (sub-scan, source rank) list, in scan order, one warp per record:
//                     all of the block's voxel rows are loaded before the first one is updated.
struct RecordLists {                               // in device memory, rewritten for every round
  const BlockRecord* ptr[kMaxBatch * kMaxRanks];
  uint32_t first[kMaxBatch * kMaxRanks + 1];       // index of every list's first record in located[]
  uint32_t n_lists;
};

// round_blocks != nullptr (round-fused apply, below): the first record of the round that names a block
// also registers it -- m.touched[h] (all ones before the launch) becomes the block's index in
// round_blocks[], cnt->round_blocks counts them.
__global__ void __launch_bounds__(256)
k_locate_records(MapView m, Counters* cnt, const RecordLists* __restrict__ lists, uint32_t* __restrict__ located,
                 uint32_t* __restrict__ round_blocks) {
  __shared__ uint32_t s_first[kMaxBatch * kMaxRanks + 1];
  const uint32_t n_lists = lists->n_lists;
  for (uint32_t i = threadIdx.x; i <= n_lists; i += blockDim.x) s_first[i] = lists->first[i];
  __syncthreads();
  const uint32_t total = s_first[n_lists];
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t g0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; g0 < total; g0 += gridDim.x * blockDim.x) {   // uniform per warp
    const uint32_t g = g0 + lane;
    uint32_t h = kNotFound;
    if (g < total) {
      uint32_t lo = 0, hi = n_lists;               // list of record g: largest l with first[l] <= g
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_first[mid] <= g) lo = mid; else hi = mid;
      }
      h = find_or_insert(m, cnt, lists->ptr[lo][g - s_first[lo]].block);
      // A new block gets its pool slot here.  Several records of the round (different scans) may name
      // the same new block: the first to swap the claim mark in reserves the slot, the others move on --
      // k_apply_located reads the slot from the table, after this kernel.
      if (h != kNotFound && m.slot[h] == kUnassigned && atomicCAS(m.slot + h, kUnassigned, kSlotClaimed) == kUnassigned) {
        const uint32_t mine = atomicAdd(&m.persist->n_slots, 1u);
        if (mine < m.pool_cap) {
          atomicExch(m.slot + h, mine);
        } else {
          cnt->overflow_pool = 1u;
          atomicExch(m.slot + h, kUnassigned);
        }
      }
      located[g] = h;
      if (round_blocks && h != kNotFound && atomicCAS(m.touched + h, kUnassigned, kSlotClaimed) == kUnassigned) {
        const uint32_t b = atomicAdd(&cnt->round_blocks, 1u);
        round_blocks[b] = h;
        atomicExch(m.touched + h, b);
      }
    }
  }
}

// Round-fused apply.  Consecutive scans touch nearly the same blocks, and a round brings a block up to
// n_lists records (one per scan): applying the lists one launch after the other reads and writes every
// such block n_lists times -- that traffic is what bounds the apply.  Instead the round's records are
// grouped by block (k_bucket_records: bucket[b * n_lists + k] = record index, any order) and ONE warp per
// block (k_apply_round) loads the block's 512 voxels into registers once, applies the block's records in
// scan order (record indices are laid out list by list, so ascending index = scan order:
// octomap_world.cc:244-262 per voxel), and writes the rows that changed back once.
__global__ void __launch_bounds__(256)
k_bucket_records(MapView m, const uint32_t* __restrict__ located, uint32_t total, uint32_t n_lists,
                 uint32_t* __restrict__ bcount, uint32_t* __restrict__ bucket) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const uint32_t h = located[g];
  if (h == kNotFound) return;
  const uint32_t b = __ldcg(m.touched + h);
  const uint32_t pos = atomicAdd(bcount + b, 1u);
  if (pos < n_lists) bucket[(size_t)b * n_lists + pos] = g;     // (always: a block appears at most once per list)
}

__global__ void __launch_bounds__(256)
k_apply_round(DevParams P, MapView m, Counters* cnt, const RecordLists* __restrict__ lists,
              const uint32_t* __restrict__ round_blocks, const uint32_t* __restrict__ bcount,
              const uint32_t* __restrict__ bucket) {
  __shared__ uint32_t s_first[kMaxBatch * kMaxRanks + 1];
  const uint32_t n_lists = lists->n_lists;            // <= 32 (the host takes the list-by-list path otherwise)
  for (uint32_t i = threadIdx.x; i <= n_lists; i += blockDim.x) s_first[i] = lists->first[i];
  __syncthreads();
  const uint32_t n_blocks = cnt->round_blocks;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t n_marks = 0;
  auto record_words = [&](uint32_t g) -> const uint32_t* {
    uint32_t lo = 0, hi = n_lists;                     // list of record g: largest l with first[l] <= g
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_first[mid] <= g) lo = mid; else hi = mid;
    }
    return reinterpret_cast<const uint32_t*>(lists->ptr[lo] + (g - s_first[lo]));
  };
  for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += warps) {
    const uint32_t h = round_blocks[b];
    const uint32_t n = min(bcount[b], n_lists);
    const uint32_t slot = __ldcg(m.slot + h);
    if (slot >= m.pool_cap) continue;                  // (no slot: the pool overflowed, the host reports it)
    // lane k holds record k of the bucket; its rank among them = its place in scan order
    const uint32_t g_mine = lane < n ? bucket[(size_t)b * n_lists + lane] : 0xFFFFFFFFu;
    uint32_t rank = 0;
    for (uint32_t k = 0; k < n; k++) rank += __shfl_sync(0xFFFFFFFFu, g_mine, k) < g_mine ? 1u : 0u;
    uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    uint32_t val[kMaskWords];
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) val[j] = __ldcg(vox + j * 32 + lane);
    uint32_t dirty = 0;                                // rows with at least one marked voxel (warp-uniform)
    // the first record's masks are requested now, every later one while its predecessor is applied
    uint32_t src = (uint32_t)__ffs(__ballot_sync(0xFFFFFFFFu, lane < n && rank == 0u)) - 1u;
    uint32_t nx_mask = record_words(__shfl_sync(0xFFFFFFFFu, g_mine, src))[2 + lane];
    for (uint32_t r = 0; r < n; r++) {
      const uint32_t mask = nx_mask;
      if (r + 1 < n) {
        src = (uint32_t)__ffs(__ballot_sync(0xFFFFFFFFu, lane < n && rank == r + 1u)) - 1u;
        nx_mask = record_words(__shfl_sync(0xFFFFFFFFu, g_mine, src))[2 + lane];
      }
      // lanes 0..15: free words with the occupied bits removed (occupied wins); lanes 16..31: occupied words
      const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, mask, 16);
      const uint32_t mine = lane < kMaskWords ? (mask & ~other) : mask;
      n_marks += __popc(mine);                         // lanes 0..15 count free voxels, lanes 16..31 occupied ones
#pragma unroll
      for (int j = 0; j < kMaskWords; j++) {
        const uint32_t fw = __shfl_sync(0xFFFFFFFFu, mine, j);
        const uint32_t ow = __shfl_sync(0xFFFFFFFFu, mine, j + kMaskWords);
        if ((fw | ow) == 0u) continue;
        dirty |= 1u << j;
        const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
        if (is_occ || is_free) val[j] = update_leaf(val[j], is_occ ? P.hit : P.miss, P.cmin, P.cmax);
      }
    }
#pragma unroll
    for (int j = 0; j < kMaskWords; j++)
      if ((dirty >> j) & 1u) vox[j * 32 + lane] = val[j];
  }
  for (int o = 8; o > 0; o >>= 1) n_marks += __shfl_xor_sync(0xFFFFFFFFu, n_marks, o);   // sums within each half warp
  if (lane == 0 && n_marks) atomicAdd(&cnt->unique_free, n_marks);
  if (lane == kMaskWords && n_marks) atomicAdd(&cnt->unique_occupied, n_marks);
}

__global__ void __launch_bounds__(256)
k_apply_located(DevParams P, MapView m, Counters* cnt, const BlockRecord* __restrict__ rec,
                const uint32_t* __restrict__ located, uint32_t n) {
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t n_free = 0, n_occ = 0;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const uint32_t h = located[i];
    if (h == kNotFound) continue;
    const uint32_t slot = __ldcg(m.slot + h);
    if (slot >= m.pool_cap) continue;              // (no slot: the pool overflowed, the host reports it)
    // a record is 34 words: key (2), free mask (16), occupied mask (16); lane l holds mask word l
    const uint32_t mask = reinterpret_cast<const uint32_t*>(rec + i)[2 + lane];
    // lanes 0..15: free words with the occupied bits removed (occupied wins); lanes 16..31: occupied words
    const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, mask, 16);
    const uint32_t mine = lane < kMaskWords ? (mask & ~other) : mask;
    uint32_t* vox = reinterpret_cast<uint32_t*>(m.pool) + (size_t)slot * kBlockVoxels;
    uint32_t val[kMaskWords];
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t any = __shfl_sync(0xFFFFFFFFu, mine, j) | __shfl_sync(0xFFFFFFFFu, mine, j + kMaskWords);
      val[j] = ((any >> lane) & 1u) ? __ldcg(vox + j * 32 + lane) : 0u;
    }
#pragma unroll
    for (int j = 0; j < kMaskWords; j++) {
      const uint32_t fw = __shfl_sync(0xFFFFFFFFu, mine, j);
      const uint32_t ow = __shfl_sync(0xFFFFFFFFu, mine, j + kMaskWords);
      const bool is_occ = (ow >> lane) & 1u, is_free = (fw >> lane) & 1u;
      if (is_occ || is_free) vox[j * 32 + lane] = update_leaf(val[j], is_occ ? P.hit : P.miss, P.cmin, P.cmax);
      if (lane == 0) { n_free += __popc(fw); n_occ += __popc(ow); }
    }
  }
  if (lane == 0) {
    if (n_free) atomicAdd(&cnt->unique_free, n_free);
    if (n_occ) atomicAdd(&cnt->unique_occupied, n_occ);
  }
}

// ---------------------------------------------------------------------------------------------
// Peer-memory exchange (vmap_peer_host.inl).  Every rank owns one MAILBOX in its device memory that
// all ranks of the box map through CUDA IPC:
//   * two ROUND buffers (round parity) of block records, packed by this rank's scans and bucketed by
//     destination, with a header that says where each (sub-scan, destination) bucket is;
//   * `ready[s]`    -- written by rank s INTO THIS mailbox: the last round rank s has published;
//   * `consumed[d]` -- written by rank d into this mailbox: the last round of MINE rank d has applied.
// A rank PUBLISHES a round by one remote store per peer after its pack kernels are done, PULLS the
// records destined to it straight out of the peers' round buffers while it applies them (the
// all-to-all is the apply kernel's own loads over NVLink), and tells every peer when it is done
// with their buffer.  Flags only ever grow; waiting is a bounded spin in a one-warp kernel.
// ---------------------------------------------------------------------------------------------
struct RoundHeader {
  uint32_t round;                               // round number + 1 once the round is sealed
  uint32_t n_subs;                              // scans of this rank in the round
  uint32_t fill;                                // records written so far
  uint32_t status;                              // 0 = fine, 1 = the round buffer overflowed (fatal)
  uint32_t count[kMaxBatch][kMaxRanks];         // records of sub-scan j for destination d
  uint32_t offset[kMaxBatch][kMaxRanks];        // first record of that bucket
};
struct MailboxHead {
  RoundHeader hdr[2];
  uint32_t ready[kMaxRanks];
  uint32_t consumed[kMaxRanks];
  uint32_t error;                               // a wait timed out / a peer reported an overflow
  uint32_t pad[63];
};
struct PeerTable {                              // by value into the kernels
  MailboxHead* head[kMaxRanks];
  BlockRecord* rec[kMaxRanks];                  // round buffers: rec[s] + parity * capacity
  uint32_t capacity;                            // records per round buffer
  uint32_t world, me;
};

constexpr long long kSpinLimitNs = 20ll * 1000 * 1000 * 1000;   // a peer that is 20 s late is gone
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// Start of a round's first pack: the round buffer of this parity was last used two rounds ago; wait
// until every rank has applied that round (want = its number + 1; 0 = nothing to wait for), then
// reset the header.
__device__ __forceinline__ void wait_consumed(const PeerTable& pt, uint32_t want) {
  MailboxHead* mine = pt.head[pt.me];
  if (threadIdx.x < pt.world && want) {
    const long long t0 = global_ns();
    while (ld_flag(&mine->consumed[threadIdx.x]) < want && !ld_flag(&mine->error)) {
      __nanosleep(256);
      if (global_ns() - t0 > kSpinLimitNs) { mine->error = 2u; break; }
    }
  }
}
// Flush: wait until every rank has applied my rounds up to number want - 1 (nobody reads my round
// buffers any more).
__global__ void k_wait_consumed(PeerTable pt, uint32_t want) { wait_consumed(pt, want); }

__global__ void k_round_open(PeerTable pt, uint32_t parity, uint32_t want) {
  MailboxHead* mine = pt.head[pt.me];
  wait_consumed(pt, want);
  __syncthreads();
  RoundHeader* h = &mine->hdr[parity];
  if (threadIdx.x == 0) { h->round = 0; h->n_subs = 0; h->fill = 0; h->status = 0; }
  for (uint32_t i = threadIdx.x; i < kMaxBatch * kMaxRanks; i += blockDim.x) {
    (&h->count[0][0])[i] = 0;
    (&h->offset[0][0])[i] = 0;
  }
}

// Peer-memory pack, step 1 (replaces k_list_touched): compact the window's touched bitmap into ONE LIST
// PER OWNER -- owner d's blocks at ms.list[d * seg_cap ...], counts[d] of them -- so that the records
// can be written one warp per block with no further bookkeeping.  A CTA takes 256 bitmap words at a
// time: its blocks are counted per owner in shared memory, the owners' list space is reserved with one
// global atomic per owner per CTA pass, positions inside the reservation come from shared cursors.
// (Round 2, first version: a count pass, and a write pass that reserved space with a returning atomic
// on `world` hot addresses per 8 blocks -- the pack took longer than the K4 update it replaces.)
__global__ void __launch_bounds__(256)
k_list_by_owner(MarkSet ms, Counters* cnt, uint32_t n_words, uint32_t world, uint32_t seg_cap, uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_cnt[kMaxRanks];
  __shared__ uint32_t s_cur[kMaxRanks];
  for (uint32_t wbase = blockIdx.x * blockDim.x; wbase < n_words; wbase += gridDim.x * blockDim.x) {   // uniform per CTA
    if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t wi = wbase + threadIdx.x;
    const uint32_t v = wi < n_words ? ms.touched_bits[wi] : 0u;
    for (uint32_t rest = v; rest; rest &= rest - 1u) {
      const uint32_t e = wi * 32u + (uint32_t)(__ffs(rest) - 1);
      atomicAdd(&s_cnt[owner_of(window_block_key(ms, e), world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world) s_cur[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]) : 0u;
    if (threadIdx.x == 0) {
      uint32_t total = 0;
      for (uint32_t d = 0; d < world; d++) total += s_cnt[d];
      if (total) atomicAdd(&cnt->n_touched, total);
    }
    __syncthreads();
    for (uint32_t rest = v; rest; rest &= rest - 1u) {
      const uint32_t e = wi * 32u + (uint32_t)(__ffs(rest) - 1);
      const uint32_t d = owner_of(window_block_key(ms, e), world);
      const uint32_t pos = atomicAdd(&s_cur[d], 1u);
      if (pos < seg_cap) ms.list[(size_t)d * seg_cap + pos] = e;
      else atomicExch(&cnt->overflow_keys, 1u);   // an owner's share of the list is full (sizing error: the host reports it)
    }
    __syncthreads();
  }
}

// Step 2 (one thread): bucket offsets of sub-scan `sub` inside the round buffer from counts[].  When the
// buffer cannot hold the records nothing is written, the header says so, and the scan's counters carry
// overflow_pack.
__global__ void k_pack_plan(PeerTable pt, uint32_t parity, uint32_t sub, const uint32_t* __restrict__ counts,
                            Counters* cnt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  RoundHeader* h = &pt.head[pt.me]->hdr[parity];
  if (cnt->overflow_table || cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys) { h->status = 2u; return; }
  uint32_t total = 0;
  for (uint32_t d = 0; d < pt.world; d++) total += counts[d];
  if (h->status || (unsigned long long)h->fill + total > pt.capacity) {
    h->status = 1u;
    cnt->overflow_pack = total ? total : 1u;
    return;
  }
  uint32_t at = h->fill;
  for (uint32_t d = 0; d < pt.world; d++) {
    h->count[sub][d] = counts[d];
    h->offset[sub][d] = at;
    at += counts[d];
  }
  h->fill = at;
  h->n_subs = sub + 1u;
}

// Step 3: one warp per listed block, owner by owner: the block's 128 bytes of masks become its record
// in the owner's bucket of the round buffer (same position as in the owner's list) and are cleared.
__global__ void __launch_bounds__(256)
k_pack_write_round(MarkSet ms, Counters* cnt, PeerTable pt, uint32_t parity, uint32_t sub, uint32_t seg_cap,
                   const uint32_t* __restrict__ counts) {
  if (cnt->overflow_table || cnt->overflow_window || cnt->overflow_ckpt || cnt->overflow_keys || cnt->overflow_pack) return;
  __shared__ uint32_t s_first[kMaxRanks + 1];     // global record index of every owner's first block
  __shared__ uint32_t s_off[kMaxRanks];           // first record of every owner's bucket
  const RoundHeader* hd = &pt.head[pt.me]->hdr[parity];
  if (threadIdx.x == 0) {
    uint32_t at = 0;
    for (uint32_t d = 0; d < pt.world; d++) { s_first[d] = at; at += counts[d]; }
    s_first[pt.world] = at;
  }
  if (threadIdx.x < pt.world) s_off[threadIdx.x] = hd->offset[sub][threadIdx.x];
  __syncthreads();
  BlockRecord* out = pt.rec[pt.me] + (size_t)parity * pt.capacity;
  const uint32_t total = s_first[pt.world];
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total; g += warps) {
    uint32_t d = 0;
    while (g >= s_first[d + 1]) d++;
    const uint32_t i = g - s_first[d];
    const uint32_t e = ms.list[(size_t)d * seg_cap + i];
    uint32_t* src = (lane < kMaskWords) ? ms.free_mask + (size_t)e * kMaskWords + lane
                                        : ms.occ_mask + (size_t)e * kMaskWords + lane - kMaskWords;
    const uint32_t wv = *src;
    if (wv) *src = 0u;
    BlockRecord* r = out + (s_off[d] + i);
    if (lane == 0) r->block = window_block_key(ms, e);
    if (lane < kMaskWords) r->free_mask[lane] = wv;
    else r->occ_mask[lane - kMaskWords] = wv;
  }
  // the scan's marks are packed: its touched bitmap is free for the next scan of this context
  const uint32_t n_words = (ms.nx * ms.ny * ms.nz + 31u) >> 5;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x)
    if (ms.touched_bits[i]) ms.touched_bits[i] = 0u;
}

// Seal round `round` (header of this parity) and tell every rank: one remote store per peer.
__global__ void k_round_publish(PeerTable pt, uint32_t parity, uint32_t round) {
  MailboxHead* mine = pt.head[pt.me];
  if (threadIdx.x == 0) mine->hdr[parity].round = round + 1u;
  __syncthreads();
  __threadfence_system();
  if (threadIdx.x < pt.world) *reinterpret_cast<volatile uint32_t*>(&pt.head[threadIdx.x]->ready[pt.me]) = round + 1u;
}

// Wait until every rank has published `round`, then collect what they hold for me:
// out[(j * world + s) * 2] = count, [.. + 1] = offset of (sub-scan j of rank s -> me); out[2 * kMaxBatch *
// world] = status (0 fine).  One CTA.
__global__ void k_round_collect(PeerTable pt, uint32_t parity, uint32_t round, uint32_t* __restrict__ out) {
  MailboxHead* mine = pt.head[pt.me];
  __shared__ uint32_t s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if (threadIdx.x < pt.world) {
    const long long t0 = global_ns();
    while (ld_flag(&mine->ready[threadIdx.x]) < round + 1u && !ld_flag(&mine->error)) {
      __nanosleep(256);
      if (global_ns() - t0 > kSpinLimitNs) { mine->error = 3u; break; }
    }
    if (ld_flag(&mine->error)) atomicExch(&s_bad, ld_flag(&mine->error));
  }
  __syncthreads();
  if (!s_bad) {
    for (uint32_t i = threadIdx.x; i < kMaxBatch * pt.world; i += blockDim.x) {
      const uint32_t j = i / pt.world, s = i - j * pt.world;
      const RoundHeader* h = &pt.head[s]->hdr[parity];
      const uint32_t st = ld_flag(&h->status);
      if (st) atomicExch(&s_bad, 16u + st);
      const bool live = j < ld_flag(&h->n_subs);
      out[2 * i] = live ? ld_flag(&h->count[j][pt.me]) : 0u;
      out[2 * i + 1] = live ? ld_flag(&h->offset[j][pt.me]) : 0u;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    out[2 * kMaxBatch * pt.world] = s_bad;
    if (s_bad && !mine->error) mine->error = s_bad;
  }
}

// After the round's records have been applied: every peer may reuse the buffer I read from.
__global__ void k_round_consumed(PeerTable pt, uint32_t round) {
  __threadfence_system();
  if (threadIdx.x < pt.world) *reinterpret_cast<volatile uint32_t*>(&pt.head[threadIdx.x]->consumed[pt.me]) = round + 1u;
}

// Sharded getLineStatus: every rank walks every segment but only probes the voxels it owns;
// the result is (index of the first non-free owned voxel << 2) | status, 0xFFFFFFFF when none.
// The global answer is the minimum over ranks (then status = value & 3, or free).
__global__ void __launch_bounds__(128)
k_query_lines_sharded(DevParams P, MapView m, const double* __restrict__ ab, uint32_t n,
                      uint32_t* __restrict__ packed, uint32_t rank, uint32_t world) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ax = (float)ab[6 * i], ay = (float)ab[6 * i + 1], az = (float)ab[6 * i + 2];
  const float bx = (float)ab[6 * i + 3], by = (float)ab[6 * i + 4], bz = (float)ab[6 * i + 5];
  uint32_t out = 0xFFFFFFFFu, idx = 0;
  unsigned long long cached_b = kEmpty;
  uint32_t cached_h = kNotFound;
  bool mine = false;
  walk_ray(P, ax, ay, az, bx, by, bz, [&](uint32_t kx, uint32_t ky, uint32_t kz) {
    const unsigned long long b = block_key(kx, ky, kz);
    if (b != cached_b) {
      cached_b = b;
      mine = owner_of(b, world) == rank;
      cached_h = mine ? find_block(m, b) : kNotFound;
    }
    if (mine) {
      const uint8_t s = voxel_status(P, m, cached_h, kx, ky, kz);
      if (s != 0) { out = (idx << 2) | s; return false; }
    }
    ++idx;
    return true;
  });
  packed[i] = out;
}

// Sharded getCellStatusPoint: the owner answers, everybody else writes 0xFF.
__global__ void __launch_bounds__(256)
k_query_points_sharded(DevParams P, MapView m, const double* __restrict__ xyz, uint32_t n,
                       uint8_t* __restrict__ status, uint32_t rank, uint32_t world) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s0 = scaled_coord(P.res_factor, xyz[3 * i]);
  const int s1 = scaled_coord(P.res_factor, xyz[3 * i + 1]);
  const int s2 = scaled_coord(P.res_factor, xyz[3 * i + 2]);
  if (!(key_in_range(s0) && key_in_range(s1) && key_in_range(s2))) {
    // search() returns NULL on every rank: rank 0 reports it
    status[i] = (rank == 0) ? (P.treat_unknown_as_occupied ? 1 : 2) : 0xFF;
    return;
  }
  const unsigned long long b = block_key(s0, s1, s2);
  if (owner_of(b, world) != rank) { status[i] = 0xFF; return; }
  status[i] = voxel_status(P, m, find_block(m, b), s0, s1, s2);
}

}  // namespace vmapdev
