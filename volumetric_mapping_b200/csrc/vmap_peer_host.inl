// vmap_peer_host.inl -- sharded map build over PEER MEMORY (included by vmap.cu after
// vmap_shard_host.inl).  Same partitioning and ordering as the NCCL path (rank r generates the
// r-th scan of every round, every owner applies a round's scans in scan order), but nothing is sent:
//
//   generate  scan k runs through the scan pipeline's three contexts (K1..K2 on its own window);
//             its touched blocks are packed into this rank's ROUND BUFFER of the round's parity,
//             bucketed by owner, bucket offsets kept in a device-side header (vmap_shard.cuh).
//   publish   after the round's last pack: one remote store per peer (k_round_publish).
//   collect   one warp waits for every peer's store and gathers (count, offset) of what each peer
//             holds for this rank; the host reads those 2 * M * G words to size table and pool
//             exactly -- the only host read of a round, taken one round late, behind the next
//             round's generate kernels.
//   apply     k_apply_update per (scan, source rank), reading the records straight out of the
//             source rank's round buffer over NVLink (CUDA IPC mapping) -- the all-to-all is the
//             apply kernel's own loads; then one remote store per peer: "your buffer is free".
//
// Scan order per voxel (octomap_world.cc:244-262) holds because a block has one owner, the owner
// applies list after list on one stream, and list order is (sub-scan, source rank) = global scan order.
namespace {

constexpr size_t kRoundInfoWords = 2 * (size_t)kMaxBatch * kMaxRanks + 1;

int peer_alloc_mailbox(vmap* h, uint64_t records_per_round) {
  if (records_per_round < 1024) records_per_round = 1024;
  if (records_per_round > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "round buffer too large");
  h->peer_capacity = (uint32_t)records_per_round;
  h->mailbox_bytes = sizeof(MailboxHead) + 2 * (size_t)records_per_round * sizeof(BlockRecord);
  VM_TRY(dev_alloc(h, &h->mailbox, h->mailbox_bytes));
  VM_CUDA(h, cudaMemset(h->mailbox, 0, sizeof(MailboxHead)));
  VM_TRY(dev_alloc(h, (void**)&h->d_round_info, 2 * kRoundInfoWords * sizeof(uint32_t)));
  VM_CUDA(h, cudaMallocHost((void**)&h->h_round_info, 2 * kRoundInfoWords * sizeof(uint32_t)));
  std::memset(h->h_round_info, 0, 2 * kRoundInfoWords * sizeof(uint32_t));
  VM_TRY(dev_alloc(h, (void**)&h->d_sub_counts, 3 * 2 * kMaxRanks * sizeof(uint32_t)));   // one set per scan context
  VM_TRY(dev_alloc(h, (void**)&h->d_lists, sizeof(RecordLists)));
  VM_CUDA(h, cudaMallocHost((void**)&h->h_lists, sizeof(RecordLists)));
  for (auto& e : h->ev_collect) VM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return VMAP_OK;
}

// VMAP_TIMING: device time of the apply round that has just been waited for (events on the apply stream)
void peer_note_apply_time(vmap* h) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev_app0, h->ev_app1) != cudaSuccess) { cudaGetLastError(); return; }
  h->dbg_ms[0] += ms;
  h->dbg_n++;
}

void peer_set_table(vmap* h, int s, void* base) {
  h->pt.head[s] = reinterpret_cast<MailboxHead*>(base);
  h->pt.rec[s] = reinterpret_cast<BlockRecord*>(reinterpret_cast<char*>(base) + sizeof(MailboxHead));
}

uint64_t peer_default_records() {
  if (const char* e = getenv("VMAP_DIST_ROUND_RECORDS")) {
    const long long v = atoll(e);
    if (v > 0) return (uint64_t)v;
  }
  return 3ull << 20;   // 3 Mi records = 408 MiB per round buffer: 8 cfg3 scans or 4 cfg4 scans twice over
}

void peer_release(vmap* h) {
  h->peer_on = false;          // (frees are immediate again: the caller has drained the device)
  empty_graveyard(h);
  if (!h->mailbox) return;
  for (int s = 0; s < kMaxRanks; s++)
    if (h->peer_ipc_base[s]) { cudaIpcCloseMemHandle(h->peer_ipc_base[s]); h->peer_ipc_base[s] = nullptr; }
  dev_free(h, h->mailbox, h->mailbox_bytes);
  h->mailbox = nullptr;
  cudaFree(h->d_round_info); h->d_round_info = nullptr;
  cudaFree(h->d_sub_counts); h->d_sub_counts = nullptr;
  if (h->d_lists) { cudaFree(h->d_lists); h->dev_bytes -= sizeof(RecordLists); }
  h->d_lists = nullptr;
  if (h->h_lists) cudaFreeHost(h->h_lists);
  h->h_lists = nullptr;
  if (h->d_located) { cudaFree(h->d_located); h->dev_bytes -= (size_t)h->located_cap * sizeof(uint32_t); }
  h->d_located = nullptr;
  h->located_cap = 0;
  if (h->h_round_info) cudaFreeHost(h->h_round_info);
  h->h_round_info = nullptr;
  for (auto& e : h->ev_collect)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  h->peer_on = false;
}

// Retire the oldest generated scan of a sharded handle: nothing to replay here -- the scan's records
// are in the round buffer; only its counters are taken and its failure flags turned into errors.
int peer_retire(vmap* h) {
  if (h->pipe_n <= 0) return VMAP_OK;
  {
    ScanCtx* oldest = other_in_flight(h, false);
    if (oldest && (!h->in_flight || oldest->seq < h->seq)) make_current(h, oldest);
  }
  VM_CUDA(h, cudaEventSynchronize(h->ev_done));
  h->in_flight = false;
  h->pipe_n--;
  const Counters& c = *h->h_cnt;
  if (c.overflow_pack)
    return fail(h, VMAP_ERR_CAPACITY, "sharded insertion: the round buffer cannot hold the round's block records (raise VMAP_DIST_ROUND_RECORDS or lower the batch)");
  if (c.overflow_window || c.overflow_ckpt || c.overflow_keys || c.overflow_table)
    return fail(h, VMAP_ERR_DIST, "sharded insertion: a scan's marks were incomplete (internal sizing error)");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  h->stats = vmap_scan_stats{};
  h->stats.points_in = h->meta.n;
  fill_stats(h, 0, ms);
  return VMAP_OK;
}

// Apply every round that is published but not applied yet.  Blocks on the round's collect event
// (every rank has published it), sizes table and pool for the records about to arrive, enqueues
// the applies and the "consumed" stores on the apply stream.
int peer_apply_pending(vmap* h) {
  cudaStream_t as = h->app_stream;
  const int G = h->shard_world;
  while (h->rounds_applied < h->rounds_published) {
    const uint64_t r = h->rounds_applied;
    const uint32_t p = (uint32_t)(r & 1u);
    VM_CUDA(h, cudaEventSynchronize(h->ev_collect[p]));
    const uint32_t* info = h->h_round_info + p * kRoundInfoWords;
    const uint32_t status = info[2 * kMaxBatch * G];
    if (status) {
      char msg[320];
      snprintf(msg, sizeof(msg), "sharded insertion: round %llu failed on the device (code %u: 1-3 a peer did not answer in time, 17 a peer's round buffer overflowed, 18 a peer's marks were incomplete)",
               (unsigned long long)r, status);
      return fail(h, status == 17 ? VMAP_ERR_CAPACITY : VMAP_ERR_DIST, msg);
    }
    size_t total = 0;
    for (int i = 0; i < kMaxBatch * G; i++) total += info[2 * i];
    if (total > 0x7FFFFFFFull) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "too many block records");
    // the previous round's apply is the only thing that can still be running on the apply stream
    VM_CUDA(h, cudaStreamSynchronize(as));
    if (h->app_pending) {
      h->app_pending = false;
      if (h->h_cnt_app->overflow_table || h->h_cnt_app->overflow_pool)
        return fail(h, VMAP_ERR_CAPACITY, "internal: map overflow inside a sharded apply round");
      *h->h_persist = *h->h_persist_app;
      peer_note_apply_time(h);
      h->totals.unique_free += h->h_cnt_app->unique_free;
      h->totals.unique_occupied += h->h_cnt_app->unique_occupied;
    }
    while (((uint64_t)h->h_persist->n_entries + total) * 2 > h->table_cap) VM_TRY(grow_table(h));
    if ((uint64_t)h->h_persist->n_slots + total > h->pool_cap) VM_TRY(grow_pool(h, (uint64_t)h->h_persist->n_slots + total));
    VM_CUDA(h, cudaMemsetAsync(h->d_cnt_app, 0, sizeof(Counters), as));
    VM_CUDA(h, cudaEventRecord(h->ev_app0, as));
    // every record of the round is located in one launch (table entry, pool slot), then the lists are
    // applied one launch each, in scan order = (sub-scan, source rank)
    RecordLists* hl = h->h_lists;                    // pinned; free again: the apply stream was drained above
    uint32_t n_lists = 0;
    hl->first[0] = 0;
    for (int j = 0; j < kMaxBatch; j++)
      for (int s = 0; s < G; s++) {
        const uint32_t cnt = info[2 * (j * G + s)], off = info[2 * (j * G + s) + 1];
        if (!cnt) continue;
        hl->ptr[n_lists] = h->pt.rec[s] + (size_t)p * h->peer_capacity + off;
        hl->first[n_lists + 1] = hl->first[n_lists] + cnt;
        n_lists++;
      }
    hl->n_lists = n_lists;
    if (n_lists) {
      // round-fused apply (k_apply_round) for up to 32 lists, list by list beyond that; VMAP_PEER_FUSED=0 forces the latter
      static const bool fused_ok = !(getenv("VMAP_PEER_FUSED") && atoi(getenv("VMAP_PEER_FUSED")) == 0);
      const bool fused = fused_ok && n_lists <= 32;
      const size_t need = fused ? 3 * total + total * n_lists : total;     // located | round_blocks | bcount | bucket
      if (need > h->located_cap) {
        dev_free(h, h->d_located, (size_t)h->located_cap * sizeof(uint32_t));
        h->d_located = nullptr;
        h->located_cap = 0;
        const size_t want = need + (need >> 2);
        VM_TRY(dev_alloc(h, (void**)&h->d_located, want * sizeof(uint32_t)));
        h->located_cap = want;
      }
      uint32_t* d_round_blocks = h->d_located + total;
      uint32_t* d_bcount = h->d_located + 2 * total;
      uint32_t* d_bucket = h->d_located + 3 * total;
      VM_CUDA(h, cudaMemcpyAsync(h->d_lists, hl, sizeof(RecordLists), cudaMemcpyHostToDevice, as));
      if (fused) {
        // m.touched (the table-mode block list, unused by sharded handles' scans) serves as "index of this
        // block among the round's blocks": all ones = not seen yet in this round
        VM_CUDA(h, cudaMemsetAsync(h->m.touched, 0xFF, h->table_cap * sizeof(uint32_t), as));
        VM_CUDA(h, cudaMemsetAsync(d_bcount, 0, total * sizeof(uint32_t), as));
      }
      static const bool split = getenv("VMAP_PEER_SERIAL") && atoi(getenv("VMAP_PEER_SERIAL")) && getenv("VMAP_TIMING");
      if (split) {
        for (auto& e : h->ev_split)
          if (!e) VM_CUDA(h, cudaEventCreate(&e));
        VM_CUDA(h, cudaEventRecord(h->ev_split[0], as));
      }
      const unsigned lgrid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)g_sm_count(h->device) * 8);
      k_locate_records<<<std::max(lgrid, 1u), 256, 0, as>>>(h->m, h->d_cnt_app, h->d_lists, h->d_located, fused ? d_round_blocks : nullptr);
      h->launches++;
      if (split) VM_CUDA(h, cudaEventRecord(h->ev_split[1], as));
      if (fused) {
        k_bucket_records<<<(unsigned)((total + 255) / 256), 256, 0, as>>>(h->m, h->d_located, (uint32_t)total, n_lists, d_bcount, d_bucket);
        if (split) VM_CUDA(h, cudaEventRecord(h->ev_split[2], as));
        const unsigned grid = (unsigned)std::min<size_t>((total + 7) / 8, (size_t)g_sm_count(h->device) * 8);
        k_apply_round<<<std::max(grid, 1u), 256, 0, as>>>(h->P, h->m, h->d_cnt_app, h->d_lists, d_round_blocks, d_bcount, d_bucket);
        h->launches += 2;
        if (split) {
          VM_CUDA(h, cudaEventRecord(h->ev_split[3], as));
          VM_CUDA(h, cudaEventSynchronize(h->ev_split[3]));
          for (int k = 0; k < 3; k++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev_split[k], h->ev_split[k + 1]);
            h->dbg_ms[1 + k] += ms;
          }
        }
      } else {
        for (uint32_t l = 0; l < n_lists; l++) {
          const uint32_t cnt = hl->first[l + 1] - hl->first[l];
          const unsigned grid = (unsigned)std::min<size_t>(((size_t)cnt + 7) / 8, (size_t)g_sm_count(h->device) * 8);
          k_apply_located<<<std::max(grid, 1u), 256, 0, as>>>(h->P, h->m, h->d_cnt_app, hl->ptr[l], h->d_located + hl->first[l], cnt);
          h->launches++;
        }
      }
    }
    k_round_consumed<<<1, 64, 0, as>>>(h->pt, (uint32_t)r);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
    VM_CUDA(h, cudaEventRecord(h->ev_app1, as));
    VM_CUDA(h, cudaMemcpyAsync(h->h_cnt_app, h->d_cnt_app, sizeof(Counters), cudaMemcpyDeviceToHost, as));
    VM_CUDA(h, cudaMemcpyAsync(h->h_persist_app, h->m.persist, sizeof(Persist), cudaMemcpyDeviceToHost, as));
    h->app_pending = true;
    h->app_records += total;
    h->rounds_applied++;
  }
  return VMAP_OK;
}

// The round under construction is complete on this rank: apply what is pending (one round late, so
// that it overlaps this round's generate kernels), then publish and collect this round.
int peer_seal(vmap* h) {
  if (h->sub_open == 0) return VMAP_OK;
  cudaStream_t as = h->app_stream;
  const uint64_t r = h->round_open;
  const uint32_t p = (uint32_t)(r & 1u);
  VM_TRY(peer_apply_pending(h));
  // after the round's packs (the two scan contexts' streams; earlier sub-scans precede these in stream order)
  if (h->in_flight) VM_CUDA(h, cudaStreamWaitEvent(as, h->ev_done, 0));
  if (h->alt.in_flight) VM_CUDA(h, cudaStreamWaitEvent(as, h->alt.ev_done, 0));
  if (h->alt2.in_flight) VM_CUDA(h, cudaStreamWaitEvent(as, h->alt2.ev_done, 0));
  k_round_publish<<<1, 64, 0, as>>>(h->pt, p, (uint32_t)r);
  k_round_collect<<<1, 128, 0, as>>>(h->pt, p, (uint32_t)r, h->d_round_info + p * kRoundInfoWords);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_CUDA(h, cudaMemcpyAsync(h->h_round_info + p * kRoundInfoWords, h->d_round_info + p * kRoundInfoWords,
                             kRoundInfoWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, as));
  VM_CUDA(h, cudaEventRecord(h->ev_collect[p], as));
  h->rounds_published = r + 1;
  h->round_open = r + 1;
  h->sub_open = 0;
  return VMAP_OK;
}

// One collective insert over peer memory (see the head of this file).
int peer_insert(vmap* h, const void* src, bool src_on_host, size_t n, size_t stride, int is_dense, const double Tm[16]) {
  if (h->capture) return fail(h, VMAP_ERR_INVALID_ARGUMENT, "key capture is not available with the peer-memory transport (VMAP_DIST_TRANSPORT=nccl)");
  VM_TRY(init_app_stream(h));
  const int n_ctx = h->pipe_ctx >= 3 ? 3 : 2;      // scan contexts the generate side rotates through
  VM_TRY(pipe_prepare(h, n_ctx == 3));
  while (h->pipe_n >= n_ctx) VM_TRY(peer_retire(h));
  if (h->in_flight) {                               // the current context is a free one from here on
    if (!h->alt.in_flight) swap_ctx(h);
    else swap_ctx2(h);
  }
  if (h->epoch >= kEpochMax) {
    while (h->pipe_n > 0) VM_TRY(peer_retire(h));
  }
  Transform T;
  fill_transform_matrix(&T, Tm);
  const size_t n1 = std::max<size_t>(n, 1);
  VM_TRY(ensure_scan(h, n1));
  const uint64_t ck = checkpoint_bound(h, n1);
  if (ck >= 0xFFFFFFFFull) return fail(h, VMAP_ERR_CAPACITY, "scan too large for the peer-memory transport");
  VM_TRY(ensure_ckpt(h, ck));
  if (!setup_dense_marks(h, T.origin))
    return fail(h, VMAP_ERR_INVALID_ARGUMENT, "the peer-memory transport needs the dense mark window (finite sensor_max_range, window within vmap_config.dense_window_bytes); use VMAP_DIST_TRANSPORT=nccl");
  if (h->win.list_cap < h->win.n_blocks) {
    VM_TRY(grow_window_list(h, (uint32_t)std::min<uint64_t>(h->win.n_blocks, 0xFFFFFFFFull)));
    h->ms.list = h->win.list; h->ms.locate = h->win.locate; h->ms.list_cap = h->win.list_cap;
    if (h->win.list_cap < h->win.n_blocks) return fail(h, VMAP_ERR_CAPACITY, "window block list could not be sized");
  }
  if (n) {
    const size_t bytes = (n - 1) * stride + 12;
    VM_TRY(ensure(h, &h->in, bytes));
    VM_CUDA(h, cudaMemcpyAsync(h->in.p, src, bytes, src_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
  }
  VM_CUDA(h, cudaEventRecord(h->ev_h2d, h->stream));
  if (h->l2_flush_bytes)
    VM_CUDA(h, cudaMemsetAsync(h->l2_flush.p, (int)(h->pipe_seq & 0xFF), h->l2_flush_bytes, h->stream));
  h->walk_overlaps = true;
  VM_TRY(next_epoch(h));
  VM_TRY(begin_scan_scratch(h, n1));
  for (bool& b : h->evk_set) b = false;
  VM_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  VM_TRY(mark_stage(h, 0));
  k_front_cloud<<<grid_for(n, 256), 256, 0, h->stream>>>(h->P, h->sb, h->d_cnt, T, (const uint8_t*)h->in.p, (uint32_t)n,
                                                        (uint32_t)stride, is_dense);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_TRY(launch_resolve_and_walk(h, n, T));
  // the touched blocks, listed owner by owner (k_list_by_owner instead of k_list_touched)
  uint32_t* counts = h->d_sub_counts + (size_t)h->ctx_id * 2 * kMaxRanks;
  VM_CUDA(h, cudaMemsetAsync(counts, 0, 2 * kMaxRanks * sizeof(uint32_t), h->stream));
  const uint32_t seg_cap = h->win.list_cap / (uint32_t)h->shard_world;
  {
    const unsigned grid = (unsigned)std::min<uint64_t>((h->win.n_words + 255) / 256, (uint64_t)g_sm_count(h->device) * 8);
    k_list_by_owner<<<std::max(grid, 1u), 256, 0, h->stream>>>(h->ms, h->d_cnt, h->win.n_words, (uint32_t)h->shard_world, seg_cap, counts);
    h->launches++;
    VM_CUDA(h, cudaGetLastError());
  }
  VM_TRY(mark_stage(h, 3));
  // the ordered half: this scan's pack comes after the pack of the scan before it (they share the
  // round buffer's cursor)
  if (ScanCtx* prev = other_in_flight(h, true)) VM_CUDA(h, cudaStreamWaitEvent(h->stream, prev->ev_done, 0));
  const uint64_t r = h->round_open;
  const uint32_t p = (uint32_t)(r & 1u), j = (uint32_t)h->sub_open;
  if (j == 0) {
    k_round_open<<<1, 64, 0, h->stream>>>(h->pt, p, r >= 2 ? (uint32_t)(r - 1) : 0u);   // round r-2, stored as number + 1
    h->launches++;
  }
  k_pack_plan<<<1, 32, 0, h->stream>>>(h->pt, p, j, counts, h->d_cnt);
  k_pack_write_round<<<g_sm_count(h->device) * 8, 256, 0, h->stream>>>(h->ms, h->d_cnt, h->pt, p, j, seg_cap, counts);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_TRY(mark_stage(h, 4));
  VM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  VM_CUDA(h, cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  VM_CUDA(h, cudaEventRecord(h->ev_done, h->stream));
  h->in_flight = true;
  h->seq = ++h->pipe_seq;
  h->pipe_n++;
  h->meta.n = n;
  h->meta.stride = stride;
  h->meta.is_dense = is_dense;
  std::memcpy(h->meta.Tm, Tm, sizeof(h->meta.Tm));
  h->sub_open++;
  // VMAP_PEER_SERIAL=1 (measurement only): nothing overlaps, so the stage events time each kernel group alone
  static const bool serial = getenv("VMAP_PEER_SERIAL") && atoi(getenv("VMAP_PEER_SERIAL"));
  if (serial) VM_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->sub_open >= h->batch_m) VM_TRY(peer_seal(h));
  if (serial && h->sub_open == 0) {
    VM_TRY(peer_apply_pending(h));
    VM_CUDA(h, cudaStreamSynchronize(h->app_stream));
  }
  VM_CUDA(h, cudaEventSynchronize(h->ev_h2d));
  return VMAP_OK;
}

// Collective flush, first half: seal and publish a partial batch (never blocks on another rank).
int peer_flush_begin(vmap* h) {
  if (h->sub_open > 0) {
    VM_TRY(init_app_stream(h));
    VM_TRY(peer_seal(h));
  }
  return VMAP_OK;
}

// Second half: apply everything published, wait until every rank has applied MY last round (so that
// nobody still reads my round buffers), retire the generated scans.
int peer_flush_end(vmap* h) {
  int first = VMAP_OK;
  std::string msg;
  auto note = [&](int st) { if (st != VMAP_OK && first == VMAP_OK) { first = st; msg = h->err; } };
  while (h->pipe_n > 0) note(peer_retire(h));
  if (h->app_stream && h->peer_on) {
    cudaStream_t as = h->app_stream;
    note(peer_apply_pending(h));
    if (h->rounds_published) {
      k_wait_consumed<<<1, 64, 0, as>>>(h->pt, (uint32_t)h->rounds_published);
      h->launches++;
    }
    // the mailbox's error word rides on the same synchronisation
    uint32_t err = 0;
    cudaMemcpyAsync(&err, &h->pt.head[h->shard_rank]->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, as);
    if (cudaStreamSynchronize(as) != cudaSuccess) note(fail(h, VMAP_ERR_CUDA, "cudaStreamSynchronize(apply stream)"));
    if (h->app_pending) {
      h->app_pending = false;
      if (h->h_cnt_app->overflow_table || h->h_cnt_app->overflow_pool)
        note(fail(h, VMAP_ERR_CAPACITY, "internal: map overflow inside a sharded apply round"));
      *h->h_persist = *h->h_persist_app;
      peer_note_apply_time(h);
      h->totals.unique_free += h->h_cnt_app->unique_free;
      h->totals.unique_occupied += h->h_cnt_app->unique_occupied;
      h->stats.blocks_total = h->h_persist->n_slots;
      h->stats.blocks_applied = h->app_records;
    }
    if (err) note(fail(h, VMAP_ERR_DIST, "sharded insertion: a rank did not answer in time (mailbox error word set)"));
    // quiescent: nothing of this rank is in flight and nobody reads its round buffers.  (Not with
    // vmap_dist_attach_local: the other handles' spinners are kernels of this very context.)
    if (!h->peer_local) empty_graveyard(h);
  }
  if (first != VMAP_OK) h->err = msg;
  return first;
}

}  // namespace
