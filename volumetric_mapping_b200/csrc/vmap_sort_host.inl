// vmap_sort_host.inl -- host launch sequence of the VMAP_DEDUPE_SORT mode (included by vmap.cu).
namespace {

void sort_free(vmap* h) {
  SortBuffers& s = h->sort;
  dev_free(h, s.keys_a, s.cap * 8);
  dev_free(h, s.keys_b, s.cap * 8);
  dev_free(h, s.ray_off, s.ray_cap * 4);
  dev_free(h, s.hist, s.hist_cap * 4);
  dev_free(h, s.bin_total, kRadix * 4);
  s = SortBuffers{};
}

int sort_ensure(vmap* h, size_t n_keys, size_t n_rays) {
  SortBuffers& s = h->sort;
  if (!s.bin_total) VM_TRY(dev_alloc(h, (void**)&s.bin_total, kRadix * 4));
  if (n_rays > s.ray_cap) {
    VM_CUDA(h, cudaStreamSynchronize(h->stream));
    dev_free(h, s.ray_off, s.ray_cap * 4);
    s.ray_off = nullptr;
    s.ray_cap = 0;
    const size_t cap = n_rays + (n_rays >> 2) + 1024;
    VM_TRY(dev_alloc(h, (void**)&s.ray_off, cap * 4));
    s.ray_cap = cap;
  }
  if (n_keys > s.cap) {
    VM_CUDA(h, cudaStreamSynchronize(h->stream));
    dev_free(h, s.keys_a, s.cap * 8);
    dev_free(h, s.keys_b, s.cap * 8);
    s.keys_a = s.keys_b = nullptr;
    s.cap = 0;
    const size_t cap = std::max<size_t>(n_keys + (n_keys >> 2), std::max<uint64_t>(h->cfg.initial_scan_keys, 1 << 16));
    if (cap > 0xFFFFFFF0ull) return fail(h, VMAP_ERR_CAPACITY, "sort mode: more than 2^32 keys in one scan");
    VM_TRY(dev_alloc(h, (void**)&s.keys_a, cap * 8));
    VM_TRY(dev_alloc(h, (void**)&s.keys_b, cap * 8));
    s.cap = cap;
  }
  const size_t tiles = (std::max<size_t>(n_keys, 1) + kSortTile - 1) / kSortTile;
  if (tiles * kRadix > s.hist_cap) {
    VM_CUDA(h, cudaStreamSynchronize(h->stream));
    dev_free(h, s.hist, s.hist_cap * 4);
    s.hist = nullptr;
    s.hist_cap = 0;
    const size_t cap = (tiles + (tiles >> 2) + 16) * kRadix;
    VM_TRY(dev_alloc(h, (void**)&s.hist, cap * 4));
    s.hist_cap = cap;
  }
  return VMAP_OK;
}

int launch_resolve_and_sort(vmap* h, size_t n, const Transform& T) {
  SortBuffers& s = h->sort;
  VM_TRY(sort_ensure(h, 0, n));
  VM_TRY(mark_stage(h, 1));
  k_resolve<<<grid_for(n, 256), 256, 0, h->stream>>>(h->sb, h->m, h->ms, h->d_cnt, (uint32_t)n, 0);
  k_order_rays<<<grid_for(n, 256), 256, 0, h->stream>>>(h->sb, (uint32_t)n);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  VM_TRY(mark_stage(h, 2));
  // emit, pass 1: per-ray key counts and slots
  if (h->P.free_filter)
    k_walk_count<true><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->d_cnt, s.ray_off, T.origin[0], T.origin[1], T.origin[2]);
  else
    k_walk_count<false><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->d_cnt, s.ray_off, T.origin[0], T.origin[1], T.origin[2]);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  VM_TRY(read_counters(h));   // n_keys, occupied_emitted: the key buffers are sized exactly
  const size_t n_free = h->h_cnt->n_keys, n_occ = h->h_cnt->occupied_emitted;
  const size_t n_keys = n_free + n_occ;
  if (n_keys == 0) return VMAP_OK;
  VM_TRY(sort_ensure(h, n_keys, n));
  // emit, pass 2
  if (h->P.free_filter)
    k_walk_emit<true><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->d_cnt, s.ray_off, s.keys_a, (uint32_t)s.cap, T.origin[0], T.origin[1], T.origin[2]);
  else
    k_walk_emit<false><<<grid_for(n, 128), 128, 0, h->stream>>>(h->P, h->sb, h->d_cnt, s.ray_off, s.keys_a, (uint32_t)s.cap, T.origin[0], T.origin[1], T.origin[2]);
  k_emit_occupied<<<grid_for(n, 256), 256, 0, h->stream>>>(h->sb, h->d_cnt, (uint32_t)n, s.keys_a, (uint32_t)n_free, (uint32_t)s.cap);
  h->launches += 2;
  VM_CUDA(h, cudaGetLastError());
  // radix sort, 7 x 8 bits
  const uint32_t tiles = (uint32_t)((n_keys + kSortTile - 1) / kSortTile);
  unsigned long long* src = s.keys_a;
  unsigned long long* dst = s.keys_b;
  for (int pass = 0; pass < kSortPasses; pass++) {
    const int shift = pass * kRadixBits;
    k_radix_hist<<<tiles, kSortThreads, 0, h->stream>>>(src, (uint32_t)n_keys, shift, s.hist, tiles);
    k_radix_scan_tiles<<<kRadix, 256, 0, h->stream>>>(s.hist, tiles, s.bin_total);
    k_radix_scatter<<<tiles, kSortThreads, 0, h->stream>>>(src, dst, (uint32_t)n_keys, shift, s.hist, tiles, s.bin_total);
    h->launches += 3;
    VM_CUDA(h, cudaGetLastError());
    std::swap(src, dst);
  }
  // unique (run heads) with occupied-wins -> block masks
  k_unique_mark<<<grid_for(n_keys, 256), 256, 0, h->stream>>>(h->m, h->d_cnt, src, (uint32_t)n_keys);
  h->launches++;
  VM_CUDA(h, cudaGetLastError());
  return VMAP_OK;
}

}  // namespace
