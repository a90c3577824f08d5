// vmap_sort_host.inl -- host launch sequence of the VMAP_DEDUPE_SORT mode (included by vmap.cu).
namespace {
void sort_free(vmap*) {}
int launch_resolve_and_sort(vmap* h, size_t, const Transform&) {
  return fail(h, VMAP_ERR_INVALID_ARGUMENT, "VMAP_DEDUPE_SORT is not available in this build");
}
}  // namespace
