// vmap_sort.cuh -- VMAP_DEDUPE_SORT key-set construction (placeholder until the sort kernels land).
#pragma once
#include "vmap_kernels.cuh"
namespace vmapdev {
struct SortBuffers {
  int unused;
};
}  // namespace vmapdev
