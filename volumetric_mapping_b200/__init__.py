"""volumetric_mapping_b200: B200-native (sm_100a) scan insertion and collision queries for
volumetric occupancy mapping, behind the OctomapWorld API of ethz-asl/volumetric_mapping.

The package is a thin host-side mirror over libvmap_b200.so (include/vmap/vmap.h); the CUDA
library is the product and there is no CPU fallback."""
from ._capi import Config, Params, ScanStats, VmapError, default_config, default_params  # noqa: F401
from .world import (DEDUPE_BITMASK, DEDUPE_SORT, OctomapWorld, kFree, kOccupied,  # noqa: F401
                    kUnknown, quat_to_matrix)
