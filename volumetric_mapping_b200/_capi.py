"""ctypes binding of include/vmap/vmap.h (libvmap_b200.so).  Loading never falls back to
anything else: a missing library or a missing CUDA device is an error."""
import ctypes as C
import os

from . import build as _build

_lib = None


class Params(C.Structure):
    """vmap_params == volumetric_mapping::OctomapParameters (octomap_world.h:54-109)."""
    _fields_ = [
        ("resolution", C.c_double),
        ("probability_hit", C.c_double),
        ("probability_miss", C.c_double),
        ("threshold_min", C.c_double),
        ("threshold_max", C.c_double),
        ("threshold_occupancy", C.c_double),
        ("filter_speckles", C.c_int32),
        ("_pad0", C.c_int32),
        ("max_free_space", C.c_double),
        ("min_height_free_space", C.c_double),
        ("sensor_max_range", C.c_double),
        ("visualize_min_z", C.c_double),
        ("visualize_max_z", C.c_double),
        ("treat_unknown_as_occupied", C.c_int32),
        ("change_detection_enabled", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("dedupe_mode", C.c_int32),
        ("initial_blocks", C.c_uint64),
        ("max_blocks", C.c_uint64),
        ("initial_scan_points", C.c_uint64),
        ("initial_scan_keys", C.c_uint64),
        ("dense_window_bytes", C.c_uint64),
    ]


class ScanStats(C.Structure):
    _fields_ = [
        ("points_in", C.c_uint64),
        ("points_valid", C.c_uint64),
        ("points_in_range", C.c_uint64),
        ("rays_cast", C.c_uint64),
        ("traversals", C.c_uint64),
        ("occupied_emitted", C.c_uint64),
        ("unique_free", C.c_uint64),
        ("unique_occupied", C.c_uint64),
        ("longest_ray", C.c_uint64),
        ("blocks_touched", C.c_uint64),
        ("blocks_total", C.c_uint64),
        ("retries", C.c_uint64),
        ("gpu_ms", C.c_double),
        ("front_ms", C.c_double),
        ("resolve_ms", C.c_double),
        ("walk_ms", C.c_double),
        ("update_ms", C.c_double),
        ("records_sent", C.c_uint64),
        ("records_received", C.c_uint64),
        ("blocks_applied", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/vmap/vmap.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "vmap_version", "vmap_default_params", "vmap_default_config", "vmap_create", "vmap_destroy",
    "vmap_reset", "vmap_set_params", "vmap_get_params", "vmap_last_error",
    "vmap_insert_pointcloud", "vmap_insert_pointcloud_dev", "vmap_insert_projected",
    "vmap_insert_disparity", "vmap_insert_disparity_dev", "vmap_reproject", "vmap_apply_keys",
    "vmap_query_points", "vmap_query_points_dev", "vmap_query_lines", "vmap_query_lines_dev",
    "vmap_leaf_count", "vmap_export_leaves", "vmap_import_leaves", "vmap_last_scan_stats",
    "vmap_set_debug_capture", "vmap_debug_last_scan_keys", "vmap_logodds_constants",
    "vmap_stream", "vmap_kernel_launches", "vmap_device_bytes", "vmap_last_walk_kernel_ms",
    "vmap_set_shard", "vmap_shard_owner", "vmap_shard_record_bytes", "vmap_shard_generate",
    "vmap_shard_generate_dev", "vmap_shard_records", "vmap_shard_apply_dev",
    "vmap_shard_query_points_dev", "vmap_shard_query_lines_dev", "vmap_dist_unique_id",
    "vmap_dist_init", "vmap_dist_shutdown", "vmap_dist_insert_pointcloud",
    "vmap_dist_insert_pointcloud_dev", "vmap_dist_query_points_dev", "vmap_dist_query_lines_dev",
    "vmap_dist_flush", "vmap_dist_set_batch", "vmap_set_debug_l2_flush", "vmap_query_points_probability",
    "vmap_query_visibility", "vmap_query_lines_bbox", "vmap_tree_size", "vmap_serialize_bt", "vmap_write_bt", "vmap_load_bt", "vmap_deserialize_bt", "vmap_serialize_ot", "vmap_deserialize_ot", "vmap_set_log_odds_bbox", "vmap_get_changed", "vmap_host_tree_stream", "vmap_host_parse_stream",
    "vmap_query_bbox_status", "vmap_query_bbox_status_dev", "vmap_check_path_collisions",
    "vmap_stage_pointcloud", "vmap_insert_staged",
    "vmap_load_occupied_pointcloud_file", "vmap_insert_occupied_points", "vmap_host_read_pointcloud_file",
    "vmap_leaf_digest", "vmap_query_lines_probes_dev",
    "vmap_dist_attach_local", "vmap_dist_flush_begin",
    "vmap_export_leaves_dev", "vmap_import_leaves_dev",
    "vmap_flush", "vmap_set_pipeline_depth", "vmap_totals", "vmap_region_begin", "vmap_region_end",
]

STATUS_NAMES = {0: "VMAP_OK", 1: "VMAP_ERR_INVALID_ARGUMENT", 2: "VMAP_ERR_CUDA",
                3: "VMAP_ERR_CAPACITY", 4: "VMAP_ERR_NO_DEVICE", 5: "VMAP_ERR_BUFFER_TOO_SMALL",
                6: "VMAP_ERR_NOT_CAPTURED", 7: "VMAP_ERR_DIST"}


class VmapError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


def lib_path():
    return os.environ.get("VMAP_LIB") or _build.LIB   # VMAP_LIB: A/B experiments only


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python -m volumetric_mapping_b200.build` (or "
                           "__graft_entry__.build()); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, sz, dp = C.c_void_p, C.c_size_t, C.POINTER(C.c_double)
    L.vmap_version.restype = C.c_char_p
    L.vmap_default_params.argtypes = [C.POINTER(Params)]
    L.vmap_default_params.restype = None
    L.vmap_default_config.argtypes = [C.POINTER(Config)]
    L.vmap_default_config.restype = None
    L.vmap_create.argtypes = [C.POINTER(Params), C.POINTER(Config), C.POINTER(vp)]
    L.vmap_destroy.argtypes = [vp]
    L.vmap_reset.argtypes = [vp]
    L.vmap_set_params.argtypes = [vp, C.POINTER(Params)]
    L.vmap_get_params.argtypes = [vp, C.POINTER(Params)]
    L.vmap_last_error.argtypes = [vp]
    L.vmap_last_error.restype = C.c_char_p
    L.vmap_insert_pointcloud.argtypes = [vp, vp, sz, sz, C.c_int, dp, vp, C.POINTER(sz)]
    L.vmap_insert_pointcloud_dev.argtypes = [vp, vp, sz, sz, dp]
    L.vmap_insert_projected.argtypes = [vp, vp, C.c_int, C.c_int, sz, dp, dp]
    L.vmap_insert_disparity.argtypes = [vp, vp, C.c_int, C.c_int, sz, dp, C.c_double, dp, dp]
    L.vmap_insert_disparity_dev.argtypes = [vp, vp, C.c_int, C.c_int, sz, dp, C.c_double, dp, dp]
    L.vmap_reproject.argtypes = [vp, vp, C.c_int, C.c_int, sz, dp, vp]
    L.vmap_apply_keys.argtypes = [vp, vp, sz, vp, sz]
    L.vmap_query_points.argtypes = [vp, vp, sz, vp]
    L.vmap_query_points_dev.argtypes = [vp, vp, sz, vp]
    L.vmap_query_lines.argtypes = [vp, vp, sz, vp]
    L.vmap_query_lines_dev.argtypes = [vp, vp, sz, vp]
    L.vmap_leaf_count.argtypes = [vp, C.POINTER(sz)]
    L.vmap_export_leaves.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    L.vmap_import_leaves.argtypes = [vp, vp, vp, sz]
    L.vmap_export_leaves_dev.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    L.vmap_import_leaves_dev.argtypes = [vp, vp, vp, sz]
    L.vmap_last_scan_stats.argtypes = [vp, C.POINTER(ScanStats)]
    L.vmap_set_debug_capture.argtypes = [vp, C.c_int]
    L.vmap_debug_last_scan_keys.argtypes = [vp, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz)]
    L.vmap_logodds_constants.argtypes = [vp, C.POINTER(C.c_float)]
    L.vmap_stream.argtypes = [vp]
    L.vmap_stream.restype = vp
    L.vmap_kernel_launches.argtypes = [vp]
    L.vmap_kernel_launches.restype = C.c_uint64
    L.vmap_device_bytes.argtypes = [vp]
    L.vmap_device_bytes.restype = C.c_uint64
    L.vmap_last_walk_kernel_ms.argtypes = [vp]
    L.vmap_last_walk_kernel_ms.restype = C.c_double
    u64p = C.POINTER(C.c_uint64)
    L.vmap_tree_size.argtypes = [vp, C.c_int, C.POINTER(sz)]
    L.vmap_leaf_digest.argtypes = [vp, u64p, u64p]
    L.vmap_flush.argtypes = [vp]
    L.vmap_set_pipeline_depth.argtypes = [vp, C.c_int]
    L.vmap_totals.argtypes = [vp, C.POINTER(ScanStats), u64p, u64p]
    L.vmap_region_begin.argtypes = [vp]
    L.vmap_region_end.argtypes = [vp, dp]
    L.vmap_query_lines_probes_dev.argtypes = [vp, vp, sz, vp, u64p]
    L.vmap_query_bbox_status.argtypes = [vp, vp, sz, dp, vp]
    L.vmap_stage_pointcloud.argtypes = [vp, vp, sz, sz]
    L.vmap_insert_staged.argtypes = [vp, C.c_int, dp]
    L.vmap_load_occupied_pointcloud_file.argtypes = [vp, C.c_char_p, C.POINTER(sz)]
    L.vmap_insert_occupied_points.argtypes = [vp, vp, sz, sz]
    L.vmap_host_read_pointcloud_file.argtypes = [C.c_char_p, vp, sz, C.POINTER(sz)]
    L.vmap_query_bbox_status_dev.argtypes = [vp, vp, sz, dp, vp]
    L.vmap_check_path_collisions.argtypes = [vp, vp, sz, dp, C.POINTER(C.c_int), C.POINTER(sz)]
    L.vmap_host_tree_stream.argtypes = [C.POINTER(Params), vp, vp, sz, C.c_int, vp, sz, C.POINTER(sz), C.POINTER(sz)]
    L.vmap_host_parse_stream.argtypes = [C.POINTER(Params), C.c_int, vp, sz, vp, vp, sz, C.POINTER(sz), dp]
    L.vmap_get_changed.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    L.vmap_set_log_odds_bbox.argtypes = [vp, vp, sz, dp, C.c_double, C.c_int]
    L.vmap_query_points_probability.argtypes = [vp, vp, sz, vp, vp]
    L.vmap_query_visibility.argtypes = [vp, vp, sz, C.c_int, vp]
    L.vmap_query_lines_bbox.argtypes = [vp, vp, sz, dp, vp]
    L.vmap_serialize_bt.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.vmap_write_bt.argtypes = [vp, C.c_char_p, C.c_int]
    L.vmap_load_bt.argtypes = [vp, C.c_char_p]
    L.vmap_deserialize_bt.argtypes = [vp, vp, sz]
    L.vmap_serialize_ot.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
    L.vmap_deserialize_ot.argtypes = [vp, C.c_int, vp, sz]
    L.vmap_set_shard.argtypes = [vp, C.c_int, C.c_int]
    L.vmap_shard_owner.argtypes = [C.c_uint16, C.c_uint16, C.c_uint16, C.c_int]
    L.vmap_shard_owner.restype = C.c_uint32
    L.vmap_shard_record_bytes.argtypes = []
    L.vmap_shard_record_bytes.restype = sz
    L.vmap_shard_generate.argtypes = [vp, vp, sz, sz, C.c_int, dp, u64p]
    L.vmap_shard_generate_dev.argtypes = [vp, vp, sz, sz, dp, u64p]
    L.vmap_shard_records.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
    L.vmap_shard_apply_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), sz]
    L.vmap_shard_query_points_dev.argtypes = [vp, vp, sz, vp]
    L.vmap_shard_query_lines_dev.argtypes = [vp, vp, sz, vp]
    L.vmap_dist_unique_id.argtypes = [vp]
    L.vmap_dist_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.vmap_dist_shutdown.argtypes = [vp]
    L.vmap_dist_flush.argtypes = [vp]
    L.vmap_dist_flush_begin.argtypes = [vp]
    L.vmap_dist_attach_local.argtypes = [C.POINTER(vp), C.c_int, C.c_uint64]
    L.vmap_dist_set_batch.argtypes = [vp, C.c_int]
    L.vmap_set_debug_l2_flush.argtypes = [vp, sz]
    L.vmap_dist_insert_pointcloud.argtypes = [vp, vp, sz, sz, C.c_int, dp]
    L.vmap_dist_insert_pointcloud_dev.argtypes = [vp, vp, sz, sz, dp]
    L.vmap_dist_query_points_dev.argtypes = [vp, vp, sz, vp]
    L.vmap_dist_query_lines_dev.argtypes = [vp, vp, sz, vp]
    _lib = L
    return L


def default_params(**overrides):
    p = Params()
    lib().vmap_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def default_config(**overrides):
    c = Config()
    lib().vmap_default_config(C.byref(c))
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c
