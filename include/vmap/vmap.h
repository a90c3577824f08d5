/*
 * vmap.h -- C ABI of the B200-native scan-insertion path ("libvmap_b200.so").
 *
 * This is the drop-in boundary that sits UNDER volumetric_mapping::OctomapWorld: the reference
 * has no FFI layer of its own, so every entry point below names the C++ member
 * it replaces in /root/reference (paths relative to that tree).  Plain pointers, sizes and PODs
 * only; every function returns a vmap_status (never throws); per-point geometric failures are
 * silent, exactly like the reference (out-of-bounds ray -> nothing inserted).
 *
 * All "host" pointers are ordinary (pageable or pinned) host memory; functions with a `_dev`
 * suffix take CUDA device pointers on the handle's device instead.  A handle is
 * thread-compatible: one mutating call in flight per handle.
 */
#ifndef VMAP_VMAP_H_
#define VMAP_VMAP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VMAP_API
#else
#define VMAP_API __attribute__((visibility("default")))
#endif

typedef enum vmap_status {
  VMAP_OK = 0,
  VMAP_ERR_INVALID_ARGUMENT = 1,
  VMAP_ERR_CUDA = 2,          /* a CUDA runtime call failed; see vmap_last_error()            */
  VMAP_ERR_CAPACITY = 3,      /* device map could not grow (out of device memory / hard cap)  */
  VMAP_ERR_NO_DEVICE = 4,     /* no usable CUDA device: there is NO CPU fallback              */
  VMAP_ERR_BUFFER_TOO_SMALL = 5,
  VMAP_ERR_NOT_CAPTURED = 6,  /* debug key capture was not enabled for the last scan          */
  VMAP_ERR_DIST = 7           /* multi-GPU exchange failure                                   */
} vmap_status;

/* CellStatus of volumetric_map_base/include/volumetric_map_base/world_base.h:54 */
enum { VMAP_CELL_FREE = 0, VMAP_CELL_OCCUPIED = 1, VMAP_CELL_UNKNOWN = 2 };

/* volumetric_mapping::OctomapParameters, octomap_world/include/octomap_world/octomap_world.h:54-109
 * (same fields, order, defaults; bools widened to int32). */
typedef struct vmap_params {
  double resolution;              /* 0.15 */
  double probability_hit;         /* 0.65 */
  double probability_miss;        /* 0.4  */
  double threshold_min;           /* 0.12 */
  double threshold_max;           /* 0.97 */
  double threshold_occupancy;     /* 0.7  */
  int32_t filter_speckles;        /* 1 (unused by the path) */
  int32_t _pad0;
  double max_free_space;          /* 0.0 = no free-space filter (octomap_world.cc:189,215) */
  double min_height_free_space;   /* 0.0 */
  double sensor_max_range;        /* 5.0; negative disables (octomap_world.cc:184) */
  double visualize_min_z;
  double visualize_max_z;
  int32_t treat_unknown_as_occupied; /* 1 */
  int32_t change_detection_enabled;  /* 0 (unused by the path) */
} vmap_params;

/* How per-scan key-set construction (octomap::KeySet free_cells / occupied_cells,
 * octomap_world.cc:126,150) is realised on the device. */
enum {
  VMAP_DEDUPE_BITMASK = 0, /* per-8^3-block 512-bit free/occupied masks OR-ed in L2 (default) */
  VMAP_DEDUPE_SORT = 1     /* emit packed keys, radix sort + unique, occupied-wins            */
};

/* Device-side sizing; everything grows on demand, these are starting points. 0 = default. */
typedef struct vmap_config {
  int32_t device;             /* CUDA device ordinal; -1 = current device                    */
  int32_t dedupe_mode;        /* VMAP_DEDUPE_*                                               */
  uint64_t initial_blocks;    /* voxel-block pool capacity (8^3 voxels = 2 KiB each)         */
  uint64_t max_blocks;        /* hard cap for growth; 0 = limited by device memory only      */
  uint64_t initial_scan_points;
  uint64_t initial_scan_keys; /* sort mode: packed-key buffer capacity                       */
  uint64_t dense_window_bytes;/* budget of the per-scan dense mark window (bitmask mode): a cube
                                 of 8^3 blocks around the sensor, 128 B per block; 0 = 6 GiB,
                                 1 = never use the window (hash-table mask columns only)      */
} vmap_config;

/* Counters of the most recent insert call (the metric's numerators). */
typedef struct vmap_scan_stats {
  uint64_t points_in;        /* points / pixels handed in                                    */
  uint64_t points_valid;     /* finite (cloud) or valid-disparity (image) points             */
  uint64_t points_in_range;  /* castRay's range predicate true (octomap_world.cc:184-185)    */
  uint64_t rays_cast;        /* points not skipped by the occupied_cells test (:133,:168)    */
  uint64_t traversals;       /* keys produced by computeRayKeys over all cast rays           */
  uint64_t occupied_emitted; /* occupied_cells->insert calls (:206)                          */
  uint64_t unique_free;      /* |free_cells| after the erase loop (:252-254)                 */
  uint64_t unique_occupied;  /* |occupied_cells|                                             */
  uint64_t longest_ray;      /* max keys of one ray                                          */
  uint64_t blocks_touched;   /* voxel blocks updated by this scan                            */
  uint64_t blocks_total;     /* voxel blocks resident in the map after this scan             */
  uint64_t retries;          /* times the scan was re-run after a device-side grow           */
  double gpu_ms;             /* device time of the scan's kernels (CUDA events)              */
  /* the same, split by stage (CUDA events on the handle's stream between the kernels):        */
  double front_ms;           /* K1  transform / projection / classification                  */
  double resolve_ms;         /* K1b skip rule, cast-ray list, occupied marks                 */
  double walk_ms;            /* K2  DDA walk + key-set construction                          */
  double update_ms;          /* K4  log-odds update of the touched voxel blocks              */
  /* sharded (multi-GPU) rounds only:                                                          */
  uint64_t records_sent;     /* block records sent to other ranks                            */
  uint64_t records_received; /* block records received from other ranks                      */
  uint64_t blocks_applied;   /* block records applied to this shard in the round             */
} vmap_scan_stats;

typedef struct vmap vmap;

VMAP_API const char* vmap_version(void);
/* Fills OctomapParameters' constructor defaults (octomap_world.h:56-69). */
VMAP_API void vmap_default_params(vmap_params* p);
VMAP_API void vmap_default_config(vmap_config* c);

/* OctomapWorld::OctomapWorld(const OctomapParameters&)  (octomap_world.cc:53-56). */
VMAP_API int vmap_create(const vmap_params* params, const vmap_config* config, vmap** out);
VMAP_API int vmap_destroy(vmap* h);
/* OctomapWorld::resetMap  (octomap_world.cc:75-80). */
VMAP_API int vmap_reset(vmap* h);
/* OctomapWorld::setOctomapParameters (octomap_world.cc:84-104): a changed resolution resets the
 * map, everything else is applied to the existing map. */
VMAP_API int vmap_set_params(vmap* h, const vmap_params* params);
/* OctomapWorld::getOctomapParameters (octomap_world.cc:106-108). */
VMAP_API int vmap_get_params(const vmap* h, vmap_params* params);
/* Message of the last failing call on this handle (h == NULL: last vmap_create failure). */
VMAP_API const char* vmap_last_error(const vmap* h);

/* OctomapWorld::insertPointcloudIntoMapImpl (octomap_world.cc:110-141), reached from
 * WorldBase::insertPointcloud (world_base.cc:182-213).
 *   xyz, n, stride_bytes : n points, 3 floats each, stride_bytes apart (16 for pcl::PointXYZ)
 *   is_dense             : pcl cloud.is_dense; when 0, non-finite points are dropped first
 *   T_rowmajor           : Transformation::getTransformationMatrix(), row-major 4x4 double
 *   xyz_world_out        : optional (may be NULL).  The reference mutates the caller's cloud in
 *                          place (NaN compaction + world-frame transform, :115-119); when
 *                          non-NULL the device result is written back here with the same
 *                          stride (may alias xyz) and *n_out receives the new point count. */
VMAP_API int vmap_insert_pointcloud(vmap* h, const float* xyz, size_t n, size_t stride_bytes,
                                    int is_dense, const double T_rowmajor[16],
                                    float* xyz_world_out, size_t* n_out);
/* Same, points already resident on the handle's device (tightly packed when stride is 12). */
VMAP_API int vmap_insert_pointcloud_dev(vmap* h, const float* d_xyz, size_t n,
                                        size_t stride_bytes, const double T_rowmajor[16]);

/* ---- the scan pipeline ------------------------------------------------------------------------
 * vmap_insert_pointcloud[_dev] is PIPELINED: the call stages the scan on the device, enqueues all
 * of its kernels and returns as soon as the caller's buffer has been consumed (and, when
 * xyz_world_out is given, rewritten); up to `depth` scans stay in flight, overlapping on the
 * device, while their updates of the map are applied strictly in insertion order.  Every other
 * entry point that reads or edits the map drains the pipeline first, so the reference's
 * observable behaviour (octomap_world.cc:110-141 is synchronous) is unchanged; what changes is
 * WHEN a failure is reported: a scan that cannot be applied (VMAP_ERR_CAPACITY) is dropped without
 * leaving marks behind and the error is returned by the call that retires it -- a later insert,
 * vmap_flush, or the next reading call.  depth 0 restores fully synchronous inserts.
 * Key capture (vmap_set_debug_capture), VMAP_DEDUPE_SORT, unlimited sensor range, windows over
 * vmap_config.dense_window_bytes and sharded handles always insert synchronously. */
VMAP_API int vmap_flush(vmap* h);
VMAP_API int vmap_set_pipeline_depth(vmap* h, int depth /* 0, 1 or 2 (default) */);
/* Sums of the per-scan counters over every scan inserted since creation / vmap_reset (drains the
 * pipeline first); *n_scans scans, *n_replayed of them had their update refused on the device for
 * lack of room and were replayed after the map had grown (both may be NULL). */
VMAP_API int vmap_totals(vmap* h, vmap_scan_stats* sums, uint64_t* n_scans, uint64_t* n_replayed);
/* Device-timed region around any number of calls: begin drains everything on the handle and
 * records a CUDA event, end drains again, records a second event and returns the milliseconds
 * between the two. */
VMAP_API int vmap_region_begin(vmap* h);
VMAP_API int vmap_region_end(vmap* h, double* ms);

/* OctomapWorld::insertProjectedDisparityIntoMapImpl (octomap_world.cc:143-175): CV_32FC3 image
 * of sensor-frame points, rows x cols, row_pitch_bytes between rows; sensor_to_world given as
 * unit quaternion (w,x,y,z) + translation, the form the reference evaluates (:162). */
VMAP_API int vmap_insert_projected(vmap* h, const float* xyz_hw3, int rows, int cols,
                                   size_t row_pitch_bytes, const double q_wxyz[4],
                                   const double t[3]);
/* WorldBase::insertDisparityImage(cv::Mat) (world_base.cc:51-87) fused with the above:
 * Q down-sampling fix-up (:55-69), cv::reprojectImageTo3D(..., handleMissingValues=true) (:75),
 * then insertion.  disparity is CV_32FC1; Q_full row-major 4x4; full_image_width =
 * full_image_size.x(). */
VMAP_API int vmap_insert_disparity(vmap* h, const float* disparity, int rows, int cols,
                                   size_t row_pitch_bytes, const double Q_full_rowmajor[16],
                                   double full_image_width, const double q_wxyz[4],
                                   const double t[3]);
VMAP_API int vmap_insert_disparity_dev(vmap* h, const float* d_disparity, int rows, int cols,
                                       size_t row_pitch_bytes, const double Q_full_rowmajor[16],
                                       double full_image_width, const double q_wxyz[4],
                                       const double t[3]);
/* Projection only (cv::reprojectImageTo3D semantics), for parity tests of kernel 1. */
VMAP_API int vmap_reproject(vmap* h, const float* disparity, int rows, int cols,
                            size_t row_pitch_bytes, const double Q_rowmajor[16],
                            float* xyz_hw3_out);

/* OctomapWorld::updateOccupancy(KeySet* free_cells, KeySet* occupied_cells)
 * (octomap_world.cc:238-264); also called by OctomapManager::loadOctomapCallback
 * (octomap_manager.cc:332-341).  Keys are 3 x uint16 (octomap::OcTreeKey). */
VMAP_API int vmap_apply_keys(vmap* h, const uint16_t* free_k3, size_t n_free,
                             const uint16_t* occ_k3, size_t n_occ);

/* OctomapWorld::getCellStatusPoint (octomap_world.cc:346-360), batched: n points (3 doubles
 * each) -> n status bytes (VMAP_CELL_*). */
VMAP_API int vmap_query_points(vmap* h, const double* xyz, size_t n, uint8_t* status);
VMAP_API int vmap_query_points_dev(vmap* h, const double* d_xyz, size_t n, uint8_t* d_status);
/* OctomapWorld::getLineStatus (octomap_world.cc:395-418), batched: n segments (start xyz, end
 * xyz = 6 doubles each) -> n status bytes. */
VMAP_API int vmap_query_lines(vmap* h, const double* start_end_xyz6, size_t n, uint8_t* status);
VMAP_API int vmap_query_lines_dev(vmap* h, const double* d_start_end_xyz6, size_t n,
                                  uint8_t* d_status);

/* Measurement twin of vmap_query_lines_dev: also returns how many voxels the reference's loop
 * (octomap_world.cc:405-415: one octree_->search per key until the first non-free one) probes in
 * total -- the `probes` of the query's algorithmic-byte figure. */
VMAP_API int vmap_query_lines_probes_dev(vmap* h, const double* d_start_end_xyz6, size_t n,
                                         uint8_t* d_status, uint64_t* probes);

/* Query variants that are fan-outs of the two kernels above, batched:
 * getCellTrueStatusPoint + getCellProbabilityPoint (octomap_world.cc:363-393): status with unknown
 * kept as VMAP_CELL_UNKNOWN; probability (may be NULL) = node->getOccupancy(), -1.0 without node. */
VMAP_API int vmap_query_points_probability(vmap* h, const double* xyz, size_t n, uint8_t* status,
                                           double* probability);
/* getVisibility (octomap_world.cc:420-449): n x (view point xyz, voxel to test xyz). */
VMAP_API int vmap_query_visibility(vmap* h, const double* view_voxel_xyz6, size_t n,
                                   int stop_at_unknown_cell, uint8_t* status);
/* getLineStatusBoundingBox (octomap_world.cc:451-493): every segment swept with the same box. */
VMAP_API int vmap_query_lines_bbox(vmap* h, const double* start_end_xyz6, size_t n,
                                   const double bounding_box_size[3], uint8_t* status);

/* getCellStatusBoundingBox (octomap_world.cc:274-344): status of a box of `bounding_box_size`
 * centred on each point -- the centre's own status when it is not free, else occupied if any
 * occupied leaf lies in the box, else unknown (occupied when treat_unknown_as_occupied) if
 * octomap's getUnknownLeafCenters finds a hole, else free.  Single-device maps only. */
VMAP_API int vmap_query_bbox_status(vmap* h, const double* xyz, size_t n,
                                    const double bounding_box_size[3], uint8_t* status);
VMAP_API int vmap_query_bbox_status_dev(vmap* h, const double* d_xyz, size_t n,
                                        const double bounding_box_size[3], uint8_t* d_status);
/* checkPathForCollisionsWithRobot (octomap_world.cc:1384-1400; checkSinglePoseCollision
 * :1402-1412): *collides = 1 and *collision_index (may be NULL) = earliest pose whose robot box
 * is not free (treat_unknown_as_occupied) / is occupied (otherwise). */
VMAP_API int vmap_check_path_collisions(vmap* h, const double* xyz, size_t n,
                                        const double robot_size[3], int* collides,
                                        size_t* collision_index);

/* Finest-level leaves (what octree_->begin_leafs() would visit after expanding pruned nodes):
 * used by the on-demand export to octomap::OcTree for the unchanged publish/save paths
 * (octomap_world.cc:820-856) and by the parity tests. */
VMAP_API int vmap_leaf_count(vmap* h, size_t* n);
VMAP_API int vmap_export_leaves(vmap* h, uint16_t* keys_k3, float* logodds, size_t capacity,
                                size_t* n);
/* Order-independent 64-bit digest of (key, log-odds bit pattern) over all leaves, and their count:
 * two maps (or a sharded map's shards summed together, the sum wraps modulo 2^64) hold bit-identical
 * leaves iff both numbers agree.  volumetric_mapping_b200.sharded.leaf_digest is the host mirror. */
VMAP_API int vmap_leaf_digest(vmap* h, uint64_t* n_leaves, uint64_t* digest);
/* Inverse (loadOctomapFromFile / setOctomapFromMsg feed): sets leaf values, no clamping. */
VMAP_API int vmap_import_leaves(vmap* h, const uint16_t* keys_k3, const float* logodds, size_t n);
/* Both with DEVICE buffers (memory of the handle's GPU), no host round trip: how a read-only
 * replica of a sharded map is assembled on every GPU for query batches (SURVEY 8(e) "gather then
 * query on every GPU over a segment slice"; the reference's queries, octomap_world.cc:346-360 and
 * :395-418, see one whole tree). */
VMAP_API int vmap_export_leaves_dev(vmap* h, uint16_t* d_keys_k3, float* d_logodds, size_t capacity,
                                    size_t* n);
VMAP_API int vmap_import_leaves_dev(vmap* h, const uint16_t* d_keys_k3, const float* d_logodds,
                                    size_t n);

/* On-demand export to octomap's tree form, built on the host from the finest-level leaves:
 * canonical pruned tree (8 equal leaf children collapse; inner value = max of children). */
/* octree_->prune(); octree_->size().  max_likelihood != 0: after toMaxLikelihood(). */
VMAP_API int vmap_tree_size(vmap* h, int max_likelihood, size_t* n_nodes);
/* OctomapWorld::writeOctomapToBinaryConst (octomap_world.cc:854-856): ".bt" stream into memory
 * (buffer == NULL: size query). */
VMAP_API int vmap_serialize_bt(vmap* h, uint8_t* buffer, size_t capacity, size_t* n_bytes);
/* OctomapWorld::writeOctomapToFile (octomap_world.cc:849-852).  OcTree::writeBinary(filename)
 * thresholds the LIVE tree to the clamping values before writing; pass threshold_live_map = 1
 * for that side effect, 0 for the const variant. */
VMAP_API int vmap_write_bt(vmap* h, const char* path, int threshold_live_map);
/* OctomapWorld::loadOctomapFromFile (octomap_world.cc:846-848) / setOctomapFromMsg (binary)
 * (:833-844): replaces the map (and its resolution) with the stream's. */
VMAP_API int vmap_load_bt(vmap* h, const char* path);
/* OctomapWorld::getOctomapFullMsg (octomap_world.cc:826-831) / OcTree::write: the full tree, one
 * float log-odds per node (inner node = max of its children).  with_header = 0: the message
 * payload (writeData); 1: a complete ".ot" stream.  vmap_deserialize_ot is the inverse
 * (setOctomapFromMsg with a full message, :833-844). */
VMAP_API int vmap_serialize_ot(vmap* h, int with_header, uint8_t* buffer, size_t capacity, size_t* n_bytes);
VMAP_API int vmap_deserialize_ot(vmap* h, int with_header, const uint8_t* buffer, size_t n_bytes);
VMAP_API int vmap_deserialize_bt(vmap* h, const uint8_t* buffer, size_t n_bytes);

/* The same tree fold and stream writers / readers on caller-supplied leaves, WITHOUT a device
 * (pure host functions; kind 0 = ".bt", 1 = full-tree data, 2 = ".ot" with header). */
VMAP_API int vmap_host_tree_stream(const vmap_params* params, const uint16_t* keys_k3,
                                   const float* logodds, size_t n, int kind, uint8_t* buffer,
                                   size_t capacity, size_t* n_bytes, size_t* n_nodes);
VMAP_API int vmap_host_parse_stream(const vmap_params* params, int kind, const uint8_t* buffer,
                                    size_t n_bytes, uint16_t* keys_k3, float* logodds,
                                    size_t capacity, size_t* n_leaves, double* resolution);

/* Pipelined vmap_insert_pointcloud for scans streamed from host memory: vmap_stage_pointcloud
 * starts an asynchronous host->device copy of a scan (the host buffer, ideally pinned, must stay
 * untouched until the matching vmap_insert_staged returns); vmap_insert_staged inserts the oldest
 * staged scan exactly like vmap_insert_pointcloud_dev.  Staging scan k+1 before inserting scan k
 * overlaps its copy with scan k's kernels.  At most two scans can be staged. */
VMAP_API int vmap_stage_pointcloud(vmap* h, const float* xyz, size_t n, size_t stride_bytes);
VMAP_API int vmap_insert_staged(vmap* h, int is_dense, const double T_rowmajor[16]);

/* OctomapManager::loadOctomapCallback's point-cloud branch (octomap_manager.cc:313-344): every
 * point of a .pcd / .ply file whose key lies inside the map becomes an occupied key; the free set
 * stays empty; updateOccupancy applies them.  vmap_insert_occupied_points is the same for points
 * already in memory; vmap_host_read_pointcloud_file only parses the file (no device needed). */
VMAP_API int vmap_load_occupied_pointcloud_file(vmap* h, const char* path, size_t* n_points);
VMAP_API int vmap_insert_occupied_points(vmap* h, const float* xyz, size_t n, size_t stride_bytes);
VMAP_API int vmap_host_read_pointcloud_file(const char* path, float* xyz, size_t capacity, size_t* n);

/* OctomapWorld::setLogOddsBoundingBox (octomap_world.cc:781-818) behind setFree / setOccupied
 * (:497-523): n boxes of one size; bound_handling = BoundHandling (octomap_world.h:43-52:
 * 0 kDefault, 1 kIncludePartialBoxes, 2 kIgnorePartialBoxes).  setFree passes the clamping
 * minimum, setOccupied the clamping maximum (vmap_logodds_constants). */
VMAP_API int vmap_set_log_odds_bbox(vmap* h, const double* positions_xyz, size_t n_positions,
                                    const double bounding_box_size[3], double log_odds_value,
                                    int bound_handling);

/* OctomapWorld::getChangedPoints (octomap_world.cc:1413-1440): keys collected by octomap's change
 * detection since the last call (created leaves; leaves whose occupancy flipped) with the current
 * occupancy of each, followed by resetChangeDetection().  Requires change_detection_enabled
 * (vmap_params; single-GPU insertion paths). */
VMAP_API int vmap_get_changed(vmap* h, uint16_t* keys_k3, uint8_t* occupied, size_t capacity, size_t* n);

/* Per-scan counters of the last insert / apply call. */
VMAP_API int vmap_last_scan_stats(const vmap* h, vmap_scan_stats* stats);
/* Parity hook: when enabled, the next inserts keep their final free / occupied key sets. */
VMAP_API int vmap_set_debug_capture(vmap* h, int enabled);
VMAP_API int vmap_debug_last_scan_keys(vmap* h, uint16_t* free_k3, size_t free_capacity,
                                       size_t* n_free, uint16_t* occ_k3, size_t occ_capacity,
                                       size_t* n_occ);
/* (float)log(p/(1-p)) constants in effect: hit, miss, clamp min, clamp max, occupancy. */
VMAP_API int vmap_logodds_constants(const vmap* h, float out5[5]);
/* ---------------------------------------------------------------------------------------------
 * Multi-GPU sharding (one handle = one map shard = one GPU = one process).
 *
 * Every 8x8x8 voxel block has one owner rank.  A scan is GENERATED on one rank (K1..K2, the
 * key sets of octomap_world.cc:126-137 as per-block masks), the touched blocks are packed into
 * one record stream per owner, the streams are exchanged, and every owner APPLIES the scans'
 * records in scan order (updateOccupancy, octomap_world.cc:238-264, restricted to its blocks).
 * vmap_dist_* does the exchange itself over NCCL (all-to-all of records over NVLink); the
 * vmap_shard_* pair exposes the two halves so a host can use any transport.
 * --------------------------------------------------------------------------------------------- */
VMAP_API int vmap_set_shard(vmap* h, int rank, int world);
/* Owner rank of the block containing key (kx,ky,kz); pure host function. */
VMAP_API uint32_t vmap_shard_owner(uint16_t kx, uint16_t ky, uint16_t kz, int world);
VMAP_API size_t vmap_shard_record_bytes(void);
/* Generate one scan on this rank; counts_out[world] = records destined to each rank (may be
 * NULL).  Nothing is applied to the map. */
VMAP_API int vmap_shard_generate(vmap* h, const float* xyz, size_t n, size_t stride_bytes,
                                 int is_dense, const double T_rowmajor[16], uint64_t* counts_out);
VMAP_API int vmap_shard_generate_dev(vmap* h, const float* d_xyz, size_t n, size_t stride_bytes,
                                     const double T_rowmajor[16], uint64_t* counts_out);
/* Device pointer / count of the records of the last generated scan destined to rank `dest`. */
VMAP_API int vmap_shard_records(vmap* h, int dest, const void** d_records, size_t* n_records);
/* Apply record lists (device memory) to this shard: list i is the i-th scan of the round. */
VMAP_API int vmap_shard_apply_dev(vmap* h, const void* const* d_record_lists,
                                  const size_t* n_records, size_t n_lists);
/* Partial query answers of this shard; the global answer is the element-wise MINIMUM over ranks
 * (points: status byte, 0xFF = not mine; lines: first-hit index << 2 | status, 0xFFFFFFFF = no
 * non-free voxel of mine on the segment -> free when that is the minimum). */
VMAP_API int vmap_shard_query_points_dev(vmap* h, const double* d_xyz, size_t n, uint8_t* d_status);
VMAP_API int vmap_shard_query_lines_dev(vmap* h, const double* d_start_end_xyz6, size_t n,
                                        uint32_t* d_packed);

/* NCCL transport.  Rank 0 obtains an id (ncclUniqueId, 128 bytes) and hands it to every rank by
 * any means (torch.distributed broadcast in bench.py); all ranks then call vmap_dist_init. */
VMAP_API int vmap_dist_unique_id(uint8_t out128[128]);
VMAP_API int vmap_dist_init(vmap* h, const uint8_t id128[128], int rank, int world);
VMAP_API int vmap_dist_shutdown(vmap* h);
/* COLLECTIVE: every rank passes its own scan; rank r's scan is inserted as the r-th scan of the
 * round on every shard (n may be 0 on ranks that have no scan left). */
VMAP_API int vmap_dist_insert_pointcloud(vmap* h, const float* xyz, size_t n, size_t stride_bytes,
                                         int is_dense, const double T_rowmajor[16]);
VMAP_API int vmap_dist_insert_pointcloud_dev(vmap* h, const float* d_xyz, size_t n,
                                             size_t stride_bytes, const double T_rowmajor[16]);
/* The collective inserts return as soon as the round's exchange and apply are ENQUEUED: they run
 * on a second stream while the next round is generated.  Every call that reads the map waits for
 * the round in flight first; vmap_dist_flush does so explicitly. */
VMAP_API int vmap_dist_flush(vmap* h);
/* TRANSPORT.  By default vmap_dist_init sets up the PEER-MEMORY transport: every rank's block
 * records stay in a round buffer in its own device memory, which all ranks of the box map through
 * CUDA IPC; a round is published with one remote store per peer, and each owner applies the
 * records straight out of its peers' buffers (the all-to-all is the apply kernel's own NVLink
 * loads).  Inserts are pipelined like vmap_insert_pointcloud; vmap_dist_flush completes them.
 * VMAP_DIST_TRANSPORT=nccl (environment), or a rank that cannot map a peer, selects the NCCL
 * send/recv transport for all ranks instead.  VMAP_DIST_ROUND_RECORDS sizes the round buffers
 * (default 3 Mi records of 136 bytes, twice per rank).
 *
 * vmap_dist_attach_local: `world` handles of ONE process on one device stand for the ranks (tests,
 * single-GPU development); mailboxes are shared by plain pointers.  Drive every handle from its own
 * host thread, as if it were a process (a rank's host may block until its peers have made progress).
 * vmap_dist_flush_begin is the non-blocking first half of vmap_dist_flush (seal + publish a partial
 * batch). */
VMAP_API int vmap_dist_attach_local(vmap** handles, int world, uint64_t records_per_round);
VMAP_API int vmap_dist_flush_begin(vmap* h);
/* Scans per rank that share one exchange round (1..8, default 1).  COLLECTIVE setting. */
VMAP_API int vmap_dist_set_batch(vmap* h, int scans_per_rank);
/* Benchmark hook: overwrite `bytes` (> L2) of scratch before every round's exchange. 0 = off. */
VMAP_API int vmap_set_debug_l2_flush(vmap* h, size_t bytes);
/* COLLECTIVE queries: every rank passes the same queries and receives the global statuses. */
VMAP_API int vmap_dist_query_points_dev(vmap* h, const double* d_xyz, size_t n, uint8_t* d_status);
VMAP_API int vmap_dist_query_lines_dev(vmap* h, const double* d_start_end_xyz6, size_t n,
                                       uint8_t* d_status);

/* cudaStream_t all of the handle's work is enqueued on (for CUDA-event timing by callers). */
VMAP_API void* vmap_stream(vmap* h);
/* Number of kernel launches issued by this handle since creation (bench.py gpu_launches). */
VMAP_API uint64_t vmap_kernel_launches(const vmap* h);
/* Device memory currently held by the handle, bytes. */
VMAP_API uint64_t vmap_device_bytes(const vmap* h);
/* Duration of the segment walker's launch (K2b, computeRayKeys + KeySet inserts: octomap_world.cc:188-190, 214-216)
 * in the last SYNCHRONOUS scan (pipeline depth 0), CUDA events directly around that one launch; 0 when the scan
 * recorded no stage marks.  bench.py's roofline divides the walk's algorithmic bytes by this. */
VMAP_API double vmap_last_walk_kernel_ms(const vmap* h);

#ifdef __cplusplus
}
#endif
#endif /* VMAP_VMAP_H_ */
