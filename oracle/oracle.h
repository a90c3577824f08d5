/*
 * oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the timed CPU baseline.  See oracle.cc for the provenance of
 * every function (reference file:line, or the un-vendored upstream library it restates).
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors and cannot be
 * compiled in this environment (octomap / PCL / OpenCV-C++ / Eigen / minkindr absent), so the
 * oracle is pinned only by (a) python cv2 4.13 `reprojectImageTo3D` (executable here, golden
 * vectors under tests/golden/), and (b) hand-derived known-answer tests (tests/test_oracle_*.py).
 */
#ifndef VMAP_ORACLE_H_
#define VMAP_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same fields, order and defaults as volumetric_mapping::OctomapParameters
 * (octomap_world/include/octomap_world/octomap_world.h:54-109); bools widened to int32. */
typedef struct oracle_params {
  double resolution;
  double probability_hit;
  double probability_miss;
  double threshold_min;
  double threshold_max;
  double threshold_occupancy;
  int32_t filter_speckles;
  int32_t _pad0;
  double max_free_space;
  double min_height_free_space;
  double sensor_max_range;
  double visualize_min_z;
  double visualize_max_z;
  int32_t treat_unknown_as_occupied;
  int32_t change_detection_enabled;
} oracle_params;

typedef struct oracle_scan_stats {
  uint64_t points_in;      /* points handed to the insert call                      */
  uint64_t points_valid;   /* after NaN filter / disparity validity                 */
  uint64_t points_in_range;/* castRay took the in-range branch                      */
  uint64_t rays_cast;      /* castRay invocations (not skipped by the occupied test)*/
  uint64_t traversals;     /* keys produced by computeRayKeys over all cast rays    */
  uint64_t occupied_emitted;/* occupied_cells->insert() calls                       */
  uint64_t unique_free;    /* |free_cells| after the occupied erase                 */
  uint64_t unique_occupied;/* |occupied_cells|                                      */
  uint64_t longest_ray;    /* max keys of a single ray                              */
  double t_keyset_s;       /* seconds: NaN filter + transform + ray casting         */
  double t_update_s;       /* seconds: the two updateNode loops                     */
  double t_inner_s;        /* seconds: updateInnerOccupancy()                       */
} oracle_scan_stats;

typedef struct oracle_world oracle_world;

void oracle_default_params(oracle_params* p);
oracle_world* oracle_create(const oracle_params* p);
void oracle_destroy(oracle_world* w);
void oracle_reset(oracle_world* w);
void oracle_set_params(oracle_world* w, const oracle_params* p);
void oracle_get_params(const oracle_world* w, oracle_params* p);
/* When 0, updateInnerOccupancy() after each scan is skipped (leaf values are unaffected);
 * default 1 = faithful to octomap_world.cc:263. */
void oracle_set_faithful_inner_update(oracle_world* w, int on);

/* insertPointcloudIntoMapImpl.  xyz is mutated in place exactly like the reference mutates
 * the caller's cloud (NaN compaction when !is_dense, then world-frame transform);
 * *n_out receives the new point count.  T is the row-major 4x4 of
 * Transformation::getTransformationMatrix(). */
void oracle_insert_pointcloud(oracle_world* w, float* xyz, size_t n, size_t stride_bytes,
                              int is_dense, const double T_rowmajor[16], size_t* n_out);
/* insertProjectedDisparityIntoMapImpl on a CV_32FC3 image. */
void oracle_insert_projected(oracle_world* w, const float* xyz, int rows, int cols,
                             size_t row_pitch_bytes, const double q_wxyz[4], const double t[3]);
/* WorldBase::insertDisparityImage(cv::Mat) : Q fix-up, reprojection, insertion. */
void oracle_insert_disparity(oracle_world* w, const float* disp, int rows, int cols,
                             size_t row_pitch_bytes, const double Q_rowmajor[16],
                             double full_image_width, const double q_wxyz[4], const double t[3]);
/* cv::reprojectImageTo3D(disp, out CV_32FC3, Q, handleMissingValues=true), cv2 4.13 semantics. */
void oracle_reproject(const float* disp, int rows, int cols, size_t row_pitch_bytes,
                      const double Q_rowmajor[16], float* out_xyz);
/* WorldBase::insertDisparityImage's Q down-sampling fix-up (world_base.cc:55-69). */
void oracle_fixup_q(const double Q_full[16], double full_image_width, int disparity_cols,
                    double Q_out[16]);
/* updateOccupancy(free, occupied) on explicit key lists (3 x uint16 each). */
void oracle_apply_keys(oracle_world* w, const uint16_t* free_k3, size_t n_free,
                       const uint16_t* occ_k3, size_t n_occ);

/* getCellStatusPoint / getLineStatus; status codes: 0 free, 1 occupied, 2 unknown. */
void oracle_query_points(const oracle_world* w, const double* xyz, size_t n, uint8_t* status);
void oracle_query_lines(const oracle_world* w, const double* a_b_xyz6, size_t n, uint8_t* status);

/* getCellTrueStatusPoint + getCellProbabilityPoint (octomap_world.cc:363-393), getVisibility
 * (:420-449), getLineStatusBoundingBox (:451-493). */
void oracle_query_points_probability(const oracle_world* w, const double* xyz, size_t n,
                                     uint8_t* status, double* probability);
void oracle_query_visibility(const oracle_world* w, const double* view_voxel_xyz6, size_t n,
                             int stop_at_unknown_cell, uint8_t* status);
void oracle_query_lines_bbox(const oracle_world* w, const double* a_b_xyz6, size_t n,
                             const double bounding_box_size[3], uint8_t* status);

/* getCellStatusBoundingBox (octomap_world.cc:274-344; octomap's leaf_bbx_iterator and
 * getUnknownLeafCenters restated) and checkPathForCollisionsWithRobot (:1384-1412). */
void oracle_query_bbox_status(const oracle_world* w, const double* xyz, size_t n,
                              const double bounding_box_size[3], uint8_t* status);
int oracle_check_path_collisions(const oracle_world* w, const double* xyz, size_t n,
                                 const double robot_size[3], size_t* collision_index);

/* getChangedPoints (octomap_world.cc:1413-1440): returns the count; fills and resets when the
 * buffers are large enough. */
size_t oracle_get_changed(oracle_world* w, uint16_t* k3, uint8_t* occupied, size_t cap);

/* setLogOddsBoundingBox (octomap_world.cc:781-818; setFree / setOccupied :497-523). */
void oracle_set_log_odds_bbox(oracle_world* w, const double* positions_xyz, size_t n,
                              const double bounding_box_size[3], double log_odds_value,
                              int insertion_method);

/* Standalone primitives (known-answer tests). */
int oracle_coord_to_key_checked(const oracle_world* w, double c, uint16_t* key);
uint16_t oracle_coord_to_key(const oracle_world* w, double c);
double oracle_key_to_coord(const oracle_world* w, uint16_t k);
/* computeRayKeys; returns -1 when the function returns false, else the number of keys
 * (written to out_k3 up to cap). */
long oracle_compute_ray_keys(const oracle_world* w, const float origin[3], const float end[3],
                             uint16_t* out_k3, size_t cap);
void oracle_quat_to_matrix(const double q_wxyz[4], const double t[3], double T_rowmajor[16]);
void oracle_logodds_constants(const oracle_world* w, float out5[5]); /* hit miss cmin cmax occ */

/* Key sets of the most recent insert (after updateOccupancy's erase): 3 x uint16 per key. */
size_t oracle_last_free_count(const oracle_world* w);
size_t oracle_last_occupied_count(const oracle_world* w);
void oracle_last_free_keys(const oracle_world* w, uint16_t* out_k3);
void oracle_last_occupied_keys(const oracle_world* w, uint16_t* out_k3);
void oracle_last_scan_stats(const oracle_world* w, oracle_scan_stats* s);

/* Finest-level leaves (pruned nodes expanded): count, then keys + float log-odds. */
size_t oracle_leaf_count(const oracle_world* w);
void oracle_export_leaves(const oracle_world* w, uint16_t* out_k3, float* out_logodds);
size_t oracle_tree_node_count(const oracle_world* w);
/* octree_->prune() (octomap_world.cc:82). */
void oracle_prune(oracle_world* w);
/* writeOctomapToFile (live != 0; thresholds and prunes the live tree like OcTree::writeBinary)
 * or writeOctomapToBinaryConst (live == 0): ".bt" stream; returns its size (buf may be NULL). */
size_t oracle_write_bt(oracle_world* w, int live, unsigned char* buf, size_t cap);
/* OcTree::write / writeData (getOctomapFullMsg payload when with_header == 0). */
size_t oracle_write_ot(const oracle_world* w, int with_header, unsigned char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
