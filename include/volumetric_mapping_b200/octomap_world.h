// octomap_world.h -- host-side mirror of volumetric_mapping::OctomapWorld for the scan-insertion
// path, implemented on the C ABI of libvmap_b200.so (include/vmap/vmap.h).  Reference:
// octomap_world/include/octomap_world/octomap_world.h:54-328 and
// octomap_world/src/octomap_world.cc:50-264,346-360,395-418.  Header-only; link with
// -lvmap_b200.  There is no CPU fallback: construction throws when no CUDA device is usable.
#ifndef VOLUMETRIC_MAPPING_B200_OCTOMAP_WORLD_H_
#define VOLUMETRIC_MAPPING_B200_OCTOMAP_WORLD_H_

#include <cstdio>
#include <cstring>
#include <ostream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "vmap/vmap.h"
#include "volumetric_mapping_b200/world_base.h"

namespace volumetric_mapping {

using compat::KeySet;
using compat::OcTreeKey;
using compat::OctomapMsg;

// octomap_world.h:43-52
enum BoundHandling { kDefault, kIncludePartialBoxes, kIgnorePartialBoxes };

// octomap_world.h:54-109 -- identical fields and defaults.
struct OctomapParameters {
  OctomapParameters()
      : resolution(0.15), probability_hit(0.65), probability_miss(0.4), threshold_min(0.12),
        threshold_max(0.97), threshold_occupancy(0.7), filter_speckles(true), max_free_space(0.0),
        min_height_free_space(0.0), sensor_max_range(5.0),
        visualize_min_z(-std::numeric_limits<double>::max()),
        visualize_max_z(std::numeric_limits<double>::max()), treat_unknown_as_occupied(true),
        change_detection_enabled(false) {}
  double resolution, probability_hit, probability_miss, threshold_min, threshold_max, threshold_occupancy;
  bool filter_speckles;
  double max_free_space, min_height_free_space, sensor_max_range, visualize_min_z, visualize_max_z;
  bool treat_unknown_as_occupied;
  bool change_detection_enabled;
};

class OctomapWorld : public WorldBase {
 public:
  typedef std::shared_ptr<OctomapWorld> Ptr;

  OctomapWorld() : handle_(nullptr), robot_size_(1.0, 1.0, 1.0) { create(OctomapParameters()); }   // octomap_world.cc:50-51
  explicit OctomapWorld(const OctomapParameters& params)                                    // :53-56
      : handle_(nullptr), robot_size_(1.0, 1.0, 1.0) { create(params); }
  // Deep copy (octomap_world.cc:59-73).  Like the reference it goes through the BINARY stream
  // (writeOctomapToBinaryConst -> readBinary -> prune), so the copy holds the max-likelihood map:
  // every leaf at the clamping minimum or maximum.
  OctomapWorld(const OctomapWorld& rhs) : handle_(nullptr), robot_size_(rhs.robot_size_) {
    create(rhs.params_);
    size_t n = 0;
    check(vmap_serialize_bt(rhs.handle_, nullptr, 0, &n));
    std::vector<uint8_t> buf(n ? n : 1);
    check(vmap_serialize_bt(rhs.handle_, buf.data(), buf.size(), &n));
    if (vmap_deserialize_bt(handle_, buf.data(), n) != VMAP_OK) std::cerr << "Could not copy octree!\n";
  }
  OctomapWorld& operator=(const OctomapWorld&) = delete;
  virtual ~OctomapWorld() { vmap_destroy(handle_); }

  void resetMap() { check(vmap_reset(handle_)); }                                           // :75-80
  void setOctomapParameters(const OctomapParameters& params) {                              // :84-104
    params_ = params;
    vmap_params p = toC(params);
    check(vmap_set_params(handle_, &p));
  }
  void getOctomapParameters(OctomapParameters* params) const { *params = params_; }         // :106-108
  void enableTreatUnknownAsOccupied() { params_.treat_unknown_as_occupied = true; setOctomapParameters(params_); }
  void disableTreatUnknownAsOccupied() { params_.treat_unknown_as_occupied = false; setOctomapParameters(params_); }
  virtual double getResolution() const { return params_.resolution; }

  // octomap_world.cc:346-360
  virtual CellStatus getCellStatusPoint(const Vector3d& point) const {
    uint8_t s = 0;
    check(vmap_query_points(handle_, point.v, 1, &s));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:395-418
  virtual CellStatus getLineStatus(const Vector3d& start, const Vector3d& end) const {
    const double ab[6] = {start.x(), start.y(), start.z(), end.x(), end.y(), end.z()};
    uint8_t s = 0;
    check(vmap_query_lines(handle_, ab, 1, &s));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:363-373
  virtual CellStatus getCellTrueStatusPoint(const Vector3d& point) const {
    uint8_t s = 0;
    check(vmap_query_points_probability(handle_, point.v, 1, &s, nullptr));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:375-393
  virtual CellStatus getCellProbabilityPoint(const Vector3d& point, double* probability) const {
    uint8_t s = 0;
    check(vmap_query_points_probability(handle_, point.v, 1, &s, probability));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:420-449
  virtual CellStatus getVisibility(const Vector3d& view_point, const Vector3d& voxel_to_test,
                                   bool stop_at_unknown_cell) const {
    const double ab[6] = {view_point.x(), view_point.y(), view_point.z(), voxel_to_test.x(), voxel_to_test.y(), voxel_to_test.z()};
    uint8_t s = 0;
    check(vmap_query_visibility(handle_, ab, 1, stop_at_unknown_cell ? 1 : 0, &s));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:451-493
  virtual CellStatus getLineStatusBoundingBox(const Vector3d& start, const Vector3d& end,
                                              const Vector3d& bounding_box_size) const {
    const double ab[6] = {start.x(), start.y(), start.z(), end.x(), end.y(), end.z()};
    uint8_t s = 0;
    check(vmap_query_lines_bbox(handle_, ab, 1, bounding_box_size.v, &s));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:274-344
  virtual CellStatus getCellStatusBoundingBox(const Vector3d& point, const Vector3d& bounding_box_size) const {
    const double p[3] = {point.x(), point.y(), point.z()};
    const double b[3] = {bounding_box_size.x(), bounding_box_size.y(), bounding_box_size.z()};
    uint8_t s = 0;
    check(vmap_query_bbox_status(handle_, p, 1, b, &s));
    return static_cast<CellStatus>(s);
  }
  // octomap_world.cc:1372-1412: collision checks of the robot's bounding box
  virtual void setRobotSize(const Vector3d& robot_size) { robot_size_ = robot_size; }
  virtual Vector3d getRobotSize() const { return robot_size_; }
  virtual bool checkCollisionWithRobot(const Vector3d& robot_position) {
    std::vector<Vector3d> one(1, robot_position);
    return checkPathForCollisionsWithRobot(one, nullptr);
  }
  virtual bool checkPathForCollisionsWithRobot(const std::vector<Vector3d>& robot_positions, size_t* collision_index) {
    if (robot_positions.empty()) return false;
    std::vector<double> xyz(3 * robot_positions.size());
    for (size_t i = 0; i < robot_positions.size(); i++) {
      xyz[3 * i] = robot_positions[i].x(); xyz[3 * i + 1] = robot_positions[i].y(); xyz[3 * i + 2] = robot_positions[i].z();
    }
    const double b[3] = {robot_size_.x(), robot_size_.y(), robot_size_.z()};
    int collides = 0;
    size_t index = 0;
    check(vmap_check_path_collisions(handle_, xyz.data(), robot_positions.size(), b, &collides, &index));
    if (collides && collision_index != nullptr) *collision_index = index;
    return collides != 0;
  }
  // Batched forms of the two queries (what planners call in bulk): one kernel launch.
  void getCellStatusPoints(const std::vector<Vector3d>& points, std::vector<uint8_t>* status) const {
    status->resize(points.size());
    if (points.empty()) return;
    check(vmap_query_points(handle_, points[0].v, points.size(), status->data()));
  }
  void getLineStatuses(const std::vector<Vector3d>& starts_ends /* start0,end0,start1,end1,... */,
                       std::vector<uint8_t>* status) const {
    status->resize(starts_ends.size() / 2);
    if (status->empty()) return;
    check(vmap_query_lines(handle_, starts_ends[0].v, status->size(), status->data()));
  }

  // octomap_world.cc:497-523: manual edits of a bounding box (setLogOddsBoundingBox :781-818)
  virtual void setFree(const Vector3d& position, const Vector3d& bounding_box_size,
                       const BoundHandling& insertion_method = kDefault) {
    float c[5];
    check(vmap_logodds_constants(handle_, c));
    check(vmap_set_log_odds_bbox(handle_, position.v, 1, bounding_box_size.v, c[2], insertion_method));
  }
  virtual void setOccupied(const Vector3d& position, const Vector3d& bounding_box_size,
                           const BoundHandling& insertion_method = kDefault) {
    float c[5];
    check(vmap_logodds_constants(handle_, c));
    check(vmap_set_log_odds_bbox(handle_, position.v, 1, bounding_box_size.v, c[3], insertion_method));
  }

  // octomap_world.cc:1413-1440 (needs change_detection_enabled; resets the tracking)
  void getChangedPoints(std::vector<Vector3d>* changed_points, std::vector<bool>* changed_states) {
    if (!changed_points) throw std::invalid_argument("getChangedPoints: null output");   // CHECK_NOTNULL
    changed_points->clear();
    if (changed_states) changed_states->clear();
    size_t n = 0;
    const int st = vmap_get_changed(handle_, nullptr, nullptr, 0, &n);
    if (st != VMAP_OK && st != VMAP_ERR_BUFFER_TOO_SMALL) check(st);
    if (!n) return;
    std::vector<uint16_t> k3(n * 3);
    std::vector<uint8_t> occ(n);
    check(vmap_get_changed(handle_, k3.data(), occ.data(), n, &n));
    for (size_t i = 0; i < n; i++) {
      // pointOctomapToEigen(octree_->keyToCoord(key)): float centres widened to double
      Vector3d c;
      for (int a = 0; a < 3; a++) c(a) = (double)(float)((double((int)k3[3 * i + a] - 32768) + 0.5) * params_.resolution);
      changed_points->push_back(c);
      if (changed_states) changed_states->push_back(occ[i] != 0);
    }
  }
  void enableChangeDetection() { params_.change_detection_enabled = true; setOctomapParameters(params_); }
  void disableChangeDetection() { params_.change_detection_enabled = false; setOctomapParameters(params_); }

  // updateOccupancy(KeySet*, KeySet*) (octomap_world.cc:238-264): protected in the reference but
  // called from OctomapManager::loadOctomapCallback (octomap_manager.cc:332-341).
  void updateOccupancy(KeySet* free_cells, KeySet* occupied_cells) {
    if (!free_cells || !occupied_cells) throw std::invalid_argument("updateOccupancy: null key set");   // CHECK_NOTNULL :240-241
    std::vector<uint16_t> f, o;
    f.reserve(free_cells->size() * 3);
    o.reserve(occupied_cells->size() * 3);
    for (const OcTreeKey& k : *occupied_cells) {
      o.push_back(k.k[0]); o.push_back(k.k[1]); o.push_back(k.k[2]);
      free_cells->erase(k);                                                                              // :252-254
    }
    for (const OcTreeKey& k : *free_cells) { f.push_back(k.k[0]); f.push_back(k.k[1]); f.push_back(k.k[2]); }
    check(vmap_apply_keys(handle_, f.data(), f.size() / 3, o.data(), o.size() / 3));
  }

  // On-demand export for the host-side publish / save paths (octomap_world.cc:820-856).
  void exportLeaves(std::vector<OcTreeKey>* keys, std::vector<float>* log_odds) const {
    size_t n = 0;
    check(vmap_leaf_count(handle_, &n));
    std::vector<uint16_t> k3(n * 3);
    log_odds->resize(n);
    if (n) check(vmap_export_leaves(handle_, k3.data(), log_odds->data(), n, &n));
    keys->resize(n);
    for (size_t i = 0; i < n; i++) (*keys)[i] = OcTreeKey(k3[3 * i], k3[3 * i + 1], k3[3 * i + 2]);
  }
  void importLeaves(const std::vector<OcTreeKey>& keys, const std::vector<float>& log_odds) {
    std::vector<uint16_t> k3(keys.size() * 3);
    for (size_t i = 0; i < keys.size(); i++) { k3[3 * i] = keys[i].k[0]; k3[3 * i + 1] = keys[i].k[1]; k3[3 * i + 2] = keys[i].k[2]; }
    check(vmap_import_leaves(handle_, k3.data(), log_odds.data(), keys.size()));
  }

  // octomap_world.cc:846-856.  writeOctomapToFile thresholds the live map like
  // OcTree::writeBinary(filename) does (toMaxLikelihood + prune before writing).
  bool loadOctomapFromFile(const std::string& filename) {
    const bool ok = vmap_load_bt(handle_, filename.c_str()) == VMAP_OK;
    refreshResolution();
    return ok;
  }
  bool writeOctomapToFile(const std::string& filename) { return vmap_write_bt(handle_, filename.c_str(), 1) == VMAP_OK; }
  // OctomapManager::loadOctomapCallback's .pcd / .ply branch (octomap_manager.cc:313-344): the
  // file's points become occupied keys, the free set stays empty.
  bool loadOccupiedPointcloudFile(const std::string& filename) {
    return vmap_load_occupied_pointcloud_file(handle_, filename.c_str(), nullptr) == VMAP_OK;
  }
  bool writeOctomapToBinaryConst(std::ostream& s) const {
    size_t n = 0;
    if (vmap_serialize_bt(handle_, nullptr, 0, &n) != VMAP_OK) return false;
    std::vector<uint8_t> buf(n ? n : 1);
    if (vmap_serialize_bt(handle_, buf.data(), buf.size(), &n) != VMAP_OK) return false;
    s.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)n);
    return s.good();
  }
  // octomap_world.cc:820-825: octomap_msgs::binaryMapToMsg -- id, resolution and the ".bt" node stream
  // WITHOUT its text header (OcTree::writeBinaryData) as the payload.
  bool getOctomapBinaryMsg(OctomapMsg* msg) const {
    if (!msg) throw std::invalid_argument("getOctomapBinaryMsg: null message");
    size_t n = 0;
    if (vmap_serialize_bt(handle_, nullptr, 0, &n) != VMAP_OK) return false;
    std::vector<uint8_t> buf(n ? n : 1);
    if (vmap_serialize_bt(handle_, buf.data(), buf.size(), &n) != VMAP_OK) return false;
    static const char kData[] = "\ndata\n";
    size_t body = n;
    for (size_t i = 0; i + 6 <= n; i++)
      if (std::memcmp(buf.data() + i, kData, 6) == 0) { body = i + 6; break; }
    msg->binary = true;
    msg->id = "OcTree";
    msg->resolution = params_.resolution;
    msg->data.assign(reinterpret_cast<const int8_t*>(buf.data()) + body, reinterpret_cast<const int8_t*>(buf.data()) + n);
    return true;
  }
  // octomap_world.cc:826-831: octomap_msgs::fullMapToMsg -- OcTree::writeData as the payload.
  bool getOctomapFullMsg(OctomapMsg* msg) const {
    if (!msg) throw std::invalid_argument("getOctomapFullMsg: null message");
    size_t n = 0;
    if (vmap_serialize_ot(handle_, 0, nullptr, 0, &n) != VMAP_OK) return false;
    std::vector<uint8_t> buf(n ? n : 1);
    if (vmap_serialize_ot(handle_, 0, buf.data(), buf.size(), &n) != VMAP_OK) return false;
    msg->binary = false;
    msg->id = "OcTree";
    msg->resolution = params_.resolution;
    msg->data.assign(reinterpret_cast<const int8_t*>(buf.data()), reinterpret_cast<const int8_t*>(buf.data()) + n);
    return true;
  }
  // octomap_world.cc:833-839
  void setOctomapFromMsg(const OctomapMsg& msg) {
    if (msg.binary) setOctomapFromBinaryMsg(msg);
    else setOctomapFromFullMsg(msg);
  }

  // octree_->prune(); octree_->size()
  size_t getNumNodesPruned() const {
    size_t n = 0;
    check(vmap_tree_size(handle_, 0, &n));
    return n;
  }

  vmap_scan_stats lastScanStats() const {
    vmap_scan_stats s;
    check(vmap_last_scan_stats(handle_, &s));
    return s;
  }
  vmap* handle() const { return handle_; }

 protected:
  // octomap_world.cc:110-141.  The reference mutates the caller's cloud in place (NaN filter,
  // world-frame transform); so does this.
  virtual void insertPointcloudIntoMapImpl(const Transformation& T_G_sensor,
                                           const PointCloud<PointXYZ>::Ptr& pointcloud) {
    const Matrix4d T = T_G_sensor.getTransformationMatrix();
    size_t n_out = pointcloud->points.size();
    float* xyz = pointcloud->points.empty() ? nullptr : &pointcloud->points[0].x;
    check(vmap_insert_pointcloud(handle_, xyz, pointcloud->points.size(), sizeof(PointXYZ),
                                 pointcloud->is_dense ? 1 : 0, T.m, xyz, &n_out));
    if (!pointcloud->is_dense) {
      pointcloud->points.resize(n_out);
      pointcloud->width = (uint32_t)n_out;
      pointcloud->height = 1;
      pointcloud->is_dense = true;
    }
  }
  // octomap_world.cc:143-175
  virtual void insertProjectedDisparityIntoMapImpl(const Transformation& sensor_to_world,
                                                   const Mat& projected_points) {
    const Vector3d t = sensor_to_world.getPosition();
    check(vmap_insert_projected(handle_, projected_points.data, projected_points.rows, projected_points.cols,
                                projected_points.step, sensor_to_world.quaternion_wxyz(), t.v));
  }
  // world_base.cc:51-87 fused with the above on the device.
  virtual void insertDisparityImageImpl(const Transformation& sensor_to_world, const Mat& disparity,
                                        const Matrix4d& Q_full, const Vector2d& full_image_size) {
    const Vector3d t = sensor_to_world.getPosition();
    check(vmap_insert_disparity(handle_, disparity.data, disparity.rows, disparity.cols, disparity.step,
                                Q_full.m, full_image_size.x(), sensor_to_world.quaternion_wxyz(), t.v));
  }

  // octomap_world.cc:781-818: manually affect the probabilities of areas within a bounding box
  // (protected in the reference as well; setFree / setOccupied pass the clamping thresholds).
  void setLogOddsBoundingBox(const Vector3d& position, const Vector3d& bounding_box_size, double log_odds_value,
                             const BoundHandling& insertion_method = kDefault) {
    check(vmap_set_log_odds_bbox(handle_, position.v, 1, bounding_box_size.v, log_odds_value, insertion_method));
  }
  void setLogOddsBoundingBox(const std::vector<Vector3d>& positions, const Vector3d& bounding_box_size,
                             double log_odds_value, const BoundHandling& insertion_method = kDefault) {
    if (positions.empty()) return;
    std::vector<double> xyz(3 * positions.size());
    for (size_t i = 0; i < positions.size(); i++) {
      xyz[3 * i] = positions[i].x(); xyz[3 * i + 1] = positions[i].y(); xyz[3 * i + 2] = positions[i].z();
    }
    check(vmap_set_log_odds_bbox(handle_, xyz.data(), positions.size(), bounding_box_size.v, log_odds_value, insertion_method));
  }
  // octomap_world.cc:841-844 (binaryMsgToMap): the payload is the node stream of a ".bt" file; the tree takes
  // the message's resolution.
  void setOctomapFromBinaryMsg(const OctomapMsg& msg) {
    char head[160];
    const int n_head = std::snprintf(head, sizeof(head), "# Octomap OcTree binary file\nid OcTree\nsize %zu\nres %.17g\ndata\n",
                                     msg.data.empty() ? (size_t)0 : (size_t)1, msg.resolution);
    std::vector<uint8_t> buf(head, head + n_head);
    buf.insert(buf.end(), reinterpret_cast<const uint8_t*>(msg.data.data()),
               reinterpret_cast<const uint8_t*>(msg.data.data()) + msg.data.size());
    if (vmap_deserialize_bt(handle_, buf.data(), buf.size()) != VMAP_OK) std::cerr << "Could not read the binary octomap message!\n";
    refreshResolution();
  }
  // octomap_world.cc:845-848 (fullMsgToMap): OcTree::readData on a tree of the message's resolution.
  void setOctomapFromFullMsg(const OctomapMsg& msg) {
    if (msg.resolution != params_.resolution) {     // a new tree of that resolution
      params_.resolution = msg.resolution;
      setOctomapParameters(params_);
    }
    const uint8_t* p = reinterpret_cast<const uint8_t*>(msg.data.data());
    const uint8_t none = 0;
    if (vmap_deserialize_ot(handle_, 0, msg.data.empty() ? &none : p, msg.data.size()) != VMAP_OK)
      std::cerr << "Could not read the full octomap message!\n";
  }

  OctomapParameters params_;

 private:
  void refreshResolution() {     // a loaded stream brings its own resolution
    vmap_params p;
    if (vmap_get_params(handle_, &p) == VMAP_OK) params_.resolution = p.resolution;
  }
  static vmap_params toC(const OctomapParameters& q) {
    vmap_params p;
    vmap_default_params(&p);
    p.resolution = q.resolution; p.probability_hit = q.probability_hit; p.probability_miss = q.probability_miss;
    p.threshold_min = q.threshold_min; p.threshold_max = q.threshold_max; p.threshold_occupancy = q.threshold_occupancy;
    p.filter_speckles = q.filter_speckles; p.max_free_space = q.max_free_space;
    p.min_height_free_space = q.min_height_free_space; p.sensor_max_range = q.sensor_max_range;
    p.visualize_min_z = q.visualize_min_z; p.visualize_max_z = q.visualize_max_z;
    p.treat_unknown_as_occupied = q.treat_unknown_as_occupied; p.change_detection_enabled = q.change_detection_enabled;
    return p;
  }
  void create(const OctomapParameters& params) {
    params_ = params;
    vmap_params p = toC(params);
    const int st = vmap_create(&p, nullptr, &handle_);
    if (st != VMAP_OK) throw std::runtime_error(std::string("OctomapWorld: ") + vmap_last_error(nullptr));
  }
  void check(int st) const {
    if (st != VMAP_OK) throw std::runtime_error(std::string("OctomapWorld: ") + vmap_last_error(handle_));
  }
  vmap* handle_;
  Vector3d robot_size_;   // Eigen::Vector3d::Ones() by default (octomap_world.cc:54)
};

}  // namespace volumetric_mapping
#endif  // VOLUMETRIC_MAPPING_B200_OCTOMAP_WORLD_H_
