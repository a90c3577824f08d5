// compat_types.h -- stand-ins for the dependency types that appear in the reference's
// WorldBase / OctomapWorld signatures (Eigen, kindr::minimal, pcl, cv::Mat, octomap keys).
//
// None of those libraries exists in this build environment.  The host-side mirror classes
// (world_base.h, octomap_world.h in this directory) are written against the small subset of
// each type that the scan-insertion path touches; a maintainer with the real libraries defines
// VMAP_USE_REAL_DEPS and the same names resolve to Eigen / kindr / pcl / OpenCV / octomap
// instead.  Nothing in here crosses the C ABI: include/vmap/vmap.h only sees
// plain pointers and sizes.
#ifndef VOLUMETRIC_MAPPING_B200_COMPAT_TYPES_H_
#define VOLUMETRIC_MAPPING_B200_COMPAT_TYPES_H_

#if defined(VMAP_USE_REAL_DEPS)
#include <Eigen/Core>
#include <kindr/minimal/quat-transformation.h>
#include <octomap/octomap.h>
#include <opencv2/core/core.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace volumetric_mapping {
namespace compat {
using Vector2d = Eigen::Vector2d;
using Vector3d = Eigen::Vector3d;
using Matrix4d = Eigen::Matrix4d;
using Matrix3Xd = Eigen::Matrix3Xd;
using Transformation = kindr::minimal::QuatTransformation;
using PointXYZ = pcl::PointXYZ;
template <class P> using PointCloud = pcl::PointCloud<P>;
using Mat = cv::Mat;
using OcTreeKey = octomap::OcTreeKey;
using KeySet = octomap::KeySet;
}  // namespace compat
}  // namespace volumetric_mapping
#include <octomap_msgs/Octomap.h>
namespace volumetric_mapping {
namespace compat {
using OctomapMsg = octomap_msgs::Octomap;
}  // namespace compat
}  // namespace volumetric_mapping
#else

#include <stddef.h>
#include <stdint.h>

#include <cmath>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

namespace volumetric_mapping {
namespace compat {

struct Vector2d {
  double v[2];
  Vector2d() : v{0, 0} {}
  Vector2d(double x, double y) : v{x, y} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};

struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  static Vector3d Zero() { return Vector3d(); }
};

struct Matrix4d {   // row(r), col(c) access like Eigen; storage row-major for the C ABI
  double m[16];
  Matrix4d() {
    for (int i = 0; i < 16; i++) m[i] = 0.0;
  }
  double& operator()(int r, int c) { return m[4 * r + c]; }
  double operator()(int r, int c) const { return m[4 * r + c]; }
  static Matrix4d Identity() {
    Matrix4d I;
    for (int i = 0; i < 4; i++) I(i, i) = 1.0;
    return I;
  }
  static Matrix4d Zero() { return Matrix4d(); }
};

struct Matrix3Xd {   // 3 x N, column = point
  std::vector<double> d;
  Matrix3Xd() {}
  explicit Matrix3Xd(size_t n) : d(3 * n, 0.0) {}
  size_t cols() const { return d.size() / 3; }
  double& operator()(int r, size_t c) { return d[3 * c + r]; }
  double operator()(int r, size_t c) const { return d[3 * c + r]; }
};

// kindr::minimal::QuatTransformation: unit quaternion (w, x, y, z) + translation.
class Transformation {
 public:
  Transformation() : q_{1, 0, 0, 0}, t_{0, 0, 0} {}
  Transformation(double qw, double qx, double qy, double qz, const Vector3d& t)
      : q_{qw, qx, qy, qz}, t_{t.x(), t.y(), t.z()} {}
  const double* quaternion_wxyz() const { return q_; }
  Vector3d getPosition() const { return Vector3d(t_[0], t_[1], t_[2]); }
  // Eigen::Quaterniond::toRotationMatrix() (same association order) + translation.
  Matrix4d getTransformationMatrix() const {
    const double w = q_[0], x = q_[1], y = q_[2], z = q_[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Matrix4d T = Matrix4d::Identity();
    T(0, 0) = 1.0 - (tyy + tzz); T(0, 1) = txy - twz; T(0, 2) = txz + twy;
    T(1, 0) = txy + twz; T(1, 1) = 1.0 - (txx + tzz); T(1, 2) = tyz - twx;
    T(2, 0) = txz - twy; T(2, 1) = tyz + twx; T(2, 2) = 1.0 - (txx + tyy);
    T(0, 3) = t_[0]; T(1, 3) = t_[1]; T(2, 3) = t_[2];
    return T;
  }

 private:
  double q_[4];
  double t_[3];
};

struct PointXYZ {   // pcl::PointXYZ: 16-byte stride
  float x, y, z, pad;
  PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad(1.0f) {}
};

template <class P>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<P> > Ptr;
  std::vector<P> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  void push_back(const P& p) { points.push_back(p); width = (uint32_t)points.size(); }
};

// cv::Mat view: rows x cols of `channels` floats, `step` bytes between rows.
struct Mat {
  int rows = 0, cols = 0, channels = 1;
  size_t step = 0;
  std::shared_ptr<std::vector<float> > storage;
  float* data = nullptr;
  Mat() {}
  Mat(int r, int c, int ch) : rows(r), cols(c), channels(ch), step((size_t)c * ch * sizeof(float)),
                              storage(new std::vector<float>((size_t)r * c * ch)) { data = storage->data(); }
  float* ptr(int r) { return reinterpret_cast<float*>(reinterpret_cast<char*>(data) + (size_t)r * step); }
  const float* ptr(int r) const { return reinterpret_cast<const float*>(reinterpret_cast<const char*>(data) + (size_t)r * step); }
};

struct OcTreeKey {   // octomap::OcTreeKey
  uint16_t k[3];
  OcTreeKey() : k{0, 0, 0} {}
  OcTreeKey(uint16_t a, uint16_t b, uint16_t c) : k{a, b, c} {}
  bool operator==(const OcTreeKey& o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
  uint16_t operator[](int i) const { return k[i]; }
  struct KeyHash {
    size_t operator()(const OcTreeKey& key) const {
      return static_cast<size_t>(key.k[0]) + 1447 * static_cast<size_t>(key.k[1]) + 345637 * static_cast<size_t>(key.k[2]);
    }
  };
};
typedef std::unordered_set<OcTreeKey, OcTreeKey::KeyHash> KeySet;

struct OctomapMsg {   // octomap_msgs::Octomap without the ROS header: what binaryMapToMsg / fullMapToMsg fill
  bool binary;
  std::string id;
  double resolution;
  std::vector<int8_t> data;
  OctomapMsg() : binary(false), resolution(0.0) {}
};

}  // namespace compat
}  // namespace volumetric_mapping
#endif  // VMAP_USE_REAL_DEPS
#endif  // VOLUMETRIC_MAPPING_B200_COMPAT_TYPES_H_
