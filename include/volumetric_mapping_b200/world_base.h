// world_base.h -- host-side mirror of volumetric_mapping::WorldBase for the scan-insertion path
// (reference: volumetric_map_base/include/volumetric_map_base/world_base.h:50-235 and
// volumetric_map_base/src/world_base.cc:41-87,146-213).  Same public entry points, same
// protected virtual hooks, same "empty world" query defaults.
#ifndef VOLUMETRIC_MAPPING_B200_WORLD_BASE_H_
#define VOLUMETRIC_MAPPING_B200_WORLD_BASE_H_

#include <limits>
#include <memory>

#include "volumetric_mapping_b200/compat_types.h"

namespace volumetric_mapping {

using compat::Mat;
using compat::Matrix3Xd;
using compat::Matrix4d;
using compat::PointCloud;
using compat::PointXYZ;
using compat::Transformation;
using compat::Vector2d;
using compat::Vector3d;

class WorldBase {
 public:
  typedef std::shared_ptr<WorldBase> Ptr;
  enum CellStatus { kFree = 0, kOccupied = 1, kUnknown = 2 };   // world_base.h:54

  WorldBase() {}
  virtual ~WorldBase() {}

  // world_base.cc:51-87.  `disparity` is CV_32FC1.  The reference rescales Q for a
  // down-sampled image (:55-69), calls cv::reprojectImageTo3D(..., handleMissingValues=true)
  // (:71-75) and dispatches to insertProjectedDisparityIntoMapImpl (:78-81).  A subclass whose
  // back end fuses those steps overrides insertDisparityImageImpl instead.
  void insertDisparityImage(const Transformation& sensor_to_world, const Mat& disparity,
                            const Matrix4d& Q_full, const Vector2d& full_image_size) {
    insertDisparityImageImpl(sensor_to_world, disparity, Q_full, full_image_size);
  }

  // world_base.cc:194-200: Eigen::Matrix3Xd -> float cloud (is_dense = true), then :202-213.
  void insertPointcloud(const Transformation& T_G_sensor, const Matrix3Xd& pointcloud_sensor) {
    PointCloud<PointXYZ>::Ptr cloud(new PointCloud<PointXYZ>);
    cloud->points.resize(pointcloud_sensor.cols());
    for (size_t i = 0; i < pointcloud_sensor.cols(); i++) {
      cloud->points[i].x = static_cast<float>(pointcloud_sensor(0, i));
      cloud->points[i].y = static_cast<float>(pointcloud_sensor(1, i));
      cloud->points[i].z = static_cast<float>(pointcloud_sensor(2, i));
    }
    cloud->width = (uint32_t)cloud->points.size();
    cloud->is_dense = true;
    insertPointcloud(T_G_sensor, cloud);
  }
  // world_base.cc:202-213 (no point weighing: OctomapWorld never implements the weighted hooks).
  void insertPointcloud(const Transformation& T_G_sensor,
                        const PointCloud<PointXYZ>::Ptr& pointcloud_sensor) {
    insertPointcloudIntoMapImpl(T_G_sensor, pointcloud_sensor);
  }

  // world_base.h:105-122: a valid, completely empty world.
  virtual CellStatus getCellStatusPoint(const Vector3d&) const { return kFree; }
  virtual CellStatus getLineStatus(const Vector3d&, const Vector3d&) const { return kFree; }
  virtual CellStatus getCellStatusBoundingBox(const Vector3d&, const Vector3d&) const { return kFree; }
  virtual CellStatus getLineStatusBoundingBox(const Vector3d&, const Vector3d&, const Vector3d&) const { return kFree; }
  virtual Vector3d getMapCenter() const { return Vector3d::Zero(); }
  virtual Vector3d getMapSize() const {
    const double m = std::numeric_limits<double>::max();
    return Vector3d(m, m, m);
  }

  // world_base.cc:146-180 (generateQ): Q from rectified camera parameters.
  Matrix4d generateQ(double Tx, double left_cx, double left_cy, double left_fx, double left_fy,
                     double right_cx, double /*right_cy*/, double /*right_fx*/, double /*right_fy*/) const {
    Matrix4d Q = Matrix4d::Zero();
    Q(0, 0) = left_fy * Tx;
    Q(0, 3) = -left_fy * left_cx * Tx;
    Q(1, 1) = left_fx * Tx;
    Q(1, 3) = -left_fx * left_cy * Tx;
    Q(2, 3) = left_fx * left_fy * Tx;
    Q(3, 2) = -left_fy;
    Q(3, 3) = left_fy * (left_cx - right_cx);
    return Q;
  }

 protected:
  // world_base.h:199-213
  virtual void insertProjectedDisparityIntoMapImpl(const Transformation&, const Mat& /*CV_32FC3*/) {}
  virtual void insertPointcloudIntoMapImpl(const Transformation&, const PointCloud<PointXYZ>::Ptr&) {}
  // Fused disparity entry (Q fix-up + reprojection + insertion in one device pass).
  virtual void insertDisparityImageImpl(const Transformation&, const Mat&, const Matrix4d&, const Vector2d&) {}
};

}  // namespace volumetric_mapping
#endif  // VOLUMETRIC_MAPPING_B200_WORLD_BASE_H_
