// demo_octomap_world.cc -- the reference's library usage ("can be run outside of a ROS node",
// README.md:41) against the B200 back end: C++ host code -> OctomapWorld mirror -> C ABI -> CUDA.
// Exits non-zero when a hand-derived known answer (tests/test_oracle_insertion.py) is not met.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "volumetric_mapping_b200/octomap_world.h"

using namespace volumetric_mapping;

#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) {                                                        \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                           \
    }                                                                     \
  } while (0)

int main() {
  OctomapParameters params;
  params.resolution = 0.1;
  params.sensor_max_range = 5.0;
  OctomapWorld world(params);

  // one point 0.55 m ahead of a sensor at the origin: 5 free voxels + 1 occupied endpoint
  PointCloud<PointXYZ>::Ptr cloud(new PointCloud<PointXYZ>);
  cloud->push_back(PointXYZ(0.55f, 0.05f, 0.05f));
  Transformation T;   // identity
  world.insertPointcloud(T, cloud);
  vmap_scan_stats st = world.lastScanStats();
  EXPECT(st.traversals == 5 && st.rays_cast == 1 && st.unique_free == 5 && st.unique_occupied == 1);

  std::vector<OcTreeKey> keys;
  std::vector<float> lo;
  world.exportLeaves(&keys, &lo);
  EXPECT(keys.size() == 6);

  // one hit is not yet "occupied" (0.619 < 0.847); two hits are
  EXPECT(world.getCellStatusPoint(Vector3d(0.55, 0.05, 0.05)) == WorldBase::kFree);
  world.insertPointcloud(T, cloud);
  EXPECT(world.getCellStatusPoint(Vector3d(0.55, 0.05, 0.05)) == WorldBase::kOccupied);
  EXPECT(world.getCellStatusPoint(Vector3d(0.25, 0.05, 0.05)) == WorldBase::kFree);
  EXPECT(world.getCellStatusPoint(Vector3d(0.25, 0.55, 0.05)) == WorldBase::kOccupied);   // unknown -> occupied
  world.disableTreatUnknownAsOccupied();
  EXPECT(world.getCellStatusPoint(Vector3d(0.25, 0.55, 0.05)) == WorldBase::kUnknown);
  // line through the free voxels stops at the occupied endpoint voxel ... which is the END voxel
  // of this segment and therefore never tested (octomap_world.cc:401-417) -> free
  EXPECT(world.getLineStatus(Vector3d(0.05, 0.05, 0.05), Vector3d(0.55, 0.05, 0.05)) == WorldBase::kFree);
  EXPECT(world.getLineStatus(Vector3d(0.05, 0.05, 0.05), Vector3d(0.75, 0.05, 0.05)) == WorldBase::kOccupied);
  EXPECT(world.getLineStatus(Vector3d(0.05, 0.05, 0.05), Vector3d(0.05, 0.45, 0.05)) == WorldBase::kUnknown);

  // Matrix3Xd overload, NaN handling of a non-dense pcl cloud, updateOccupancy on key sets
  Matrix3Xd m(2);
  m(0, 0) = -0.55; m(1, 0) = 0.05; m(2, 0) = 0.05;
  m(0, 1) = 0.05; m(1, 1) = -0.75; m(2, 1) = 0.05;
  world.insertPointcloud(T, m);
  EXPECT(world.lastScanStats().rays_cast == 2);
  PointCloud<PointXYZ>::Ptr nan_cloud(new PointCloud<PointXYZ>);
  nan_cloud->push_back(PointXYZ(0.05f, 0.05f, 0.95f));
  nan_cloud->push_back(PointXYZ(NAN, 0.f, 0.f));
  nan_cloud->is_dense = false;
  world.insertPointcloud(T, nan_cloud);
  EXPECT(nan_cloud->points.size() == 1 && nan_cloud->is_dense);
  KeySet free_cells, occupied_cells;
  free_cells.insert(OcTreeKey(32900, 32900, 32900));
  free_cells.insert(OcTreeKey(32901, 32900, 32900));
  occupied_cells.insert(OcTreeKey(32901, 32900, 32900));   // occupied wins
  world.updateOccupancy(&free_cells, &occupied_cells);
  EXPECT(world.lastScanStats().unique_free == 1 && world.lastScanStats().unique_occupied == 1);

  // a 64 x 48 disparity image through the fused path
  Mat disp(48, 64, 1);
  for (int v = 0; v < 48; v++)
    for (int u = 0; u < 64; u++) disp.ptr(v)[u] = (u % 7 == 0) ? -1.0f : 20.0f + 0.5f * u + 0.25f * v;
  const Matrix4d Q = world.generateQ(-0.110074, 32.0, 24.0, 45.0, 45.0, 31.5, 24.0, 45.0, 45.0);
  Transformation cam(0.5, -0.5, 0.5, -0.5, Vector3d(0.3, -0.2, 1.1));
  world.insertDisparityImage(cam, disp, Q, Vector2d(64.0, 48.0));
  st = world.lastScanStats();
  EXPECT(st.points_in == 64 * 48 && st.points_valid > 0 && st.rays_cast > 0);
  std::printf("demo ok: %llu valid pixels, %llu rays, %llu traversals, %.3f ms on the device\n",
              (unsigned long long)st.points_valid, (unsigned long long)st.rays_cast,
              (unsigned long long)st.traversals, st.gpu_ms);
  // save / load round trip through the octomap binary format
  std::ostringstream bt;
  EXPECT(world.writeOctomapToBinaryConst(bt));
  EXPECT(bt.str().compare(0, 28, "# Octomap OcTree binary file") == 0);
  EXPECT(world.writeOctomapToFile("/tmp/vmap_demo.bt"));
  OctomapWorld other;
  EXPECT(other.loadOctomapFromFile("/tmp/vmap_demo.bt"));
  std::ostringstream bt2;
  EXPECT(other.writeOctomapToBinaryConst(bt2) && bt2.str() == bt.str());
  // the two message forms (octomap_world.cc:820-844): binary = the .bt node stream without its header,
  // full = OcTree::writeData; a world of another resolution takes the message's
  OctomapMsg bin_msg, full_msg;
  EXPECT(world.getOctomapBinaryMsg(&bin_msg) && bin_msg.binary && bin_msg.id == "OcTree" && bin_msg.resolution == 0.1);
  EXPECT(!bin_msg.data.empty() && bin_msg.data.size() < bt.str().size());
  EXPECT(std::memcmp(bin_msg.data.data(), bt.str().data() + (bt.str().size() - bin_msg.data.size()), bin_msg.data.size()) == 0);
  OctomapWorld from_bin;                       // 0.15 m by default
  from_bin.setOctomapFromMsg(bin_msg);
  EXPECT(from_bin.getResolution() == 0.1);
  std::ostringstream bt3;
  EXPECT(from_bin.writeOctomapToBinaryConst(bt3) && bt3.str() == bt.str());
  EXPECT(world.getOctomapFullMsg(&full_msg) && !full_msg.binary && full_msg.resolution == 0.1);
  OctomapWorld from_full;
  from_full.setOctomapFromMsg(full_msg);
  EXPECT(from_full.getResolution() == 0.1);
  OctomapMsg full_again;
  EXPECT(from_full.getOctomapFullMsg(&full_again) && full_again.data == full_msg.data);
  std::vector<OcTreeKey> keys2;
  std::vector<float> lo2;
  world.exportLeaves(&keys, &lo);
  from_full.exportLeaves(&keys2, &lo2);
  EXPECT(keys2.size() == keys.size() && !keys.empty());
  // box edits: setOccupied writes the clamping maximum, setFree the minimum (octomap_world.cc:497-523)
  world.setOccupied(Vector3d(2.05, 2.05, 2.05), Vector3d(0.3, 0.3, 0.3));
  EXPECT(world.getCellStatusPoint(Vector3d(2.05, 2.05, 2.05)) == WorldBase::kOccupied);
  world.setFree(Vector3d(2.05, 2.05, 2.05), Vector3d(0.3, 0.3, 0.3));
  EXPECT(world.getCellStatusPoint(Vector3d(2.05, 2.05, 2.05)) == WorldBase::kFree);
  world.resetMap();
  world.exportLeaves(&keys, &lo);
  EXPECT(keys.empty());
  return 0;
}
