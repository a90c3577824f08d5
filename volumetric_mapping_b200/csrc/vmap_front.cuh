// vmap_front.cuh -- per-point arithmetic of the front-end kernels (K1), free of block- and
// warp-level primitives so that the CPU test-suite can run it against the cv2 golden vectors and
// the oracle (tests/native/front_check.cc):
//   matrix_transform   pcl::transformPointCloud(Matrix4d) on one point (octomap_world.cc:118-119)
//   quat_transform     kindr QuatTransformation * point (:159-163)
//   reproject_pixel    cv::reprojectImageTo3D, OpenCV 4.13 arithmetic (world_base.cc:71-75)
//   projected_valid    isValidPoint && z >= 0 (:156, 232-236)
//   classify_math      castRay's range test / clip (:184-185, 210-212), the endpoint key and its
//                      bounds check (:131, 166, 205), and the step count that orders the work list
#pragma once

#include "vmap_device.cuh"

namespace vmapdev {

struct Transform {
  double m[12];   // row-major 3x4 of getTransformationMatrix()
  double q[4];    // w x y z
  double t[3];
  float origin[3];
};

// static_cast<float>(m(r,0)*x + m(r,1)*y + m(r,2)*z + m(r,3)), left to right, double
__device__ __forceinline__ void matrix_transform(const Transform& T, float fx, float fy, float fz, float* wx, float* wy,
                                                 float* wz) {
  const double x = fx, y = fy, z = fz;
  *wx = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[0], x), __dmul_rn(T.m[1], y)), __dmul_rn(T.m[2], z)), T.m[3]);
  *wy = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[4], x), __dmul_rn(T.m[5], y)), __dmul_rn(T.m[6], z)), T.m[7]);
  *wz = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[8], x), __dmul_rn(T.m[9], y)), __dmul_rn(T.m[10], z)), T.m[11]);
}

// Eigen::Quaterniond::_transformVector(v) + t  (kindr QuatTransformation::operator*)
__device__ __forceinline__ void quat_transform(const Transform& T, double vx, double vy, double vz,
                                               double* ox, double* oy, double* oz) {
  const double w = T.q[0], x = T.q[1], y = T.q[2], z = T.q[3];
  double u0 = __dsub_rn(__dmul_rn(y, vz), __dmul_rn(z, vy));
  double u1 = __dsub_rn(__dmul_rn(z, vx), __dmul_rn(x, vz));
  double u2 = __dsub_rn(__dmul_rn(x, vy), __dmul_rn(y, vx));
  u0 = __dadd_rn(u0, u0); u1 = __dadd_rn(u1, u1); u2 = __dadd_rn(u2, u2);
  const double c0 = __dsub_rn(__dmul_rn(y, u2), __dmul_rn(z, u1));
  const double c1 = __dsub_rn(__dmul_rn(z, u0), __dmul_rn(x, u2));
  const double c2 = __dsub_rn(__dmul_rn(x, u1), __dmul_rn(y, u0));
  *ox = __dadd_rn(__dadd_rn(__dadd_rn(vx, __dmul_rn(w, u0)), c0), T.t[0]);
  *oy = __dadd_rn(__dadd_rn(__dadd_rn(vy, __dmul_rn(w, u1)), c1), T.t[1]);
  *oz = __dadd_rn(__dadd_rn(__dadd_rn(vz, __dmul_rn(w, u2)), c2), T.t[2]);
}

// isValidPoint(pt) && !(pt.z < 0)   (octomap_world.cc:156, 232-236)
__device__ __forceinline__ bool projected_valid(float sz) { return (sz != 10000.0f) && !isinf(sz) && !(sz < 0.0f); }

// cv::reprojectImageTo3D, OpenCV 4.13 arithmetic: one pixel.
__device__ __forceinline__ void reproject_pixel(const double* Q, double min_disp, uint32_t u,
                                                uint32_t v, float disp, float* X, float* Y, float* Z) {
  const double x = (double)u, y = (double)v, d = (double)disp;
  double h[4];
#pragma unroll
  for (int r = 0; r < 4; r++)
    h[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Q[4 * r], x), __dmul_rn(Q[4 * r + 1], y)),
                               __dmul_rn(Q[4 * r + 2], d)),
                     __dmul_rn(Q[4 * r + 3], 1.0));
  const double iw = __ddiv_rn(1.0, h[3]);
  *X = (float)__dmul_rn((double)(float)h[0], iw);
  *Y = (float)__dmul_rn((double)(float)h[1], iw);
  *Z = (float)__dmul_rn((double)(float)h[2], iw);
  if (fabs(__dsub_rn(d, min_disp)) <= 1.1920928955078125e-07) *Z = 10000.0f;
}

struct QMat { double q[16]; };

struct PointClass {
  unsigned long long key;   // unchecked endpoint key k0 | k1 << 16 | k2 << 32
  bool bound;               // coordToKeyChecked(point) succeeds
  bool in_range;            // castRay's in-range branch
  float ex, ey, ez;         // ray end: the point, or the clipped end when out of range
  uint32_t len;             // DDA steps of the ray (key-space Manhattan distance), 0 when it is not walked
};
__device__ __forceinline__ PointClass classify_math(const DevParams& P, const float* origin, float px, float py, float pz) {
  PointClass c;
  const int s0 = scaled_coord(P.res_factor, (double)px);
  const int s1 = scaled_coord(P.res_factor, (double)py);
  const int s2 = scaled_coord(P.res_factor, (double)pz);
  c.key = (unsigned long long)((uint32_t)s0 & 0xFFFFu) | ((unsigned long long)((uint32_t)s1 & 0xFFFFu) << 16) |
          ((unsigned long long)((uint32_t)s2 & 0xFFFFu) << 32);
  c.bound = key_in_range(s0) && key_in_range(s1) && key_in_range(s2);
  // castRay: (point - sensor_origin).norm() <= sensor_max_range, norm() in double on a float sum
  float dx = __fsub_rn(px, origin[0]), dy = __fsub_rn(py, origin[1]), dz = __fsub_rn(pz, origin[2]);
  const double len = __dsqrt_rn((double)norm_sq_f(dx, dy, dz));
  const bool in_range = (P.max_range < 0.0) || (len <= P.max_range);
  float ex = px, ey = py, ez = pz;
  if (!in_range) {
    // new_end = origin + (point - origin).normalized() * sensor_max_range   (all float)
    if (len > 0) {
      const float fl = (float)len;
      dx = __fdiv_rn(dx, fl); dy = __fdiv_rn(dy, fl); dz = __fdiv_rn(dz, fl);
    }
    const float mr = (float)P.max_range;
    ex = __fadd_rn(origin[0], __fmul_rn(dx, mr));
    ey = __fadd_rn(origin[1], __fmul_rn(dy, mr));
    ez = __fadd_rn(origin[2], __fmul_rn(dz, mr));
  }
  c.in_range = in_range;
  c.ex = ex; c.ey = ey; c.ez = ez;
  {
    // number of DDA steps the ray will take (exact up to rounding): orders the work list
    const int o0 = scaled_coord(P.res_factor, (double)origin[0]), o1 = scaled_coord(P.res_factor, (double)origin[1]),
              o2 = scaled_coord(P.res_factor, (double)origin[2]);
    const int e0 = scaled_coord(P.res_factor, (double)ex), e1 = scaled_coord(P.res_factor, (double)ey),
              e2 = scaled_coord(P.res_factor, (double)ez);
    const bool walk = key_in_range(o0) && key_in_range(o1) && key_in_range(o2) && key_in_range(e0) &&
                      key_in_range(e1) && key_in_range(e2);
    c.len = walk ? (uint32_t)(abs(e0 - o0) + abs(e1 - o1) + abs(e2 - o2)) : 0u;
  }
  return c;
}

}  // namespace vmapdev
