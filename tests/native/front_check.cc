// Host build of the product's front-end arithmetic (vmap_front.cuh) for the CPU test-suite, driven
// by tests/test_front_host_emulation.py: projection against the cv2 golden vectors, the two
// transforms and castRay's range test / clip against independent restatements.
#include "../../volumetric_mapping_b200/csrc/vmap_front.cuh"

using namespace vmapdev;

extern "C" {

// cv::reprojectImageTo3D(handleMissingValues = true) on a rows x cols float image
void fc_reproject(const float* disparity, int rows, int cols, const double* Q16, float* xyz) {
  // image minimum (what k_min_disparity computes; NaNs are skipped, FLT_MAX for an all-NaN image)
  double min_disp = 3.4028234663852886e38;
  bool first = true;
  for (long i = 0; i < (long)rows * cols; i++)
    if (disparity[i] == disparity[i] && (first || (double)disparity[i] < min_disp)) { min_disp = disparity[i]; first = false; }
  for (int v = 0; v < rows; v++)
    for (int u = 0; u < cols; u++) {
      float* o = xyz + 3 * ((long)v * cols + u);
      reproject_pixel(Q16, min_disp, (uint32_t)u, (uint32_t)v, disparity[(long)v * cols + u], o, o + 1, o + 2);
    }
}

void fc_matrix_transform(const double* T16, const float* xyz, size_t n, float* out) {
  Transform T;
  memset(&T, 0, sizeof(T));
  for (int i = 0; i < 12; i++) T.m[i] = T16[i];
  for (size_t i = 0; i < n; i++) matrix_transform(T, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], out + 3 * i, out + 3 * i + 1, out + 3 * i + 2);
}

void fc_quat_transform(const double* q_wxyz, const double* t3, const float* xyz, size_t n, float* out, uint8_t* valid) {
  Transform T;
  memset(&T, 0, sizeof(T));
  for (int i = 0; i < 4; i++) T.q[i] = q_wxyz[i];
  for (int i = 0; i < 3; i++) T.t[i] = t3[i];
  for (size_t i = 0; i < n; i++) {
    double x, y, z;
    quat_transform(T, (double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2], &x, &y, &z);
    out[3 * i] = (float)x; out[3 * i + 1] = (float)y; out[3 * i + 2] = (float)z;
    valid[i] = projected_valid(xyz[3 * i + 2]) ? 1 : 0;
  }
}

// flags: bit 0 bound, bit 1 in range
void fc_classify(double res, double max_range, const float* origin, const float* xyz, size_t n, uint64_t* key,
                 uint8_t* flags, float* end, uint32_t* len) {
  DevParams P;
  memset(&P, 0, sizeof(P));
  P.res = res;
  P.res_factor = 1.0 / res;
  P.max_range = max_range;
  for (size_t i = 0; i < n; i++) {
    const PointClass c = classify_math(P, origin, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    key[i] = c.key;
    flags[i] = (uint8_t)((c.bound ? 1 : 0) | (c.in_range ? 2 : 0));
    end[3 * i] = c.ex; end[3 * i + 1] = c.ey; end[3 * i + 2] = c.ez;
    len[i] = c.len;
  }
}
}

// updateNode's leaf arithmetic: apply a sequence of hits (1) / misses (0) to one voxel, starting
// from "no node"; returns the final value bits
extern "C" uint32_t fc_update_sequence(const uint8_t* hits, size_t n, float hit, float miss, float cmin, float cmax,
                                       uint32_t* trace) {
  uint32_t bits = kUnknownBits;
  for (size_t i = 0; i < n; i++) {
    bits = update_leaf(bits, hits[i] ? hit : miss, cmin, cmax);
    if (trace) trace[i] = bits;
  }
  return bits;
}
