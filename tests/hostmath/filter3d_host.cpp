// tests/hostmath/filter3d_host.cpp -- TEST INFRASTRUCTURE: csrc/filter3d.cuh compiled for the host, with the same
// two-pass structure as the CUDA launcher (per-point minimum, then the "unseen points get the largest seen distance" fill).
#include <math.h>

#include "../../gaussian-opacity-fields_b200/csrc/filter3d.cuh"

extern "C" void hm_filter3d(int P, const float* xyz, int n_cams, const float* cams, float* filter_3D) {
  float focal = 0.f;
  for (int c = 0; c < n_cams; ++c) focal = cams[c * F3_CAM_FLOATS + 12] > focal ? cams[c * F3_CAM_FLOATS + 12] : focal;
  float dmax = -INFINITY;
  for (int i = 0; i < P; ++i) {
    bool seen;
    const float d = f3_min_depth(xyz + 3 * i, cams, n_cams, &seen);
    filter_3D[i] = seen ? d : -1.0f;
    if (seen && d > dmax) dmax = d;
  }
  const float k = sqrtf(0.2f);
  for (int i = 0; i < P; ++i) filter_3D[i] = (filter_3D[i] < 0.f ? dmax : filter_3D[i]) / focal * k;
}
