// filter3d.cuh -- GaussianModel.compute_3D_filter (scene/gaussian_model.py:262-311) for one point: the smallest camera-space
// depth over the cameras that see the point (depth > 0.2 and projection inside the image enlarged by 15 %).  STAGED
// (SURVEY.md 8(f) rank 4), host/device so that tests/hostmath can run it on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define F3_HD __host__ __device__ __forceinline__
#else
#define F3_HD static inline
#endif

// one camera: R (3x3 as stored by the reference, used as xyz @ R), T, focal_x, focal_y, width, height = 16 floats
#define F3_CAM_FLOATS 16

// returns the minimal valid depth (100000 when no camera sees the point) and sets *seen
F3_HD float f3_min_depth(const float* xyz, const float* cams, int n_cams, bool* seen) {
  float dist = 100000.0f;
  bool any = false;
  for (int c = 0; c < n_cams; ++c) {
    const float* k = cams + (size_t)c * F3_CAM_FLOATS;
    // xyz_cam = xyz @ R + T   (row vector times the stored matrix)
    const float xc = xyz[0] * k[0] + xyz[1] * k[3] + xyz[2] * k[6] + k[9];
    const float yc = xyz[0] * k[1] + xyz[1] * k[4] + xyz[2] * k[7] + k[10];
    const float zc = xyz[0] * k[2] + xyz[1] * k[5] + xyz[2] * k[8] + k[11];
    const bool valid_depth = zc > 0.2f;
    const float z = zc < 0.001f ? 0.001f : zc;                     // torch.clamp(z, min=0.001)
    const float fx = k[12], fy = k[13], W = k[14], H = k[15];
    const float x = xc / z * fx + W / 2.0f, y = yc / z * fy + H / 2.0f;
    const bool in_screen = (x >= -0.15f * W) && (x <= W * 1.15f) && (y >= -0.15f * H) && (y <= 1.15f * H);
    if (valid_depth && in_screen) {
      dist = z < dist ? z : dist;
      any = true;
    }
  }
  *seen = any;
  return dist;
}
