// filter3d.cu -- kernels and C ABI of compute_3D_filter (see filter3d.cuh; scene/gaussian_model.py:262-311).
// The per-point arithmetic is verified on the CPU (tests/test_filter3d_host.py), the wrapper by tests/test_gpu_param_ops.py.
// Pass 1: per point the minimal valid depth over all cameras (one thread per point, camera table read through the
// read-only path) and the maximum of the seen depths (block reduction + atomicMax on the float's bit pattern: depths are
// positive).  Pass 2: unseen points take that maximum; filter = distance / max focal * sqrt(0.2) in the reference's order
// (:306: float division by the focal length, then the float product with 0.2 ** 0.5).  When NO point is seen the reference
// raises (max() of an empty tensor, :301); here *scratch4 stays 0 and the caller must treat that as the error.
#include <math.h>

#include "filter3d.cuh"
#include "gof_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_filter3d_min(int P, const float* __restrict__ xyz, int n_cams, const float* __restrict__ cams,
                                                      float* __restrict__ dist, unsigned int* __restrict__ dmax_bits) {
  __shared__ float s_max[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float mine = 0.0f;                                   // seen depths are > 0.2
  if (i < P) {
    const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    bool seen;
    const float d = f3_min_depth(p, cams, n_cams, &seen);
    dist[i] = seen ? d : -1.0f;
    if (seen) mine = d;
  }
  s_max[threadIdx.x] = mine;
  for (int stride = 128; stride >= 1; stride >>= 1) {
    __syncthreads();
    if ((int)threadIdx.x < stride) s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + stride]);
  }
  __syncthreads();   // the load of s_max[0] below is hoisted in front of the thread test (racecheck: read by all threads)
  if (threadIdx.x == 0 && s_max[0] > 0.0f) atomicMax(dmax_bits, __float_as_uint(s_max[0]));   // positive floats order like uints
}

__global__ void __launch_bounds__(256) k_filter3d_fill(int P, float* __restrict__ dist, const unsigned int* __restrict__ dmax_bits,
                                                       float max_focal) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float d = dist[i];
  dist[i] = __fmul_rn(__fdiv_rn(d < 0.0f ? __uint_as_float(*dmax_bits) : d, max_focal), 0.4472135954999579f);
}

}  // namespace

// xyz [P,3], cams [n_cams,16] = (R 3x3 as stored by the reference, T, focal_x, focal_y, width, height), filter_3D [P] out,
// scratch: 4 bytes; all device pointers except max_focal (host value = largest focal_x of the cameras).
extern "C" GOF_API int gof_compute_3d_filter(int P, const float* xyz, int n_cams, const float* cams, float max_focal, float* filter_3D,
                                             void* scratch4, void* stream) {
  if (P < 0 || n_cams < 0) { gof_set_error("compute_3d_filter: bad sizes"); return GOF_E_INVALID; }
  if (P == 0) return GOF_OK;
  if (!xyz || (n_cams > 0 && !cams) || !filter_3D || !scratch4 || !(max_focal > 0.f)) { gof_set_error("compute_3d_filter: bad arguments"); return GOF_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* dmax = static_cast<unsigned int*>(scratch4);
  GOF_CUDA_OK(cudaMemsetAsync(dmax, 0, 4, st));
  const unsigned blocks = (unsigned)((P + 255) / 256);
  GOF_LAUNCH("filter3d_min", st, k_filter3d_min<<<blocks, 256, 0, st>>>(P, xyz, n_cams, cams, filter_3D, dmax));
  GOF_LAUNCH_CHECK(false, st);
  GOF_LAUNCH("filter3d_fill", st, k_filter3d_fill<<<blocks, 256, 0, st>>>(P, filter_3D, dmax, max_focal));
  GOF_LAUNCH_CHECK(false, st);
  return GOF_OK;
}
