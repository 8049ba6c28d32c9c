// conv_wgrad.cu -- weight and bias gradient of a 3x3, stride-1, pad-1 convolution with few channels at full image resolution:
// the two convolutions at the tail of the reference's AppearanceNetwork (scene/appearance_network.py:28-29, 16 -> 16 and 16 -> 3
// channels at 1056x1920 in BASELINE config C4).  Why hand-written: profiled on B200 (profiles/r2_appearance_profile_cudnn.txt),
// cuDNN answers these shapes with its generic fp32 `wgrad_alg0_engine` -- 2.1 ms of the 5.2 ms appearance step, more than the
// whole forward of the network.  dW[co][ci][ky][kx] = sum_p gy[co][p] * x[ci][p + (ky-1, kx-1)] is a reduction over two million
// pixels into at most 2304 numbers: here one thread owns a few (co, ci) pairs and keeps their nine sums each in registers while
// persistent CTAs sweep pixel tiles staged in shared memory by cp.async (a 3x3 window of x slides along the row); one atomic flush
// per CTA.  fp32 accumulation (cuDNN's TF32 path rounds the products to 10 bits).
#include "gof_common.cuh"

namespace {

constexpr int TW = 32, TH = 8;   // pixel tile

__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}

// COPT output channels per thread (register tile: 3 + COPT shared loads per 9*COPT FMAs), RG row groups per tile (so that small
// channel pairs still fill a CTA): THREADS = CO/COPT * CI * RG.
template <int CO, int CI, int COPT, int RG>
__global__ void __launch_bounds__((CO / COPT) * CI * RG) k_conv3x3_wgrad(const float* __restrict__ x, const float* __restrict__ gy, int H, int W,
                                                                          int tiles_x, int tiles, float* __restrict__ dW, float* __restrict__ db) {
  constexpr int THREADS = (CO / COPT) * CI * RG;
  constexpr int ROWS = TH / RG;
  static_assert(CO % COPT == 0 && TH % RG == 0 && THREADS >= 64 && THREADS <= 1024, "tile shape");
  constexpr int XP = (TH + 2) * (TW + 2) + 1;   // plane pitch of the x tile (+1: the CI planes fall into distinct banks)
  constexpr int GP = TH * TW + 1;
  __shared__ float s_x[CI * XP];
  __shared__ float s_g[CO * GP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sx_base = (uint32_t)__cvta_generic_to_shared(s_x), sg_base = (uint32_t)__cvta_generic_to_shared(s_g);
  const int ci = tid % CI, cog = (tid / CI) % (CO / COPT), rg = tid / (CI * (CO / COPT));
  float acc[COPT][9], accb[COPT];
#pragma unroll
  for (int j = 0; j < COPT; ++j) {
    accb[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[j][k] = 0.f;
  }
  const size_t HW = (size_t)H * W;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    __syncthreads();
    // Fill with 4-byte cp.async (zero-filled outside the image): every copy of the tile is in flight before the first one is
    // waited for -- a load/store loop serialised ~40 DRAM round trips per thread and tile (first version: 0.59 ms for the
    // 16 -> 3 layer).  One warp per tile row: the row/plane index arithmetic is per row, not per element.
    for (int r = warp; r < CI * (TH + 2); r += THREADS / 32) {
      const int c = r / (TH + 2), yy = r - c * (TH + 2);
      const int iy = y0 + yy - 1;
      const bool rowok = iy >= 0 && iy < H;
      const float* src = x + (size_t)c * HW + (size_t)(rowok ? iy : 0) * W;
      const uint32_t dst = sx_base + 4u * (uint32_t)(c * XP + yy * (TW + 2));
#pragma unroll
      for (int xx = lane; xx < TW + 2; xx += 32) {
        const int ix = x0 + xx - 1;
        const bool ok = rowok && ix >= 0 && ix < W;
        cp_async4_zfill(dst + 4u * (uint32_t)xx, src + (ok ? ix : 0), ok);
      }
    }
    for (int r = warp; r < CO * TH; r += THREADS / 32) {
      const int c = r / TH, yy = r - c * TH;
      const int iy = y0 + yy, ix = x0 + lane;
      const bool ok = iy < H && ix < W;
      cp_async4_zfill(sg_base + 4u * (uint32_t)(c * GP + yy * TW + lane), gy + (size_t)c * HW + (ok ? (size_t)iy * W + ix : 0), ok);
    }
    gof_cp_async_wait_all();
    __syncthreads();
    const float* xs = s_x + ci * XP;
    const float* gs = s_g + cog * COPT * GP;
#pragma unroll 1
    for (int yy = rg * ROWS; yy < rg * ROWS + ROWS; ++yy) {
      const float* r0 = xs + yy * (TW + 2);
      const float* r1 = r0 + (TW + 2);
      const float* r2 = r1 + (TW + 2);
      float a0 = r0[0], a1 = r0[1], b0 = r1[0], b1 = r1[1], c0 = r2[0], c1 = r2[1];
#pragma unroll 8
      for (int xx = 0; xx < TW; ++xx) {
        const float a2 = r0[xx + 2], b2 = r1[xx + 2], c2 = r2[xx + 2];
#pragma unroll
        for (int j = 0; j < COPT; ++j) {
          const float g = gs[j * GP + yy * TW + xx];
          acc[j][0] = fmaf(g, a0, acc[j][0]); acc[j][1] = fmaf(g, a1, acc[j][1]); acc[j][2] = fmaf(g, a2, acc[j][2]);
          acc[j][3] = fmaf(g, b0, acc[j][3]); acc[j][4] = fmaf(g, b1, acc[j][4]); acc[j][5] = fmaf(g, b2, acc[j][5]);
          acc[j][6] = fmaf(g, c0, acc[j][6]); acc[j][7] = fmaf(g, c1, acc[j][7]); acc[j][8] = fmaf(g, c2, acc[j][8]);
          if (ci == 0) accb[j] += g;
        }
        a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < COPT; ++j) {
    const int co = cog * COPT + j;
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(dW + ((size_t)co * CI + ci) * 9 + k, acc[j][k]);
    if (ci == 0 && db != nullptr) atomicAdd(db + co, accb[j]);
  }
}

template <int CO, int CI, int COPT, int RG>
int launch(const float* x, const float* gy, int H, int W, float* dW, float* db, cudaStream_t st) {
  constexpr int THREADS = (CO / COPT) * CI * RG;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH, tiles = tiles_x * tiles_y;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  static int per_sm = 0;   // resident CTAs per SM of this instantiation: the persistent grid is exactly one wave
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv3x3_wgrad<CO, CI, COPT, RG>, THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  }
  const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
  GOF_LAUNCH("conv3x3_wgrad", st, (k_conv3x3_wgrad<CO, CI, COPT, RG><<<grid, THREADS, 0, st>>>(x, gy, H, W, tiles_x, tiles, dW, db)));
  GOF_LAUNCH_CHECK(false, st);
  return GOF_OK;
}

}  // namespace

// x [CI,H,W], gy [CO,H,W] (contiguous fp32, batch 1) -> dW [CO,CI,3,3] and db [CO] (may be NULL), both ACCUMULATED INTO (zero them
// first).  Supported shapes: the channel pairs of the reference's AppearanceNetwork tail and its last upsample block.
extern "C" GOF_API int gof_conv3x3_wgrad(int CO, int CI, int H, int W, const float* x, const float* gy, float* dW, float* db, void* stream) {
  if (H <= 0 || W <= 0 || !x || !gy || !dW) { gof_set_error("conv3x3_wgrad: bad arguments"); return GOF_E_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (CO == 16 && CI == 16) return launch<16, 16, 4, 4>(x, gy, H, W, dW, db, st);
  if (CO == 3 && CI == 16) return launch<3, 16, 3, 8>(x, gy, H, W, dW, db, st);
  if (CO == 16 && CI == 8) return launch<16, 8, 4, 8>(x, gy, H, W, dW, db, st);
  gof_set_error("conv3x3_wgrad: unsupported channel pair %d -> %d", CI, CO);
  return GOF_E_INVALID;
}
