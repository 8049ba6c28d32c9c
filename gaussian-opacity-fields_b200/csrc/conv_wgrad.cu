// conv_wgrad.cu -- weight and bias gradient of a 3x3, stride-1, pad-1 convolution with few channels at full image resolution:
// the two convolutions at the tail of the reference's AppearanceNetwork (scene/appearance_network.py:28-29, 16 -> 16 and 16 -> 3
// channels at 1056x1920 in BASELINE config C4).  Why hand-written: profiled on B200 (profiles/r2_appearance_profile_cudnn.txt),
// cuDNN answers these shapes with its generic fp32 `wgrad_alg0_engine` -- 2.1 ms of the 5.2 ms appearance step, more than the
// whole forward of the network.  dW[co][ci][ky][kx] = sum_p gy[co][p] * x[ci][p + (ky-1, kx-1)] is a reduction over two million
// pixels into at most 2304 numbers: here one thread owns a few (co, ci) pairs and keeps their nine sums each in registers while
// persistent CTAs sweep pixel tiles staged in shared memory (a 3x3 window of x slides along the row); one atomic flush per CTA.  fp32 accumulation (cuDNN's TF32 path rounds the products to 10 bits).
#include "gof_common.cuh"

namespace {

constexpr int TW = 32, TH = 8;   // pixel tile

// COPT output channels per thread (register tile: 3 + COPT shared loads per 9*COPT FMAs), RG row groups per tile (so that small
// channel pairs still fill a CTA): THREADS = CO/COPT * CI * RG.
template <int CO, int CI, int COPT, int RG, bool VEC>
__global__ void __launch_bounds__((CO / COPT) * CI * RG) k_conv3x3_wgrad(const float* __restrict__ x, const float* __restrict__ gy, int H, int W,
                                                                          int tiles_x, int tiles, float* __restrict__ dW, float* __restrict__ db) {
  constexpr int THREADS = (CO / COPT) * CI * RG;
  constexpr int ROWS = TH / RG;
  static_assert(CO % COPT == 0 && TH % RG == 0 && THREADS >= 64 && THREADS <= 1024, "tile shape");
  // Shared-memory pitches chosen for the lanes of a warp (CI channels x two row groups or output-channel groups): x planes
  // 350 floats apart (== -2 mod 32: the CI planes take distinct even banks) with rows 35 apart (odd: the second row group takes the
  // odd banks); g rows 33 apart, planes 265 apart (4 * 265 == 4 mod 32: the output-channel groups take distinct banks).
  constexpr int XROW = TW + 3, XP = (TH + 2) * XROW;
  constexpr int GROW = TW + 1, GP = TH * GROW + 1;
  static_assert(TW == 32 && TH == 8, "pitches are worked out for 32x8 tiles");
  __shared__ float s_x[CI * XP];
  __shared__ float s_g[CO * GP];
  const int tid = threadIdx.x;
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
    if (VEC) {
      // Fill, W % 4 == 0 and 16-byte aligned planes: the 32 interior floats of a tile row are eight aligned LDG.128, the two halo
      // columns scalar loads -- 12 DRAM round trips per thread and tile instead of 42 (ptxas pairs every load with its stores
      // through one register whatever the source order, so the number of loads is what counts; ncu of the scalar version: 23
      // warps per issue waiting on the scoreboard, 0.93 ms for the 16 -> 3 layer).
      constexpr int XG = CI * (TH + 2) * (TW / 4), XH = CI * (TH + 2) * 2, GG = CO * TH * (TW / 4);
#pragma unroll 2
      for (int g = tid; g < XG; g += THREADS) {
        const int c = g / ((TH + 2) * (TW / 4)), r = g - c * ((TH + 2) * (TW / 4));
        const int yy = r / (TW / 4), k = r - yy * (TW / 4);
        const int iy = y0 + yy - 1, ix = x0 + 4 * k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < H && ix < W) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)c * HW + (size_t)iy * W + ix));
        float* d = s_x + c * XP + yy * XROW + 1 + 4 * k;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
#pragma unroll 2
      for (int h = tid; h < XH; h += THREADS) {
        const int c = h / ((TH + 2) * 2), r = h - c * ((TH + 2) * 2);
        const int yy = r >> 1, side = r & 1;
        const int iy = y0 + yy - 1, ix = side ? x0 + TW : x0 - 1;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (size_t)c * HW + (size_t)iy * W + ix);
        s_x[c * XP + yy * XROW + (side ? TW + 1 : 0)] = v;
      }
#pragma unroll 2
      for (int g = tid; g < GG; g += THREADS) {
        const int c = g / (TH * (TW / 4)), r = g - c * (TH * (TW / 4));
        const int yy = r / (TW / 4), k = r - yy * (TW / 4);
        const int iy = y0 + yy, ix = x0 + 4 * k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy < H && ix < W) v = __ldg(reinterpret_cast<const float4*>(gy + (size_t)c * HW + (size_t)iy * W + ix));
        float* d = s_g + c * GP + yy * GROW + 4 * k;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    } else {
      // any W / alignment: scalar loads
      constexpr int NX = CI * (TH + 2) * (TW + 2), NG = CO * TH * TW;
      for (int e = tid; e < NX; e += THREADS) {
        const int c = e / ((TH + 2) * (TW + 2)), r = e - c * ((TH + 2) * (TW + 2));
        const int yy = r / (TW + 2), xx = r - yy * (TW + 2);
        const int iy = y0 + yy - 1, ix = x0 + xx - 1;
        s_x[c * XP + yy * XROW + xx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(x + (size_t)c * HW + (size_t)iy * W + ix) : 0.f;
      }
      for (int e = tid; e < NG; e += THREADS) {
        const int c = e / (TH * TW), r = e - c * (TH * TW);
        const int iy = y0 + r / TW, ix = x0 + (r & (TW - 1));
        s_g[c * GP + (r / TW) * GROW + (r & (TW - 1))] = (iy < H && ix < W) ? __ldg(gy + (size_t)c * HW + (size_t)iy * W + ix) : 0.f;
      }
    }
    __syncthreads();
    const float* xs = s_x + ci * XP;
    const float* gs = s_g + cog * COPT * GP;
#pragma unroll 1
    for (int yy = rg * ROWS; yy < rg * ROWS + ROWS; ++yy) {
      const float* r0 = xs + yy * XROW;
      const float* r1 = r0 + XROW;
      const float* r2 = r1 + XROW;
      float a0 = r0[0], a1 = r0[1], b0 = r1[0], b1 = r1[1], c0 = r2[0], c1 = r2[1];
#pragma unroll 8
      for (int xx = 0; xx < TW; ++xx) {
        const float a2 = r0[xx + 2], b2 = r1[xx + 2], c2 = r2[xx + 2];
#pragma unroll
        for (int j = 0; j < COPT; ++j) {
          const float g = gs[j * GP + yy * GROW + xx];
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

template <int CO, int CI, int COPT, int RG, bool VEC>
int launch_v(const float* x, const float* gy, int H, int W, float* dW, float* db, cudaStream_t st) {
  constexpr int THREADS = (CO / COPT) * CI * RG;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH, tiles = tiles_x * tiles_y;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  static int per_sm = 0;   // resident CTAs per SM of this instantiation: the persistent grid is exactly one wave
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_conv3x3_wgrad<CO, CI, COPT, RG, VEC>, THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  }
  const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
  GOF_LAUNCH("conv3x3_wgrad", st, (k_conv3x3_wgrad<CO, CI, COPT, RG, VEC><<<grid, THREADS, 0, st>>>(x, gy, H, W, tiles_x, tiles, dW, db)));
  GOF_LAUNCH_CHECK(false, st);
  return GOF_OK;
}

template <int CO, int CI, int COPT, int RG>
int launch(const float* x, const float* gy, int H, int W, float* dW, float* db, cudaStream_t st) {
  const bool vec = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy)) & 15u) == 0;
  return vec ? launch_v<CO, CI, COPT, RG, true>(x, gy, H, W, dW, db, st) : launch_v<CO, CI, COPT, RG, false>(x, gy, H, W, dW, db, st);
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
