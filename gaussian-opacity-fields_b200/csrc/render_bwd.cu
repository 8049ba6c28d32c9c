// render_bwd.cu -- K7: per-tile back-to-front gradient pass (backward.cu:634-955).
//
// The reference issues 17 global float atomics per contributing (pixel,Gaussian) pair
// (backward.cu:836,905-912,943-952).  Here the 17 partial gradients are first summed over the warp with
// shuffles, accumulated per (tile,Gaussian) in shared memory, and flushed with at most 17 global atomics
// per (tile,Gaussian) instance.  Pairs are re-evaluated with exactly the forward's operation sequence
// (gof_math.cuh) so that the recomputed alpha equals the forward's; the traversal starts at the last
// Gaussian any pixel of the tile actually blended instead of at the end of the tile list.
#include "gof_common.cuh"
#include "gof_math.cuh"

namespace {

struct BwdArgs {
  int W, H, grid_x;
  float focal_x, focal_y;
  const uint2* ranges;
  const uint32_t* point_list;
  const GofSplat* splat;
  const GofSplatBwd* splat_bwd;
  const float* bg;
  const float* accum;        // [4][tiles*256]
  const uint32_t* ncontrib;  // [2][tiles*256]
  const float* dL_dpix;      // [9][H][W]
  size_t plane;
  float* dL_dmean2D;   // [P,3]
  float* dL_dopacity;  // [P]
  float* dL_dcolor;    // [P,3]
  float* dL_dv2g;      // [P,10]
};

constexpr int BATCH = GOF_BLOCK_SIZE;
constexpr int NGRAD = 17;   // 3 colour + 3 mean2D + 1 opacity + 10 view2gaussian

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__global__ void __launch_bounds__(GOF_BLOCK_SIZE) k_render_backward(const BwdArgs a) {
  __shared__ float4 s_rec[BATCH][4];
  __shared__ float4 s_recb[BATCH][2];
  __shared__ float s_thr[BATCH];
  __shared__ uint32_t s_id[BATCH];
  __shared__ float s_grad[NGRAD][BATCH];
  __shared__ uint32_t s_max;

  const int tile = blockIdx.x;
  const int tile_x = tile % a.grid_x, tile_y = tile / a.grid_x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t pix_x = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
  const uint32_t pix_y = tile_y * 16 + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = pix_x < (uint32_t)a.W && pix_y < (uint32_t)a.H;
  const float rx = gof_ray(pix_x, a.W, a.focal_x);
  const float ry = gof_ray(pix_y, a.H, a.focal_y);

  const uint2 range = a.ranges[tile];
  const size_t slot = (size_t)tile * 256 + threadIdx.x;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)pix_y * a.W + pix_x;

  // backward.cu:692-723
  const float T_final = inside ? a.accum[slot] : 0.f;
  float T = T_final;
  const float final_D = inside ? a.accum[a.plane + slot] : 0.f;
  const float final_A = 1.f - T_final;
  const uint32_t last_contributor = inside ? a.ncontrib[slot] : 0u;
  const uint32_t max_contributor = inside ? a.ncontrib[a.plane + slot] : 0u;
  float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f, dn0 = 0.f, dn1 = 0.f, dn2 = 0.f, ddepth = 0.f, dreg = 0.f;
  if (inside) {
    dpix0 = a.dL_dpix[0 * HW + pid]; dpix1 = a.dL_dpix[1 * HW + pid]; dpix2 = a.dL_dpix[2 * HW + pid];
    dn0 = a.dL_dpix[3 * HW + pid]; dn1 = a.dL_dpix[4 * HW + pid]; dn2 = a.dL_dpix[5 * HW + pid];
    ddepth = a.dL_dpix[6 * HW + pid];
    dreg = a.dL_dpix[8 * HW + pid];   // channel 7 (alpha) receives no gradient, backward.cu:697,717-723
  }
  const float bg_dot_dpixel = a.bg[0] * dpix0 + a.bg[1] * dpix1 + a.bg[2] * dpix2;

  // traversal starts at the deepest Gaussian any pixel of this tile blended
  if (threadIdx.x == 0) s_max = 0u;
  __syncthreads();
  {
    uint32_t m = last_contributor;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) atomicMax(&s_max, m);
  }
  __syncthreads();
  const int used = (int)min(s_max, range.y - range.x);
  const int rounds = (used + BATCH - 1) / BATCH;

  float last_alpha = 0.f;
  float last_c0 = 0.f, last_c1 = 0.f, last_c2 = 0.f, acc_c0 = 0.f, acc_c1 = 0.f, acc_c2 = 0.f;
  float last_n0 = 0.f, last_n1 = 0.f, last_n2 = 0.f, acc_n0 = 0.f, acc_n1 = 0.f, acc_n2 = 0.f;
  const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

  int toDo = used;
  for (int i = 0; i < rounds; ++i, toDo -= BATCH) {
    __syncthreads();
    // stage batch i in REVERSE order: element j of the batch is list entry used-1-(i*BATCH+j)
    const int progress = i * BATCH + (int)threadIdx.x;
    if (progress < used) {
      const uint32_t g = a.point_list[range.x + (uint32_t)(used - 1 - progress)];
      s_id[threadIdx.x] = g;
      const float4* src = reinterpret_cast<const float4*>(a.splat + g);
      const float4 r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2), r3 = __ldg(src + 3);
      s_rec[threadIdx.x][0] = r0; s_rec[threadIdx.x][1] = r1; s_rec[threadIdx.x][2] = r2; s_rec[threadIdx.x][3] = r3;
      const float4* srcb = reinterpret_cast<const float4*>(a.splat_bwd + g);
      s_recb[threadIdx.x][0] = __ldg(srcb); s_recb[threadIdx.x][1] = __ldg(srcb + 1);
      const float op = r2.z;
      s_thr[threadIdx.x] = (op > 0.f) ? (-logf(255.0f * op) - 2e-3f) : __int_as_float(0x7f800000);
    }
#pragma unroll
    for (int k = 0; k < NGRAD; ++k) s_grad[k][threadIdx.x] = 0.f;
    __syncthreads();

    const int nb = toDo < BATCH ? toDo : BATCH;
    for (int j = 0; j < nb; ++j) {
      // zero-based index of this Gaussian in the tile list == the reference's `contributor` after its
      // decrement (backward.cu:763)
      const uint32_t contributor = (uint32_t)(used - 1 - (i * BATCH + j));
      bool contrib = inside && contributor < last_contributor;

      const float4 q0 = s_rec[j][0], q1 = s_rec[j][1], q2 = s_rec[j][2];
      const float v[10] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
      GofPair p;
      float t = 0.f, G = 0.f, alpha = 0.f;
      if (contrib) {
        p = gof_pair_geom(v, rx, ry);
        const float bh = 0.5f * p.BB;
        const float qf = bh * bh * __frcp_rn(p.AA);
        const float pw = -0.5f * (v[9] - qf);
        const float bound = fmaf(fabsf(qf), 3e-7f, pw);
        if (bound < s_thr[j] && fabsf(p.AA) < 1e30f) contrib = false;
      }
      if (contrib) {
        t = gof_pair_t(p);
        if ((double)t <= GOF_NEAR_PLANE_D) contrib = false;
      }
      if (contrib) {
        const float power = gof_pair_power(p, v[9]);
        G = F_EXP(power);
        alpha = fminf(F_MUL(q2.z, G), GOF_ALPHA_MAX);
        if (alpha < GOF_ALPHA_MIN) contrib = false;
      }
      if (!__any_sync(0xffffffffu, contrib)) continue;

      float g[NGRAD];
#pragma unroll
      for (int k = 0; k < NGRAD; ++k) g[k] = 0.f;
      if (contrib) {
        // backward.cu:806-817
        const double td = (double)t;
        const float m = gof_mapped_t(t);
        const float dm_dt = (float)(20.0 / ((99.8 * td) * td));
        const float len = gof_normal_length(p);
        const float nn0 = -p.n0 / len, nn1 = -p.n1 / len, nn2 = -p.n2 / len;
        T = T / (1.f - alpha);
        const float w = alpha * T;
        float dL_dalpha = 0.f;
        // colour, :824-837
        const float4 q3 = s_rec[j][3];
        const float c0 = q2.w, c1 = q3.x, c2 = q3.y;
        acc_c0 = last_alpha * last_c0 + (1.f - last_alpha) * acc_c0; last_c0 = c0;
        acc_c1 = last_alpha * last_c1 + (1.f - last_alpha) * acc_c1; last_c1 = c1;
        acc_c2 = last_alpha * last_c2 + (1.f - last_alpha) * acc_c2; last_c2 = c2;
        dL_dalpha += (c0 - acc_c0) * dpix0;
        dL_dalpha += (c1 - acc_c1) * dpix1;
        dL_dalpha += (c2 - acc_c2) * dpix2;
        g[0] = w * dpix0; g[1] = w * dpix1; g[2] = w * dpix2;
        // distortion: only the depth path survives ("detach weight", :848-858)
        const float dL_dmax_t = 2.0f * (T * alpha) * (m * final_A - final_D) * dreg * dm_dt;
        // normal, :860-877
        acc_n0 = last_alpha * last_n0 + (1.f - last_alpha) * acc_n0; last_n0 = nn0;
        acc_n1 = last_alpha * last_n1 + (1.f - last_alpha) * acc_n1; last_n1 = nn1;
        acc_n2 = last_alpha * last_n2 + (1.f - last_alpha) * acc_n2; last_n2 = nn2;
        dL_dalpha += (nn0 - acc_n0) * dn0;
        dL_dalpha += (nn1 - acc_n1) * dn1;
        dL_dalpha += (nn2 - acc_n2) * dn2;
        const float dnn0 = w * dn0, dnn1 = w * dn1, dnn2 = w * dn2;
        float dL_dlength = dnn0 * p.n0 + dnn1 * p.n1 + dnn2 * p.n2;
        dL_dlength *= 1.f / (len * len);
        float dnrm0 = (-dnn0 + dL_dlength * p.n0) / len;
        float dnrm1 = (-dnn1 + dL_dlength * p.n1) / len;
        float dnrm2 = (-dnn2 + dL_dlength * p.n2) / len;
        // :879-893
        float dL_dt = dL_dmax_t;
        if (contributor == max_contributor - 1u) dL_dt += ddepth;
        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
        // :896-912  2D-mean statistic and opacity
        const float4 b0 = s_recb[j][0], b1 = s_recb[j][1];   // (mx,my,cx,cy) (cz,..)
        const float dx = b0.x - (float)pix_x, dy = b0.y - (float)pix_y;
        const float dL_dG = q2.z * dL_dalpha;
        const float gdx = G * dx, gdy = G * dy;
        const float dG_ddelx = -gdx * b0.z - gdy * b0.w;
        const float dG_ddely = -gdy * b1.x - gdx * b0.w;
        g[3] = dL_dG * dG_ddelx * ddelx_dx;
        g[4] = dL_dG * dG_ddely * ddely_dy;
        g[5] = fabsf(g[3]) + fabsf(g[4]);
        g[6] = G * dL_dalpha;
        // :914-928
        const float dL_dpower = dL_dG * G;
        const float dL_dmin = dL_dpower * -0.5f;
        const double AA = (double)p.AA, BB = (double)p.BB;
        const double boa = BB / AA;
        double dL_dA = (double)dL_dmin * boa * boa / 4.0;
        double dL_dB = (double)dL_dmin * -BB / (2 * AA);
        const float dL_dC = dL_dmin;
        dL_dA += (double)dL_dt * BB / (2 * AA * AA);
        dL_dB += (double)dL_dt * -1.0 / (2 * AA);
        // :938-952
        dnrm0 = (float)(dnrm0 + dL_dA * rx);
        dnrm1 = (float)(dnrm1 + dL_dA * ry);
        dnrm2 = (float)(dnrm2 + dL_dA);
        g[7] = dnrm0 * rx;
        g[8] = dnrm0 * ry + dnrm1 * rx;
        g[9] = dnrm0 + dnrm2 * rx;
        g[10] = dnrm1 * ry;
        g[11] = dnrm1 + dnrm2 * ry;
        g[12] = dnrm2;
        g[13] = (float)(dL_dB * 2 * rx);
        g[14] = (float)(dL_dB * 2 * ry);
        g[15] = (float)(dL_dB * 2);
        g[16] = dL_dC;
      }
#pragma unroll
      for (int k = 0; k < NGRAD; ++k) {
        const float s = warp_sum(g[k]);
        if (lane == k) atomicAdd(&s_grad[k][j], s);
      }
    }
    __syncthreads();
    // flush this batch: thread j owns Gaussian j of the batch
    if ((int)threadIdx.x < nb) {
      const uint32_t gid = s_id[threadIdx.x];
      float g[NGRAD];
      bool any = false;
#pragma unroll
      for (int k = 0; k < NGRAD; ++k) {
        g[k] = s_grad[k][threadIdx.x];
        any |= (g[k] != 0.f);
      }
      if (any) {
        atomicAdd(a.dL_dcolor + 3 * (size_t)gid + 0, g[0]);
        atomicAdd(a.dL_dcolor + 3 * (size_t)gid + 1, g[1]);
        atomicAdd(a.dL_dcolor + 3 * (size_t)gid + 2, g[2]);
        atomicAdd(a.dL_dmean2D + 3 * (size_t)gid + 0, g[3]);
        atomicAdd(a.dL_dmean2D + 3 * (size_t)gid + 1, g[4]);
        atomicAdd(a.dL_dmean2D + 3 * (size_t)gid + 2, g[5]);
        atomicAdd(a.dL_dopacity + gid, g[6]);
#pragma unroll
        for (int k = 0; k < 10; ++k) atomicAdd(a.dL_dv2g + 10 * (size_t)gid + k, g[7 + k]);
      }
    }
  }
}

}  // namespace

int gof_launch_render_backward(const gof_scene_t* s, const GofView& v, const char* geom, const GofGeomLayout& GL,
                               const char* bin, const GofBinLayout& BL, const char* img, const GofImageLayout& IL,
                               const float* dL_dpix, float* dL_dmean2D, float* dL_dopacity, float* dL_dcolor,
                               float* dL_dv2g, cudaStream_t st) {
  BwdArgs a;
  a.W = v.W; a.H = v.H; a.grid_x = v.grid_x; a.focal_x = v.focal_x; a.focal_y = v.focal_y;
  a.ranges = reinterpret_cast<const uint2*>(img + IL.ranges);
  a.point_list = reinterpret_cast<const uint32_t*>(bin + BL.point_list);
  a.splat = reinterpret_cast<const GofSplat*>(geom + GL.splat);
  a.splat_bwd = reinterpret_cast<const GofSplatBwd*>(geom + GL.splat_bwd);
  a.bg = s->background;
  a.accum = reinterpret_cast<const float*>(img + IL.accum);
  a.ncontrib = reinterpret_cast<const uint32_t*>(img + IL.ncontrib);
  a.dL_dpix = dL_dpix;
  a.plane = (size_t)v.tiles * 256;
  a.dL_dmean2D = dL_dmean2D; a.dL_dopacity = dL_dopacity; a.dL_dcolor = dL_dcolor; a.dL_dv2g = dL_dv2g;
  GOF_LAUNCH("render_bwd", st, k_render_backward<<<v.tiles, GOF_BLOCK_SIZE, 0, st>>>(a));
  GOF_LAUNCH_CHECK(s->debug, st);
  return GOF_OK;
}
