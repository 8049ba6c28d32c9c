// binning_legacy.cu -- the round-1 binning (three launches per radix pass / per scan), kept behind GOF_BINNING=legacy for A/B
// timing and as the cross-check of binning.cu's one-sweep passes (tests/test_gpu_binning.py).  Same results by construction.
//
// Tile binning without the reference's 64-bit global sort.
//
// The reference emits one (tile<<32 | depth_bits, gaussian) pair per (Gaussian,tile) instance and runs
// cub::DeviceRadixSort over 32+log2(tiles) bits of R instances (rasterizer_impl.cu:70-111, 355-363:
// 6 passes x 24 B x R at 1080p).  The order it defines is (tile, depth bits, Gaussian index) because the
// radix sort is stable and instances are emitted in ascending Gaussian index.  We produce the SAME order
// (bit-exact point_list / ranges) with far less traffic by splitting the key LSD-style:
//   1. stable radix sort of the P Gaussians by depth bits (4 x 8-bit passes over P pairs),
//   2. scan tiles_touched in that order, emit instances (tile id, gaussian) in that order,
//   3. stable radix sort of the R instances by tile id only (13 bits at 1080p -> 2 passes, u16 keys),
//   4. tile ranges from the sorted tile ids (rasterizer_impl.cu:149-171).
// All passes are deterministic (no atomics decide an output position).
#include "gof_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// exclusive scan of u32 (3 small kernels: block sums, spine, downsweep). Loader functor lets the scan read
// through a gather (tiles_touched in depth order).
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;   // 2048

struct LoadDirect {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(size_t i) const { return p[i]; }
};
struct LoadGather {
  const uint32_t* src;
  const uint32_t* idx;
  __device__ __forceinline__ uint32_t operator()(size_t i) const { return src[idx[i]]; }
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += n;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  __shared__ uint32_t s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t incl = warp_incl_scan(v);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0u;
    uint32_t wi = warp_incl_scan(w);
    if (lane < SCAN_THREADS / 32) s_warp[lane] = wi - w;
    if (lane == SCAN_THREADS / 32 - 1) s_total = wi;
  }
  __syncthreads();
  const uint32_t r = s_warp[warp] + incl - v;
  *total = s_total;
  __syncthreads();
  return r;
}

template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(Load ld, size_t n, uint32_t* block_sums) {
  const size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += ld(i);
  }
  uint32_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place; grand total to *total_out
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_spine(uint32_t* block_sums, int nblocks, uint32_t* total_out) {
  uint32_t carry = 0;
  for (int base = 0; base < nblocks; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, &total);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// downsweep: thread t owns SCAN_ITEMS consecutive elements
template <class Load, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(Load ld, size_t n, const uint32_t* block_sums,
                                                           uint32_t* out) {
  const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? ld(base + k) : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_excl_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = INCLUSIVE ? run + v[k] : run;
    run += v[k];
  }
}

template <class Load, bool INCLUSIVE>
int scan_u32(Load ld, size_t n, uint32_t* out, uint32_t* tmp, uint32_t* total_out, bool debug, cudaStream_t st) {
  if (n == 0) {
    if (total_out) GOF_CUDA_OK(cudaMemsetAsync(total_out, 0, 4, st));
    return GOF_OK;
  }
  const int nb = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  GOF_LAUNCH("scan", st, k_scan_reduce<Load><<<nb, SCAN_THREADS, 0, st>>>(ld, n, tmp));
  GOF_LAUNCH_CHECK(debug, st);
  GOF_LAUNCH("scan", st, k_scan_spine<<<1, SCAN_THREADS, 0, st>>>(tmp, nb, total_out));
  GOF_LAUNCH_CHECK(debug, st);
  GOF_LAUNCH("scan", st, k_scan_down<Load, INCLUSIVE><<<nb, SCAN_THREADS, 0, st>>>(ld, n, tmp, out));
  GOF_LAUNCH_CHECK(debug, st);
  return GOF_OK;
}

// ------------------------------------------------------------------------------------------------
// One stable LSD radix pass = histogram, row scan, scatter.  Block b owns the key chunk
// [b*CHUNK, (b+1)*CHUNK); warp w of the block owns a contiguous 1/8 of it, processed 32 keys per round,
// so (block, warp, round, lane) order == input order and ranks are stable.

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift, uint32_t mask) {
  return ((uint32_t)k >> shift) & mask;
}

template <typename KeyT>
__global__ void __launch_bounds__(GOF_BLOCK_SIZE) k_radix_hist(const KeyT* __restrict__ keys, size_t n, int shift,
                                                              uint32_t mask, uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t s_h[GOF_RADIX];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * GOF_SORT_CHUNK;
#pragma unroll
  for (int k = 0; k < GOF_SORT_ITEMS; ++k) {
    const size_t i = base + (size_t)k * GOF_BLOCK_SIZE + threadIdx.x;
    if (i < n) atomicAdd(&s_h[digit_of<KeyT>(keys[i], shift, mask)], 1u);
  }
  __syncthreads();
  if (threadIdx.x <= mask) hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];
}

// one block per digit: exclusive scan of that digit's row over blocks; row total -> totals[digit]
__global__ void __launch_bounds__(SCAN_THREADS) k_radix_rowscan(uint32_t* hist, int nblocks, uint32_t* totals) {
  uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
  uint32_t carry = 0;
  for (int base = 0; base < nblocks; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? row[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, &total);
    if (i < nblocks) row[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

template <typename KeyT>
__global__ void __launch_bounds__(GOF_BLOCK_SIZE) k_radix_scatter(const KeyT* __restrict__ keys_in,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 KeyT* __restrict__ keys_out,
                                                                 uint32_t* __restrict__ vals_out, size_t n, int shift,
                                                                 uint32_t mask, const uint32_t* __restrict__ hist,
                                                                 const uint32_t* __restrict__ totals, int nblocks) {
  constexpr int WARPS = GOF_BLOCK_SIZE / 32;
  constexpr int ROUNDS = GOF_SORT_CHUNK / GOF_BLOCK_SIZE;   // 16 rounds of 32 keys per warp
  __shared__ uint32_t s_cnt[WARPS][GOF_RADIX];              // per-warp digit counters -> bases
  __shared__ uint32_t s_dig[GOF_RADIX];                     // global start of each digit
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) s_cnt[w][threadIdx.x] = 0;
  // exclusive scan over digit totals (256 values, one per thread)
  {
    const uint32_t v = threadIdx.x <= mask ? totals[threadIdx.x] : 0u;
    uint32_t total;
    s_dig[threadIdx.x] = block_excl_scan(v, &total);
  }
  __syncthreads();

  const size_t wbase = (size_t)blockIdx.x * GOF_SORT_CHUNK + (size_t)warp * (ROUNDS * 32);
  KeyT key[ROUNDS];
  uint32_t val[ROUNDS];
  uint32_t rank[ROUNDS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + (size_t)r * 32 + lane;
    const bool valid = i < n;
    uint32_t d = GOF_RADIX;   // sentinel digit for the ragged tail
    if (valid) {
      key[r] = keys_in[i];
      val[r] = vals_in[i];
      d = digit_of<KeyT>(key[r], shift, mask);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && valid) {
      old = s_cnt[warp][d];
      s_cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt);
    __syncwarp();   // orders the leaders' counter updates before the next round's reads (racecheck: same warp, different lanes)
  }
  __syncthreads();
  // per digit: exclusive scan over the 8 warps, offset by this block's global base
  if (threadIdx.x <= mask) {
    uint32_t acc = s_dig[threadIdx.x] + hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const uint32_t c = s_cnt[w][threadIdx.x];
      s_cnt[w][threadIdx.x] = acc;
      acc += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const size_t i = wbase + (size_t)r * 32 + lane;
    if (i < n) {
      const uint32_t d = digit_of<KeyT>(key[r], shift, mask);
      const uint32_t pos = s_cnt[warp][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

template <typename KeyT>
int radix_pass(const KeyT* kin, const uint32_t* vin, KeyT* kout, uint32_t* vout, size_t n, int shift, int bits,
               uint32_t* hist, bool debug, cudaStream_t st) {
  const int nb = gof_sort_blocks(n);
  const uint32_t mask = (1u << bits) - 1u;
  uint32_t* totals = hist + (size_t)GOF_RADIX * nb;
  GOF_LAUNCH("radix_hist", st, k_radix_hist<KeyT><<<nb, GOF_BLOCK_SIZE, 0, st>>>(kin, n, shift, mask, hist, nb));
  GOF_LAUNCH_CHECK(debug, st);
  GOF_LAUNCH("radix_rowscan", st, k_radix_rowscan<<<(int)mask + 1, SCAN_THREADS, 0, st>>>(hist, nb, totals));
  GOF_LAUNCH_CHECK(debug, st);
  GOF_LAUNCH("radix_scatter", st, k_radix_scatter<KeyT><<<nb, GOF_BLOCK_SIZE, 0, st>>>(kin, vin, kout, vout, n, shift, mask, hist, totals, nb));
  GOF_LAUNCH_CHECK(debug, st);
  return GOF_OK;
}

// ------------------------------------------------------------------------------------------------
// instance emission, warp-cooperative: a warp owns 32 consecutive depth-ordered Gaussians and writes each
// one's tile list with all 32 lanes (coalesced), instead of one thread looping over all tiles of its
// Gaussian (duplicateWithKeys, rasterizer_impl.cu:70-111).
template <typename KeyT>
__global__ void __launch_bounds__(256) k_emit_instances(int P, const uint32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ incl_offsets,
                                                       const uint2* __restrict__ rect,
                                                       const uint32_t* __restrict__ tiles, int grid_x,
                                                       KeyT* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;   // position in depth order
  uint32_t g = 0, n = 0, off = 0;
  uint2 rc = make_uint2(0u, 0u);
  if (k < P) {
    g = order[k];
    n = tiles[g];
    if (n) {
      rc = rect[g];
      off = incl_offsets[k] - n;
    }
  }
  uint32_t active = __ballot_sync(0xffffffffu, n != 0);
  while (active) {
    const int src = __ffs(active) - 1;
    active &= active - 1;
    const uint32_t gg = __shfl_sync(0xffffffffu, g, src);
    const uint32_t nn = __shfl_sync(0xffffffffu, n, src);
    const uint32_t oo = __shfl_sync(0xffffffffu, off, src);
    const uint32_t r0 = __shfl_sync(0xffffffffu, rc.x, src);
    const uint32_t r1 = __shfl_sync(0xffffffffu, rc.y, src);
    const uint32_t xmin = r0 & 0xffffu, ymin = r0 >> 16;
    const uint32_t w = (r1 & 0xffffu) - xmin;
    for (uint32_t t = lane; t < nn; t += 32) {
      const uint32_t dy = t / w, dx = t - dy * w;
      keys[oo + t] = (KeyT)((ymin + dy) * (uint32_t)grid_x + xmin + dx);
      vals[oo + t] = gg;
    }
  }
}

// rasterizer_impl.cu:149-171 identifyTileRanges on the sorted tile ids (ranges pre-zeroed, :365)
template <typename KeyT>
__global__ void __launch_bounds__(256) k_tile_ranges(size_t L, const KeyT* __restrict__ keys, uint2* __restrict__ ranges) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)keys[idx];
  if (idx == 0)
    ranges[cur].x = 0;
  else {
    const uint32_t prev = (uint32_t)keys[idx - 1];
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == L - 1) ranges[cur].y = (uint32_t)L;
}

template <typename KeyT>
int bin_tiles_t(int P, size_t R, const GofView& v, char* geom, const GofGeomLayout& GL, char* bin,
                const GofBinLayout& BL, char* img, const GofImageLayout& IL, bool debug, cudaStream_t st) {
  uint2* ranges = reinterpret_cast<uint2*>(img + IL.ranges);
  GOF_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)v.tiles * sizeof(uint2), st));
  if (R == 0) return GOF_OK;
  KeyT* ka = reinterpret_cast<KeyT*>(bin + BL.key_a);
  KeyT* kb = reinterpret_cast<KeyT*>(bin + BL.key_b);
  uint32_t* va = reinterpret_cast<uint32_t*>(bin + BL.val_a);
  uint32_t* vb = reinterpret_cast<uint32_t*>(bin + BL.val_b);
  uint32_t* hist = reinterpret_cast<uint32_t*>(bin + BL.hist);
  // the depth sort always runs 4 passes: its result is back in the *_a buffers of the geometry state
  const uint32_t* order = reinterpret_cast<const uint32_t*>(geom + GL.val_a);
  GOF_LAUNCH("emit_instances", st, k_emit_instances<KeyT><<<(P + 255) / 256, 256, 0, st>>>(
      P, order, reinterpret_cast<const uint32_t*>(geom + GL.offsets), reinterpret_cast<const uint2*>(geom + GL.rect),
      reinterpret_cast<const uint32_t*>(geom + GL.tiles), v.grid_x, ka, va));
  GOF_LAUNCH_CHECK(debug, st);
  int shift = 0;
  for (int p = 0; p < BL.passes; ++p) {
    const bool a2b = (p % 2 == 0);
    int rc = radix_pass<KeyT>(a2b ? ka : kb, a2b ? va : vb, a2b ? kb : ka, a2b ? vb : va, R, shift, BL.bits[p], hist,
                              debug, st);
    if (rc != GOF_OK) return rc;
    shift += BL.bits[p];
  }
  const KeyT* sorted = reinterpret_cast<const KeyT*>(bin + BL.sorted_keys);
  GOF_LAUNCH("tile_ranges", st, k_tile_ranges<KeyT><<<(unsigned)((R + 255) / 256), 256, 0, st>>>(R, sorted, ranges));
  GOF_LAUNCH_CHECK(debug, st);
  return GOF_OK;
}

}  // namespace

// stable sort of P (depth bits, gaussian id) pairs, then the inclusive scan of tiles_touched in that order.
// Input keys/values are in key_a/val_a (written by the preprocess kernel); 4 passes -> result back in *_a.
int legacy_gof_depth_sort_and_offsets(int P, char* geom, const GofGeomLayout& L, bool debug, cudaStream_t st) {
  uint32_t* ka = reinterpret_cast<uint32_t*>(geom + L.key_a);
  uint32_t* kb = reinterpret_cast<uint32_t*>(geom + L.key_b);
  uint32_t* va = reinterpret_cast<uint32_t*>(geom + L.val_a);
  uint32_t* vb = reinterpret_cast<uint32_t*>(geom + L.val_b);
  uint32_t* hist = reinterpret_cast<uint32_t*>(geom + L.hist);
  for (int p = 0; p < 4; ++p) {
    const bool a2b = (p % 2 == 0);
    int rc = radix_pass<uint32_t>(a2b ? ka : kb, a2b ? va : vb, a2b ? kb : ka, a2b ? vb : va, (size_t)P, 8 * p, 8, hist,
                                  debug, st);
    if (rc != GOF_OK) return rc;
  }
  LoadGather ld{reinterpret_cast<const uint32_t*>(geom + L.tiles), va};
  return scan_u32<LoadGather, true>(ld, (size_t)P, reinterpret_cast<uint32_t*>(geom + L.offsets),
                                    reinterpret_cast<uint32_t*>(geom + L.scan_tmp),
                                    reinterpret_cast<uint32_t*>(geom + L.total), debug, st);
}

// Stable sort of `n` (tile id, index) pairs by tile id (ids < 2^nbits) for the integrate path's query points, then
// the per-tile ranges of the first ids < num_tiles (ranges must hold num_tiles + 1 uint2; the last slot absorbs the
// sentinel id given to points outside the image).  Buffers: keys/vals ping-pong (u32), hist as in gof_bin_layout.
int legacy_gof_sort_points_by_tile(size_t n, int nbits, uint32_t* ka, uint32_t* kb, uint32_t* va, uint32_t* vb, uint32_t* hist,
                            uint2* ranges, int num_tiles, bool debug, cudaStream_t st, int* result_in_b) {
  GOF_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)(num_tiles + 1) * sizeof(uint2), st));
  *result_in_b = 0;
  if (n == 0) return GOF_OK;
  const int passes = (nbits + 7) / 8;
  int shift = 0, rem = nbits;
  for (int p = 0; p < passes; ++p) {
    const int b = (rem + (passes - p) - 1) / (passes - p);
    const bool a2b = (p % 2 == 0);
    int rc = radix_pass<uint32_t>(a2b ? ka : kb, a2b ? va : vb, a2b ? kb : ka, a2b ? vb : va, n, shift, b, hist, debug, st);
    if (rc != GOF_OK) return rc;
    shift += b; rem -= b;
  }
  *result_in_b = passes % 2;
  const uint32_t* sorted = (passes % 2) ? kb : ka;
  GOF_LAUNCH("tile_ranges", st, k_tile_ranges<uint32_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, sorted, ranges));
  GOF_LAUNCH_CHECK(debug, st);
  return GOF_OK;
}

// Stable LSD radix sort of n (u32 key, u32 value) pairs on the low `nbits` key bits.  Ping-pong buffers a/b (input in a);
// *result_in_b tells where the result is.  hist: GOF_RADIX * (gof_sort_blocks(n) + 1) u32.
int legacy_gof_sort_pairs_u32(uint32_t* ka, uint32_t* kb, uint32_t* va, uint32_t* vb, uint32_t* hist, size_t n, int nbits, bool debug,
                       cudaStream_t st, int* result_in_b) {
  *result_in_b = 0;
  if (n == 0 || nbits <= 0) return GOF_OK;
  const int passes = (nbits + 7) / 8;
  int shift = 0, rem = nbits;
  for (int p = 0; p < passes; ++p) {
    const int b = (rem + (passes - p) - 1) / (passes - p);
    const bool a2b = (p % 2 == 0);
    int rc = radix_pass<uint32_t>(a2b ? ka : kb, a2b ? va : vb, a2b ? kb : ka, a2b ? vb : va, n, shift, b, hist, debug, st);
    if (rc != GOF_OK) return rc;
    shift += b; rem -= b;
  }
  *result_in_b = passes % 2;
  return GOF_OK;
}

// exclusive scan of n u32 (in != out allowed); total (if non-NULL) receives the sum; tmp: n/2048 + 2 u32
int legacy_gof_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t* tmp, uint32_t* total, size_t n, bool debug, cudaStream_t st) {
  LoadDirect ld{in};
  return scan_u32<LoadDirect, false>(ld, n, out, tmp, total, debug, st);
}

int legacy_gof_bin_tiles(int P, size_t R, const GofView& v, char* geom, const GofGeomLayout& GL, char* bin,
                  const GofBinLayout& BL, char* img, const GofImageLayout& IL, bool debug, cudaStream_t st) {
  if (BL.key_bytes == 2) return bin_tiles_t<uint16_t>(P, R, v, geom, GL, bin, BL, img, IL, debug, st);
  return bin_tiles_t<uint32_t>(P, R, v, geom, GL, bin, BL, img, IL, debug, st);
}
