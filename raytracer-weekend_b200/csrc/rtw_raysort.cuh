// Ray reordering between the iterations of the wavefront (rtw_render.cu).
//
// Why.  A closest-hit query (world.hit, hittable/mod.rs:57-69) does not depend on the order in which the rays of an
// iteration are traced, so the order is free.  The wavefront keeps a path in the same pool slot for its whole life:
// after a few bounces slot i and slot i + 1 hold rays in unrelated parts of the scene.  On a hierarchy that does not fit
// the caches (config C5: 11 M primitives, 0.9 GB of node + primitive records against 126 MB of L2) every warp then
// drags its own 32 root-to-leaf paths through HBM: the r02 capture shows 2 kB of DRAM reads per ray at an L2 hit rate
// of 50 %.  Tracing the rays in the order of the scene cell they START in (the point where they enter the scene box)
// makes the ~300 k rays in flight at any time work on one region of the tree: the same records are now hit in L2 / L1.
//
// How.  The shade kernel (and the path start) writes a 32-bit key per slot: Morton code of the entry cell (10 bits per
// axis, truncated to RTW_RAYSORT_POS_BITS), optionally followed by the direction octant; dead slots get the largest key.  Between
// shade and the next traversal an LSD radix sort (8 bits per pass, stable: the rtw_bvh.cu scheme with 32-bit keys and the
// entry count read from the device) turns the iteration's entry list — identity or the live-slot queue — into `order`,
// which the traversal kernel reads instead.  The path state itself never moves: shade still streams the slots coalesced.
#pragma once
#include "rtw_device.cuh"

namespace rtw {

#ifndef RTW_RAYSORT_POS_BITS
#define RTW_RAYSORT_POS_BITS 8   // Morton bits of the entry cell kept in the key (of 30).  r02 A/B on C5 (traverse + sort ms per
                                 // 88 iterations): 8 bits, no octant, ONE pass 379.9 + 6.1; 13 + octant 377.7 + 15.5; 21 + octant
                                 // 368.8 + 24.7; unsorted 446.4 — coarse cells already give the L2 locality, finer keys only pay
                                 // for more passes
#endif
#ifndef RTW_RAYSORT_OCTANT
#define RTW_RAYSORT_OCTANT 0     // append the direction octant (3 bits) to the key
#endif
#define RTW_RAYSORT_KEY_BITS (RTW_RAYSORT_POS_BITS + (RTW_RAYSORT_OCTANT ? 3 : 0))
#define RTW_RAYSORT_PASSES ((RTW_RAYSORT_KEY_BITS + 7) / 8)
#define RTW_RAYSORT_DEAD_KEY 0xFFFFFFFFu
#define RTW_RAYSORT_THREADS 256
#define RTW_RAYSORT_MAX_BLOCKS 1184

struct RaySortGrid {   // per frame: the scene box, scaled to 1024 cells per axis
  float lo[3], hi[3], scale[3];
  uint32_t enabled;
};

struct RaySortDev {
  uint32_t* key;       // [pool] written by shade / path start
  uint32_t* keys[2];   // [pool] ping-pong of the passes
  uint32_t* vals[2];   // [pool]; vals[(RTW_RAYSORT_PASSES - 1) & 1] is the final order
  uint32_t* hist;      // [256][blocks]
  uint32_t* totals;    // [256] per-digit totals of the current pass
};

__device__ __forceinline__ uint32_t expand10(uint32_t x) {  // 10 bits -> every third bit
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

// Not parity relevant (the key only orders the work): approximate reciprocals and explicit FMAs.
__device__ __forceinline__ uint32_t ray_sort_key(const RaySortGrid& g, v3 o, v3 d) {
  float ix, iy, iz;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(d.y));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(d.z));
  const float ax = (g.lo[0] - o.x) * ix, bx = (g.hi[0] - o.x) * ix;
  const float ay = (g.lo[1] - o.y) * iy, by = (g.hi[1] - o.y) * iy;
  const float az = (g.lo[2] - o.z) * iz, bz = (g.hi[2] - o.z) * iz;
  float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));  // NaNs (0 * inf) drop out
  if (!(tn < 3.0e38f)) tn = 0.0f;
  const float px = __fmaf_rn(d.x, tn, o.x), py = __fmaf_rn(d.y, tn, o.y), pz = __fmaf_rn(d.z, tn, o.z);
  const float qx = fminf(fmaxf((px - g.lo[0]) * g.scale[0], 0.0f), 1023.0f);
  const float qy = fminf(fmaxf((py - g.lo[1]) * g.scale[1], 0.0f), 1023.0f);
  const float qz = fminf(fmaxf((pz - g.lo[2]) * g.scale[2], 0.0f), 1023.0f);
  const uint32_t m = (expand10((uint32_t)qx) << 2) | (expand10((uint32_t)qy) << 1) | expand10((uint32_t)qz);
#if RTW_RAYSORT_OCTANT
  const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
  return ((m >> (30 - RTW_RAYSORT_POS_BITS)) << 3) | oct;
#else
  return m >> (30 - RTW_RAYSORT_POS_BITS);
#endif
}

// entries of the iteration: n = *count; entry i is slot list[i] (list == nullptr: slot i)
__device__ __forceinline__ uint32_t raysort_chunk(uint32_t n) {
  const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
  return (per + RTW_RAYSORT_THREADS - 1) & ~(uint32_t)(RTW_RAYSORT_THREADS - 1);
}

// FIRST: keys come from key[slot] through the entry list; later passes read the previous pass's output.
template <bool FIRST>
__global__ void __launch_bounds__(RTW_RAYSORT_THREADS) k_raysort_hist(const uint32_t* __restrict__ count, const uint32_t* __restrict__ qmode,
                                                                     const uint32_t* __restrict__ list, const uint32_t* __restrict__ keys,
                                                                     int shift, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = *count;
  const bool use_list = FIRST && *qmode != 0u;
  const uint32_t chunk = raysort_chunk(n);
  const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
  for (uint32_t i = begin + threadIdx.x; i < end; i += RTW_RAYSORT_THREADS) {
    const uint32_t k = FIRST ? keys[use_list ? list[i] : i] : keys[i];
    atomicAdd(&h[(k >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// Row scan: block d turns hist[d][0..blocks) into its exclusive prefix and writes the digit's total.  The offset of
// the digit itself (exclusive scan of the 256 totals) is added by every scatter block in its prologue.
__global__ void __launch_bounds__(1024) k_raysort_scan(uint32_t blocks, uint32_t* __restrict__ hist, uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sums[32];
  uint32_t* row = hist + (size_t)blockIdx.x * blocks;
  const uint32_t per = (blocks + 1023u) / 1024u;  // 1 or 2
  const uint32_t begin = min(blocks, threadIdx.x * per), end = min(blocks, begin + per);
  uint32_t sum = 0;
  for (uint32_t i = begin; i < end; ++i) sum += row[i];
  uint32_t x = sum;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= (uint32_t)off) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = warp_sums[lane];
    uint32_t ws = w;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
      if (lane >= (uint32_t)off) ws += y;
    }
    warp_sums[lane] = ws - w;
    if (lane == 31) totals[blockIdx.x] = ws;
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + (x - sum);
  for (uint32_t i = begin; i < end; ++i) {
    const uint32_t v = row[i];
    row[i] = run;
    run += v;
  }
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(RTW_RAYSORT_THREADS) k_raysort_scatter(const uint32_t* __restrict__ count, const uint32_t* __restrict__ qmode,
                                                                        const uint32_t* __restrict__ list, const uint32_t* __restrict__ keys_in,
                                                                        const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                                        uint32_t* __restrict__ vals_out, int shift,
                                                                        const uint32_t* __restrict__ hist,
                                                                        const uint32_t* __restrict__ totals) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t warp_cnt[RTW_RAYSORT_THREADS / 32][256];
  __shared__ uint32_t wsum[8];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {  // offset of digit threadIdx.x = exclusive scan of the 256 digit totals (256 threads = 8 warps)
    const uint32_t t = totals[threadIdx.x];
    uint32_t x = t;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= (uint32_t)off) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t pre = 0;
    for (uint32_t w = 0; w < warp; ++w) pre += wsum[w];
    base[threadIdx.x] = pre + (x - t) + hist[threadIdx.x * gridDim.x + blockIdx.x];
  }
  for (uint32_t w = 0; w < RTW_RAYSORT_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = *count;
  const bool use_list = FIRST && *qmode != 0u;
  const uint32_t chunk = raysort_chunk(n);
  const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
  for (uint32_t tile = begin; tile < end; tile += RTW_RAYSORT_THREADS) {
    const uint32_t i = tile + threadIdx.x;
    const bool valid = i < end;
    uint32_t key = 0, val = 0;
    if (valid) {
      if (FIRST) {
        val = use_list ? list[i] : i;
        key = keys_in[val];
      } else {
        key = keys_in[i];
        val = vals_in[i];
      }
    }
    const uint32_t digit = valid ? ((key >> shift) & 255u) : (256u + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[warp][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t pre = 0;
      for (uint32_t w = 0; w < warp; ++w) pre += warp_cnt[w][digit];
      const uint32_t pos = base[digit] + pre + rank;
      if (!LAST) keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    uint32_t tot = 0;
    for (uint32_t w = 0; w < RTW_RAYSORT_THREADS / 32; ++w) {
      tot += warp_cnt[w][threadIdx.x];
      warp_cnt[w][threadIdx.x] = 0;
    }
    base[threadIdx.x] += tot;
    __syncthreads();
  }
}

}  // namespace rtw
