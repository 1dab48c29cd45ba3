// LBVH build on the GPU: per-primitive world boxes -> scene bounds ->
// 63-bit Morton codes -> LSD radix sort -> Karras 2012 hierarchy -> atomic bottom-up refit.
//
// Node layout in HBM: internal node i owns the 64-byte PAIR nodes[2i], nodes[2i+1] = the records
// of its left and right child (32 bytes each: box + link + meta, see rtw_bvh_node).  One traversal
// step fetches a pair with four LDG.128 and knows both child boxes, so it can descend front to back.
//
// Boxes follow the reference's bounding_box() of every primitive (cited per case) pushed through
// the instance chain the way Translation / YRotation do it (transformations.rs:40-47, 77-111), and
// are then padded by 2^-20 of the scene's largest coordinate: the traversal tests primitives in
// OBJECT space with a transformed ray whose rounding differs from the world-space slab test.
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rtw_scene.cuh"

namespace rtw {

namespace {

// ---- helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  float f;
#ifdef __CUDA_ARCH__
  f = __uint_as_float(b);
#else
  memcpy(&f, &b, 4);
#endif
  return f;
}

struct Box {
  v3 lo, hi;
};

// triangular.rs:79-93
__device__ __forceinline__ void tri_min_max(float a, float b, float c, float& lo, float& hi) {
  lo = fminf(a, fminf(b, c));
  hi = fmaxf(a, fmaxf(b, c));
  if (fabsf(lo - hi) < 0.0002f) { lo = lo - 0.0001f; hi = hi + 0.0001f; }
}

// transformations.rs:77-111
__device__ __forceinline__ Box rotate_box(Box b, float sin_theta, float cos_theta) {
  const float INF = __int_as_float(0x7f800000);
  v3 mn = mk(INF, INF, INF), mx = mk(-INF, -INF, -INF);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      for (int k = 0; k < 2; ++k) {
        float fi = (float)i, fj = (float)j, fk = (float)k;
        float x = fi * b.hi.x + (1.0f - fi) * b.lo.x;
        float y = fj * b.hi.y + (1.0f - fj) * b.lo.y;
        float z = fk * b.hi.z + (1.0f - fk) * b.lo.z;
        float new_x = cos_theta * x + sin_theta * z;
        float new_z = (-sin_theta) * x + cos_theta * z;
        mn.x = fminf(mn.x, new_x); mx.x = fmaxf(mx.x, new_x);
        mn.y = fminf(mn.y, y);     mx.y = fmaxf(mx.y, y);
        mn.z = fminf(mn.z, new_z); mx.z = fmaxf(mx.z, new_z);
      }
  Box r; r.lo = mn; r.hi = mx;
  return r;
}

// Kernel 1: object-space box per the reference, pushed through the instance chain; also rewrites
// the raw parameters into the traversal encoding (canonical order).
//   raw SPHERE  g0=(c, r)
//   raw MSPHERE g0=(c0, r) g1=(c1, time0) g2=(time1)      -> enc g1=(c1-c0, time0) g2=(time1-time0)
//   raw RECT    g0=(a0,a1,b0,b1) g1=(k)
//   raw TRI     g0=(a, b.x) g1=(b.yz, c.xy) g2=(c.z)       -> enc (a, e1, e2, n) in 12 floats
__global__ void k_prim_setup(uint32_t n, const float4* __restrict__ raw, const uint32_t* __restrict__ meta,
                             const uint2* __restrict__ inst_range, const InstOp* __restrict__ inst_ops, float time0,
                             float time1, float4* __restrict__ enc, float4* __restrict__ box_lo,
                             float4* __restrict__ box_hi) {
  uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n) return;
  float4 g0 = raw[3 * (size_t)id], g1 = raw[3 * (size_t)id + 1], g2 = raw[3 * (size_t)id + 2];
  uint32_t m = meta[id];
  uint32_t type = m & 7u, inst = m >> RTW_META_TYPE_BITS;
  Box b;
  switch (type) {
    case PT_SPHERE: {  // spherical.rs:98-104 ; [QUIRK] negative radius gives min > max there: use |r|
      float r = fabsf(g0.w);
      v3 c = mk(g0.x, g0.y, g0.z), rv = mk(r, r, r);
      b.lo = c - rv; b.hi = c + rv;
      break;
    }
    case PT_MSPHERE: {  // spherical.rs:138-150
      v3 c0 = mk(g0.x, g0.y, g0.z), c1 = mk(g1.x, g1.y, g1.z);
      float t0 = g1.w, t1 = g2.x;
      v3 dc = c1 - c0;
      float dt = t1 - t0;
      g1 = make_float4(dc.x, dc.y, dc.z, t0);
      g2 = make_float4(dt, 0.f, 0.f, 0.f);
      v3 sc = c0 + ((time0 - t0) / dt) * dc;
      v3 ec = c0 + ((time1 - t0) / dt) * dc;
      float r = fabsf(g0.w);
      v3 rv = mk(r, r, r);
      b.lo = mk(fminf(sc.x - rv.x, ec.x - rv.x), fminf(sc.y - rv.y, ec.y - rv.y), fminf(sc.z - rv.z, ec.z - rv.z));
      b.hi = mk(fmaxf(sc.x + rv.x, ec.x + rv.x), fmaxf(sc.y + rv.y, ec.y + rv.y), fmaxf(sc.z + rv.z, ec.z + rv.z));
      break;
    }
    case PT_MEDIUM_SPHERE: {  // volumes.rs:80-82 -> spherical.rs:98-104 ; g1.x = -1/density
      float r = fabsf(g0.w);
      v3 c = mk(g0.x, g0.y, g0.z), rv = mk(r, r, r);
      b.lo = c - rv; b.hi = c + rv;
      break;
    }
    case PT_MEDIUM_BOX:  // volumes.rs:80-82 -> rectangular.rs:242-244 ; g0 = (p0, p1.x), g1 = (p1.yz, -1/density)
      b.lo = mk(g0.x, g0.y, g0.z); b.hi = mk(g0.w, g1.x, g1.y);
      break;
    case PT_RECT_YZ:  // rectangular.rs:161-166
      b.lo = mk(g1.x - 0.0001f, g0.x, g0.z); b.hi = mk(g1.x + 0.0001f, g0.y, g0.w);
      break;
    case PT_RECT_XZ:  // rectangular.rs:110-115
      b.lo = mk(g0.x, g1.x - 0.0001f, g0.z); b.hi = mk(g0.y, g1.x + 0.0001f, g0.w);
      break;
    case PT_RECT_XY:  // rectangular.rs:59-64
      b.lo = mk(g0.x, g0.z, g1.x - 0.0001f); b.hi = mk(g0.y, g0.w, g1.x + 0.0001f);
      break;
    default: {  // triangular.rs:140-149 ; encode triangular.rs:101-103
      v3 va = mk(g0.x, g0.y, g0.z), vb = mk(g0.w, g1.x, g1.y), vc = mk(g1.z, g1.w, g2.x);
      tri_min_max(va.x, vb.x, vc.x, b.lo.x, b.hi.x);
      tri_min_max(va.y, vb.y, vc.y, b.lo.y, b.hi.y);
      tri_min_max(va.z, vb.z, vc.z, b.lo.z, b.hi.z);
      v3 e1 = vb - va, e2 = vc - va;
      v3 nn = cross(e1, e2);
      g0 = make_float4(va.x, va.y, va.z, e1.x);
      g1 = make_float4(e1.y, e1.z, e2.x, e2.y);
      g2 = make_float4(e2.z, nn.x, nn.y, nn.z);
      break;
    }
  }
  enc[3 * (size_t)id] = g0; enc[3 * (size_t)id + 1] = g1; enc[3 * (size_t)id + 2] = g2;
  if (inst != 0) {
    uint2 rg = inst_range[inst];
    for (int k = (int)rg.y - 1; k >= 0; --k) {  // innermost wrapper first
      InstOp op = inst_ops[rg.x + k];
      if (op.kind == OP_TRANSLATE) {  // transformations.rs:40-47
        v3 off = mk(op.a, op.b, op.c);
        b.lo = b.lo + off; b.hi = b.hi + off;
      } else {
        b = rotate_box(b, op.a, op.b);
      }
    }
  }
  box_lo[id] = make_float4(b.lo.x, b.lo.y, b.lo.z, 0.f);
  box_hi[id] = make_float4(b.hi.x, b.hi.y, b.hi.z, 0.f);
}

// Kernel 2: bounds of the box centroids (6 ordered uints) and the largest finite |coordinate| (1).
__global__ void k_bounds(uint32_t n, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                         uint32_t* __restrict__ out /*7*/) {
  const float INF = __int_as_float(0x7f800000);
  float mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
  float amax = 0.f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 lo = box_lo[i], hi = box_hi[i];
    float c[3] = {0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z};
    float e[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
    for (int a = 0; a < 3; ++a)
      if (isfinite(c[a])) { mn[a] = fminf(mn[a], c[a]); mx[a] = fmaxf(mx[a], c[a]); }
    for (int a = 0; a < 6; ++a)
      if (isfinite(e[a])) amax = fmaxf(amax, fabsf(e[a]));
  }
  for (int off = 16; off > 0; off >>= 1) {
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  }
  if ((threadIdx.x & 31) == 0) {
    for (int a = 0; a < 3; ++a) {
      atomicMin(&out[a], f2ord(mn[a]));
      atomicMax(&out[3 + a], f2ord(mx[a]));
    }
    atomicMax(&out[6], f2ord(amax));
  }
}

__device__ __forceinline__ uint64_t expand21(uint32_t v) {  // spread 21 bits to every third bit
  uint64_t x = v & 0x1FFFFFu;
  x = (x | (x << 32)) & 0x1F00000000FFFFull;
  x = (x | (x << 16)) & 0x1F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// Kernel 3: 63-bit Morton code of the box centroid inside the centroid bounds.
__global__ void k_morton(uint32_t n, const float4* __restrict__ box_lo, const float4* __restrict__ box_hi,
                         const uint32_t* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 lo = box_lo[i], hi = box_hi[i];
  float c[3] = {0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z};
  uint32_t q[3];
  for (int a = 0; a < 3; ++a) {
    float mn = ord2f(bounds[a]), mx = ord2f(bounds[3 + a]);
    float ext = mx - mn;
    float f = (ext > 0.f && isfinite(c[a])) ? (c[a] - mn) / ext : 0.f;
    f = fminf(fmaxf(f, 0.f), 1.f);
    q[a] = min((uint32_t)(f * 2097152.0f), 2097151u);
  }
  keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
  vals[i] = i;
}

// ---- LSD radix sort, 8 bits per pass, stable ---------------------------------------------------
#define SORT_THREADS 256

__global__ void k_sort_hist(uint32_t n, uint32_t chunk, const uint64_t* __restrict__ keys, int shift,
                            uint32_t* __restrict__ hist /*[256][gridDim.x]*/) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint32_t begin = blockIdx.x * chunk;
  uint32_t end = min(n, begin + chunk);
  for (uint32_t i = begin + threadIdx.x; i < end; i += SORT_THREADS) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist (bin-major) with one block
__global__ void k_sort_scan(uint32_t total, uint32_t* __restrict__ hist) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < total; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = (i < total) ? hist[i] : 0u;
    uint32_t x = v;
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
      if ((threadIdx.x & 31) >= off) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = (threadIdx.x < (blockDim.x >> 5)) ? warp_sums[threadIdx.x] : 0u;
      uint32_t ws = w;
      for (int off = 1; off < 32; off <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
        if (threadIdx.x >= off) ws += y;
      }
      warp_sums[threadIdx.x] = ws - w;  // exclusive
    }
    __syncthreads();
    uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + (x - v);
    if (i < total) hist[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
}

__global__ void k_sort_scatter(uint32_t n, uint32_t chunk, const uint64_t* __restrict__ keys_in,
                               const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
                               uint32_t* __restrict__ vals_out, int shift, const uint32_t* __restrict__ hist) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t warp_cnt[SORT_THREADS / 32][256];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  base[threadIdx.x] = hist[threadIdx.x * gridDim.x + blockIdx.x];
  for (uint32_t w = 0; w < SORT_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
  __syncthreads();
  uint32_t begin = blockIdx.x * chunk;
  uint32_t end = min(n, begin + chunk);
  for (uint32_t tile = begin; tile < end; tile += SORT_THREADS) {
    uint32_t i = tile + threadIdx.x;
    bool valid = i < end;
    uint64_t key = valid ? keys_in[i] : 0ull;
    uint32_t val = valid ? vals_in[i] : 0u;
    uint32_t digit = valid ? ((uint32_t)(key >> shift) & 255u) : (256u + lane);
    uint32_t peers = __match_any_sync(0xffffffffu, digit);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[warp][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t pre = 0;
      for (uint32_t w = 0; w < warp; ++w) pre += warp_cnt[w][digit];
      uint32_t pos = base[digit] + pre + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
    uint32_t tot = 0;
    for (uint32_t w = 0; w < SORT_THREADS / 32; ++w) {
      tot += warp_cnt[w][threadIdx.x];
      warp_cnt[w][threadIdx.x] = 0;
    }
    base[threadIdx.x] += tot;
    __syncthreads();
  }
}

// ---- Karras 2012 ---------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
  return __clzll((long long)(a ^ b));
}

// parent encoding: (pair index << 1) | side.  node_range[i] = (first slot, slot count) of the subtree.
__global__ void k_karras(int n, const uint64_t* __restrict__ keys, uint32_t* __restrict__ node_parent,
                         uint32_t* __restrict__ leaf_parent, uint2* __restrict__ node_range) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  uint32_t me = (uint32_t)i << 1;
  if (lo == gamma) leaf_parent[gamma] = me | 0u; else node_parent[gamma] = me | 0u;
  if (hi == gamma + 1) leaf_parent[gamma + 1] = me | 1u; else node_parent[gamma + 1] = me | 1u;
  if (i == 0) node_parent[0] = 0xFFFFFFFFu;
  node_range[i] = make_uint2((uint32_t)lo, (uint32_t)(hi - lo + 1));
}

// Kernel: compact pairs for hierarchies that do not fit the caches.  Beyond L2 a child-pair fetch is a random
// 64-byte gather, and HBM serves those at ~1.3 TB/s whatever the parallelism (tools/gather_peak.cu): the only way to
// go faster is to move fewer bytes.  A compact pair is 32 bytes: both child boxes as 16-bit coordinates on a uniform
// grid over the scene's root box (lo rounded down, hi rounded up, checked with the traversal's own dequantisation
// expression, so the compact box always CONTAINS the exact one: culling stays conservative), plus one word per
// child: an internal link (pair index) or 0x80000000 | (count - 1) << 26 | first slot for a leaf.
struct QuantGrid {
  float lo[3], step[3];
};
__device__ __forceinline__ uint32_t quant_down(float x, float lo, float step) {
  float q = floorf((x - lo) / step);
  q = fminf(fmaxf(q, 0.0f), 65535.0f);
  while (q > 0.0f && __fmaf_rn(q, step, lo) > x) q -= 1.0f;
  return (uint32_t)q;
}
__device__ __forceinline__ uint32_t quant_up(float x, float lo, float step) {
  float q = ceilf((x - lo) / step);
  q = fminf(fmaxf(q, 0.0f), 65535.0f);
  while (q < 65535.0f && __fmaf_rn(q, step, lo) < x) q += 1.0f;
  return (uint32_t)q;
}
__global__ void k_build_compact(uint32_t num_nodes, const float4* __restrict__ nodes, QuantGrid g, uint4* __restrict__ out,
                                uint32_t* __restrict__ bad) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_nodes) return;
  for (int c = 0; c < 2; ++c) {
    const float4 lo = nodes[4 * (size_t)i + 2 * c], hi = nodes[4 * (size_t)i + 2 * c + 1];
    const uint32_t lx = quant_down(lo.x, g.lo[0], g.step[0]), ly = quant_down(lo.y, g.lo[1], g.step[1]),
                   lz = quant_down(lo.z, g.lo[2], g.step[2]);
    const uint32_t hx = quant_up(hi.x, g.lo[0], g.step[0]), hy = quant_up(hi.y, g.lo[1], g.step[1]),
                   hz = quant_up(hi.z, g.lo[2], g.step[2]);
    // containment must hold exactly (a coordinate outside the grid, a NaN box): otherwise the scene keeps the fp32 pairs
    if (!(__fmaf_rn((float)lx, g.step[0], g.lo[0]) <= lo.x && __fmaf_rn((float)ly, g.step[1], g.lo[1]) <= lo.y &&
          __fmaf_rn((float)lz, g.step[2], g.lo[2]) <= lo.z && __fmaf_rn((float)hx, g.step[0], g.lo[0]) >= hi.x &&
          __fmaf_rn((float)hy, g.step[1], g.lo[1]) >= hi.y && __fmaf_rn((float)hz, g.step[2], g.lo[2]) >= hi.z))
      atomicExch(bad, 1u);
    const int32_t link = __float_as_int(lo.w);
    const uint32_t meta = __float_as_uint(hi.w);
    uint32_t w;
    if (link >= 0) {
      w = (uint32_t)link;
    } else {
      const uint32_t first = (uint32_t)(~link);
      if (link == (int32_t)0x80000000 || first >= (1u << 26) || meta == 0u || meta > 32u) atomicExch(bad, 1u);
      w = 0x80000000u | ((meta - 1u) << 26) | (first & 0x3FFFFFFu);
    }
    out[2 * (size_t)i + c] = make_uint4(lx | (ly << 16), lz | (hx << 16), hy | (hz << 16), w);
  }
}

// Kernel: the 4-wide view of the tree.  Record i holds the children of pair i's two children (a child that is a
// leaf stands for itself), i.e. the four boxes a ray meets two levels below pair i, in one 128-byte record: a
// traversal step through it replaces two dependent pair fetches by one.  Unused entries: link = RTW_LINK_DONE
// and an empty box.  Only pairs at even depth are ever visited through this view; building it for every pair
// keeps the indices identical to the binary tree's.
__global__ void k_build_wide(uint32_t num_nodes, const float4* __restrict__ nodes, float4* __restrict__ nodes4) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_nodes) return;
  const float INF = __int_as_float(0x7f800000);
  float4 out[8];
  int n = 0;
  for (int c = 0; c < 2; ++c) {
    const float4 r0 = nodes[4 * (size_t)i + 2 * c], r1 = nodes[4 * (size_t)i + 2 * c + 1];
    const int32_t link = __float_as_int(r0.w);
    if (link >= 0) {
      const float4* q = nodes + 4 * (size_t)link;
      out[n++] = q[0]; out[n++] = q[1]; out[n++] = q[2]; out[n++] = q[3];
    } else {
      out[n++] = r0; out[n++] = r1;
    }
  }
  for (; n < 8; n += 2) {
    out[n] = make_float4(INF, INF, INF, __int_as_float((int32_t)0x80000000));
    out[n + 1] = make_float4(-INF, -INF, -INF, 0.f);
  }
  for (int k = 0; k < 8; ++k) nodes4[8 * (size_t)i + k] = out[k];
}

#if RTW_TOP_TREE > 0
// Kernel (one thread): the first `cap` pairs of the tree in breadth-first order, for the shared-memory top of
// tree of the traversal kernels.  A child link that points to a pair inside the copy is re-targeted to its index
// in the copy and tagged RTW_LINK_TOP; every other link (deeper pairs, leaves) is kept.
__global__ void k_top_tree(const float4* __restrict__ nodes, uint32_t num_nodes, uint32_t cap, float4* __restrict__ top,
                           uint32_t* __restrict__ top_src, uint32_t* __restrict__ top_count) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t n = 0;
  if (num_nodes > 0 && cap > 0) { top_src[0] = 0; n = 1; }
  for (uint32_t t = 0; t < n; ++t) {
    const float4* src = nodes + 4 * (size_t)top_src[t];
    float4 rec[4] = {src[0], src[1], src[2], src[3]};
    for (int c = 0; c < 2; ++c) {
      const int32_t link = __float_as_int(rec[2 * c].w);
      if (link >= 0 && n < cap) {
        top_src[n] = (uint32_t)link;
        rec[2 * c].w = __int_as_float((int32_t)(RTW_LINK_TOP | n));
        n++;
      }
    }
    for (int k = 0; k < 4; ++k) top[4 * (size_t)t + k] = rec[k];
  }
  *top_count = n;
}
#endif

// Kernel: leaves.  Slot s holds primitive vals[s]; copies its geometry and meta into slot order.
__global__ void k_emit_leaves(uint32_t n, const uint32_t* __restrict__ vals, const float4* __restrict__ enc,
                              const uint32_t* __restrict__ meta, const uint32_t* __restrict__ prim_mat,
                              const int32_t* __restrict__ prim_shade, float4* __restrict__ geom,
                              int32_t* __restrict__ slot_prim, uint32_t* __restrict__ slot_meta, int2* __restrict__ slot_ms) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t id = vals[s];
  slot_prim[s] = (int32_t)id;
  slot_meta[s] = meta[id];
  slot_ms[s] = make_int2((int)prim_mat[id], prim_shade[id]);
  geom[3 * (size_t)s] = enc[3 * (size_t)id];
  geom[3 * (size_t)s + 1] = enc[3 * (size_t)id + 1];
  geom[3 * (size_t)s + 2] = enc[3 * (size_t)id + 2];
}

__device__ __forceinline__ void store_record(float4* nodes, uint32_t rec, v3 lo, v3 hi, int32_t link, uint32_t meta) {
  __stcg(&nodes[2 * (size_t)rec], make_float4(lo.x, lo.y, lo.z, __int_as_float(link)));
  __stcg(&nodes[2 * (size_t)rec + 1], make_float4(hi.x, hi.y, hi.z, __uint_as_float(meta)));
}

__device__ __forceinline__ float half_area(v3 lo, v3 hi) {
  float dx = fmaxf(hi.x - lo.x, 0.f), dy = fmaxf(hi.y - lo.y, 0.f), dz = fmaxf(hi.z - lo.z, 0.f);
  return dx * dy + dy * dz + dz * dx;
}

// Kernel: bottom-up refit + SAH leaf collapse.  One thread per Karras leaf; the second thread to
// reach a pair owns it (atomic flag), unions the child boxes (Aabb::surrounding_box, aabb.rs:74-88)
// and decides whether the subtree stays a hierarchy or becomes ONE leaf over its contiguous slot
// range:  C_leaf = n * C_PRIM   vs   C_split = C_PAIR + (SA(l) C(l) + SA(r) C(r)) / SA(node).
// Leaves are ranges: link = ~first_slot, meta = primitive count.  Scenes whose boxes all overlap
// (a Cornell box: every wall spans the room) collapse into a single converged primitive loop.
struct RefitParams {
  float pad;        // box padding (see file header)
  float c_pair;     // cost of fetching + testing one child pair
  float c_prim;     // cost of one primitive test
  uint32_t max_leaf;
  uint32_t force_flat;  // collapse the root unconditionally (tiny scene)
};

__global__ void k_refit(uint32_t n, const uint32_t* __restrict__ vals, const float4* __restrict__ box_lo,
                        const float4* __restrict__ box_hi, RefitParams rp, const uint32_t* __restrict__ node_parent,
                        const uint32_t* __restrict__ leaf_parent, const uint2* __restrict__ node_range,
                        uint32_t* __restrict__ flags, uint32_t* __restrict__ heights, float* __restrict__ node_cost,
                        uint32_t* __restrict__ collapsed, float4* __restrict__ nodes, float* __restrict__ out_root,
                        uint32_t* __restrict__ out_info) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t id = vals[s];
  float4 l4 = box_lo[id], h4 = box_hi[id];
  v3 lo = mk(l4.x - rp.pad, l4.y - rp.pad, l4.z - rp.pad), hi = mk(h4.x + rp.pad, h4.y + rp.pad, h4.z + rp.pad);
  const float INF = __int_as_float(0x7f800000);
  if (n == 1) {
    store_record(nodes, 0, lo, hi, ~0, 1u);
    store_record(nodes, 1, mk(INF, INF, INF), mk(-INF, -INF, -INF), ~0, 0u);
    out_root[0] = lo.x; out_root[1] = lo.y; out_root[2] = lo.z;
    out_root[3] = hi.x; out_root[4] = hi.y; out_root[5] = hi.z;
    out_info[0] = 1; out_info[1] = 1;
    return;
  }
  uint32_t par = leaf_parent[s];
  store_record(nodes, par, lo, hi, ~(int32_t)s, 1u);
  uint32_t h = 1;  // height of the subtree rooted at the parent pair, counted in pairs
  for (;;) {
    uint32_t p = par >> 1;
    atomicMax(&heights[p], h);
    __threadfence();
    if (atomicAdd(&flags[p], 1u) == 0u) return;  // first to arrive: the sibling will finish the pair
    __threadfence();
    float4 a0 = __ldcg(&nodes[4 * (size_t)p]), a1 = __ldcg(&nodes[4 * (size_t)p + 1]);
    float4 b0 = __ldcg(&nodes[4 * (size_t)p + 2]), b1 = __ldcg(&nodes[4 * (size_t)p + 3]);
    v3 ulo = mk(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z));
    v3 uhi = mk(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z));
    // SAH: cost of the children as they stand now (leaf range or hierarchy)
    int32_t la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
    float ca = la < 0 ? (float)__float_as_uint(a1.w) * rp.c_prim : __ldcg(&node_cost[la]);
    float cb = lb < 0 ? (float)__float_as_uint(b1.w) * rp.c_prim : __ldcg(&node_cost[lb]);
    float sa = half_area(ulo, uhi);
    float c_split = rp.c_pair + (sa > 0.f ? (half_area(mk(a0.x, a0.y, a0.z), mk(a1.x, a1.y, a1.z)) * ca +
                                              half_area(mk(b0.x, b0.y, b0.z), mk(b1.x, b1.y, b1.z)) * cb) / sa
                                           : ca + cb);
    uint2 rg = node_range[p];
    float c_leaf = (float)rg.y * rp.c_prim;
    bool collapse = rg.y <= rp.max_leaf && !(c_leaf > c_split);
    if (p == 0 && rp.force_flat) collapse = true;
    node_cost[p] = collapse ? c_leaf : c_split;
    collapsed[p] = collapse ? 1u : 0u;
    h = atomicMax(&heights[p], 0u);
    if (p == 0) {
      if (collapse) {  // the whole scene is one leaf: pair 0 = {that leaf, empty}
        store_record(nodes, 0, ulo, uhi, ~(int32_t)rg.x, rg.y);
        store_record(nodes, 1, mk(INF, INF, INF), mk(-INF, -INF, -INF), ~0, 0u);
      }
      out_root[0] = ulo.x; out_root[1] = ulo.y; out_root[2] = ulo.z;
      out_root[3] = uhi.x; out_root[4] = uhi.y; out_root[5] = uhi.z;
      out_info[0] = h; out_info[1] = collapse ? 1u : 0u;
      return;
    }
    par = node_parent[p];
    if (collapse) store_record(nodes, par, ulo, uhi, ~(int32_t)rg.x, rg.y);
    else store_record(nodes, par, ulo, uhi, (int32_t)p, 0u);
    h = h + 1;
  }
}

// Kernel: inside every final multi-primitive leaf, order the primitives by (instance, type) — the
// traversal re-transforms / re-permutes the ray only when that key changes.  A leaf is final when
// its pair collapsed and its parent did not.  Stable insertion sort of <= max_leaf entries of `vals`.
__global__ void k_sort_leaf_ranges(uint32_t n_internal, const uint32_t* __restrict__ collapsed,
                                   const uint32_t* __restrict__ node_parent, const uint2* __restrict__ node_range,
                                   const uint32_t* __restrict__ meta, uint32_t* __restrict__ vals) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_internal || !collapsed[p]) return;
  // final = no ancestor collapsed.  (Checking only the parent is not enough: a grandparent may collapse
  // although the parent did not, and two threads would then sort overlapping ranges concurrently.)
  for (uint32_t q = p; q != 0;) {
    q = node_parent[q] >> 1;
    if (collapsed[q]) return;
  }
  uint2 rg = node_range[p];
  for (uint32_t i = 1; i < rg.y; ++i) {
    uint32_t v = vals[rg.x + i];
    uint32_t key = meta[v];
    uint32_t j = i;
    while (j > 0 && meta[vals[rg.x + j - 1]] > key) {
      vals[rg.x + j] = vals[rg.x + j - 1];
      --j;
    }
    vals[rg.x + j] = v;
  }
}

// ---- tree rotations (r02) ---------------------------------------------------------------------------------------------
// A Karras tree splits where the Morton prefix changes: spatial medians.  Measured offline on the trees this file builds
// (tools/dump_bvh.py + tools/sah_study.py): the expected number of pair visits, sum of SA(internal node) / SA(root), is
// 1.42x (cow) / 1.54x (monument) / 2.0x (jumpy-balls: one huge sphere next to 485 small ones) that of a full-sweep SAH
// tree over the same leaves — and ONE bottom-up pass of tree rotations (Kensler 2008) closes almost all of that gap
// (cow 3.93 -> 2.92 vs 2.78 for the sweep build; monument 14.2 -> 9.9 -> 9.1 vs 9.2; jumpy 2.0 -> 1.0).
// The pass has the shape of k_refit: one thread per leaf record climbs, the second thread to arrive at a pair owns it, so
// the whole subtree below is final.  At pair p with children A, B it looks at the (up to) four grandchildren and applies
// the best of: B <-> A1, B <-> A2, A <-> B1, A <-> B2, A1 <-> B1, A1 <-> B2 — whichever lowers the surface area of the
// nodes below p most (p's own box does not change).  Only child RECORDS move between pairs; pair indices, leaves (slot
// ranges) and everything the traversal relies on stay as they are, and closest hits do not depend on the tree's shape.
// A rotation may make p deeper; the stack of the walk is RTW_STACK_SIZE entries, so a candidate is refused when it would
// lift p above max(its current height, kRotateHeightCap), and the build falls back to the unrotated tree if the root still
// ends up too deep.
constexpr uint32_t kRotateHeightCap = 56;

// live[p]: pair p is part of the final tree (neither it nor an ancestor collapsed into a leaf)
__global__ void k_rotate_live(uint32_t n_internal, const uint32_t* __restrict__ collapsed, const uint32_t* __restrict__ node_parent,
                              uint32_t* __restrict__ live) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_internal) return;
  uint32_t ok = collapsed[p] ? 0u : 1u;
  for (uint32_t q = p; ok && q != 0;) {
    q = node_parent[q] >> 1;
    if (collapsed[q]) ok = 0u;
  }
  live[p] = ok;
}

struct ChildRec {
  float4 lo, hi;  // lo.w = link, hi.w = meta
};
__device__ __forceinline__ ChildRec load_rec(const float4* nodes, uint32_t rec) {
  ChildRec r;
  r.lo = __ldcg(&nodes[2 * (size_t)rec]);
  r.hi = __ldcg(&nodes[2 * (size_t)rec + 1]);
  return r;
}
__device__ __forceinline__ void put_rec(float4* nodes, uint32_t* node_parent, uint32_t pair, uint32_t side, const ChildRec& r) {
  __stcg(&nodes[2 * (size_t)(2 * pair + side)], r.lo);
  __stcg(&nodes[2 * (size_t)(2 * pair + side) + 1], r.hi);
  const int32_t link = __float_as_int(r.lo.w);
  if (link >= 0) node_parent[link] = (pair << 1) | side;
}
__device__ __forceinline__ ChildRec union_rec(const ChildRec& a, const ChildRec& b, int32_t link) {
  ChildRec r;
  r.lo = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), __int_as_float(link));
  r.hi = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), __uint_as_float(0u));
  return r;
}
__device__ __forceinline__ float rec_area(const ChildRec& r) {
  return half_area(mk(r.lo.x, r.lo.y, r.lo.z), mk(r.hi.x, r.hi.y, r.hi.z));
}
__device__ __forceinline__ uint32_t rec_height(const ChildRec& r, const uint32_t* heights) {
  const int32_t link = __float_as_int(r.lo.w);
  return link >= 0 ? __ldcg(&heights[link]) : 0u;
}

// start[t]: record t = (pair t / 2, side t % 2) is a leaf record of a live pair NOW — taken before the pass, because the pass
// itself moves leaf records between pairs
__global__ void k_rotate_starts(uint32_t n_internal, const uint32_t* __restrict__ live, const float4* __restrict__ nodes,
                                uint8_t* __restrict__ start) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n_internal) return;
  start[t] = (live[t >> 1] && __float_as_int(nodes[2 * (size_t)t].w) < 0) ? 1 : 0;
}

__global__ void k_rotate(uint32_t n_internal, const uint8_t* __restrict__ start, uint32_t* __restrict__ node_parent,
                         uint32_t* __restrict__ flags, uint32_t* __restrict__ heights, float4* __restrict__ nodes,
                         uint32_t* __restrict__ out_info /* [0] root height, [2] rotations applied */) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t p = t >> 1;
  if (p >= n_internal || !start[t]) return;
  uint32_t applied = 0;
  for (;;) {
    __threadfence();
    if (atomicAdd(&flags[p], 1u) == 0u) break;  // first to arrive: the sibling subtree's thread will own the pair
    __threadfence();
    ChildRec A = load_rec(nodes, 2 * p), B = load_rec(nodes, 2 * p + 1);
    const int32_t la = __float_as_int(A.lo.w), lb = __float_as_int(B.lo.w);
    ChildRec A1 = A, A2 = A, B1 = B, B2 = B;
    uint32_t hA1 = 0, hA2 = 0, hB1 = 0, hB2 = 0;
    if (la >= 0) { A1 = load_rec(nodes, 2 * (uint32_t)la); A2 = load_rec(nodes, 2 * (uint32_t)la + 1); hA1 = rec_height(A1, heights); hA2 = rec_height(A2, heights); }
    if (lb >= 0) { B1 = load_rec(nodes, 2 * (uint32_t)lb); B2 = load_rec(nodes, 2 * (uint32_t)lb + 1); hB1 = rec_height(B1, heights); hB2 = rec_height(B2, heights); }
    const uint32_t hA = la >= 0 ? 1u + max(hA1, hA2) : 0u, hB = lb >= 0 ? 1u + max(hB1, hB2) : 0u;
    const uint32_t h_cur = 1u + max(hA, hB);
    const uint32_t h_max = max(h_cur, kRotateHeightCap);
    const float sA = rec_area(A), sB = rec_area(B);
    float best = -1e-6f * (sA + sB);  // a rotation must win by more than rounding noise
    int which = 0;
    uint32_t h_new = h_cur;
    auto consider = [&](int id, float delta, uint32_t h) {
      if (delta < best && h <= h_max) { best = delta; which = id; h_new = h; }
    };
    if (la >= 0) {
      consider(1, rec_area(union_rec(B, A2, 0)) - sA, 1u + max(hA1, 1u + max(hB, hA2)));  // B <-> A1
      consider(2, rec_area(union_rec(A1, B, 0)) - sA, 1u + max(hA2, 1u + max(hA1, hB)));  // B <-> A2
    }
    if (lb >= 0) {
      consider(3, rec_area(union_rec(A, B2, 0)) - sB, 1u + max(hB1, 1u + max(hA, hB2)));  // A <-> B1
      consider(4, rec_area(union_rec(B1, A, 0)) - sB, 1u + max(hB2, 1u + max(hB1, hA)));  // A <-> B2
    }
    if (la >= 0 && lb >= 0) {
      consider(5, rec_area(union_rec(B1, A2, 0)) + rec_area(union_rec(A1, B2, 0)) - sA - sB,
               2u + max(max(hB1, hA2), max(hA1, hB2)));  // A1 <-> B1
      consider(6, rec_area(union_rec(B2, A2, 0)) + rec_area(union_rec(B1, A1, 0)) - sA - sB,
               2u + max(max(hB2, hA2), max(hB1, hA1)));  // A1 <-> B2
    }
    const uint32_t ua = (uint32_t)la, ub = (uint32_t)lb;
    switch (which) {
      case 1: put_rec(nodes, node_parent, ua, 0, B); put_rec(nodes, node_parent, p, 0, union_rec(B, A2, la)); put_rec(nodes, node_parent, p, 1, A1); break;
      case 2: put_rec(nodes, node_parent, ua, 1, B); put_rec(nodes, node_parent, p, 0, union_rec(A1, B, la)); put_rec(nodes, node_parent, p, 1, A2); break;
      case 3: put_rec(nodes, node_parent, ub, 0, A); put_rec(nodes, node_parent, p, 1, union_rec(A, B2, lb)); put_rec(nodes, node_parent, p, 0, B1); break;
      case 4: put_rec(nodes, node_parent, ub, 1, A); put_rec(nodes, node_parent, p, 1, union_rec(B1, A, lb)); put_rec(nodes, node_parent, p, 0, B2); break;
      case 5:
        put_rec(nodes, node_parent, ua, 0, B1); put_rec(nodes, node_parent, ub, 0, A1);
        put_rec(nodes, node_parent, p, 0, union_rec(B1, A2, la)); put_rec(nodes, node_parent, p, 1, union_rec(A1, B2, lb));
        break;
      case 6:
        put_rec(nodes, node_parent, ua, 0, B2); put_rec(nodes, node_parent, ub, 1, A1);
        put_rec(nodes, node_parent, p, 0, union_rec(B2, A2, la)); put_rec(nodes, node_parent, p, 1, union_rec(B1, A1, lb));
        break;
      default: break;
    }
    if (which) {
      applied++;
      // the heights of the rewritten children (their pairs keep their indices)
      if (which == 1) heights[ua] = 1u + max(hB, hA2);
      if (which == 2) heights[ua] = 1u + max(hA1, hB);
      if (which == 3) heights[ub] = 1u + max(hA, hB2);
      if (which == 4) heights[ub] = 1u + max(hB1, hA);
      if (which == 5) { heights[ua] = 1u + max(hB1, hA2); heights[ub] = 1u + max(hA1, hB2); }
      if (which == 6) { heights[ua] = 1u + max(hB2, hA2); heights[ub] = 1u + max(hB1, hA1); }
    }
    heights[p] = h_new;
    if (p == 0) {
      out_info[0] = h_new;
      break;
    }
    p = node_parent[p] >> 1;
  }
  if (applied) atomicAdd(&out_info[2], applied);
}

template <class T>
int dev_alloc(rtw_scene* s, T** out, size_t count) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = dev_malloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(RTW_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
  }
  s->allocations.push_back(p);
  s->allocation_bytes.push_back(bytes);
  s->device_bytes += bytes;
  *out = (T*)p;
  return RTW_OK;
}
template <class P, class T>
int dev_upload(rtw_scene* s, P* out, const std::vector<T>& v) {
  T* p = nullptr;
  int rc = dev_alloc(s, &p, v.size());
  if (rc) return rc;
  if (!v.empty()) RTW_CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = p;
  return RTW_OK;
}

}  // namespace

void free_scene_device(rtw_scene* s) {
  for (void* p : s->allocations) mem_free(p);
  s->allocations.clear();
  s->allocation_bytes.clear();
  s->device_bytes = 0;
}

int build_scene_device(rtw_scene* s, float time0, float time1, rtw_build_stats* stats) {
  const uint32_t n = (uint32_t)s->prim_meta.size();
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  cudaEvent_t ev[3];
  for (auto& e : ev) RTW_CUDA_TRY(cudaEventCreate(&e));
  RTW_CUDA_TRY(cudaEventRecord(ev[0]));

  // ---- upload -----------------------------------------------------------------------------------
  SceneDev& d = s->dev;
  int rc;
  if ((rc = dev_upload(s, &d.prim_meta, s->prim_meta))) return rc;
  if ((rc = dev_upload(s, &d.prim_mat, s->prim_mat))) return rc;
  if ((rc = dev_upload(s, &d.prim_shade, s->prim_shade))) return rc;
  if ((rc = dev_upload(s, &d.tri_shade, s->tri_shade))) return rc;
  if ((rc = dev_upload(s, &d.inst_range, s->inst_range))) return rc;
  if ((rc = dev_upload(s, &d.inst_ops, s->inst_ops))) return rc;
  {  // SolidColor albedo / emit textures are copied into the material record (MaterialRec::solid)
    std::vector<MaterialRec> mats = s->materials;
    for (MaterialRec& m : mats) {
      m.solid = 0;
      if (m.type != MT_METAL && m.type != MT_DIELECTRIC && m.tex >= 0 && (size_t)m.tex < s->textures.size() &&
          s->textures[m.tex].type == TT_SOLID) {
        m.solid = 1;
        m.r = s->textures[m.tex].f0; m.g = s->textures[m.tex].f1; m.b = s->textures[m.tex].f2;
      }
    }
    d.all_diffuse_solid = 1;
    for (const MaterialRec& m : mats)
      if (!((m.type == MT_LAMBERTIAN || m.type == MT_DIFFUSE_LIGHT) && m.solid)) d.all_diffuse_solid = 0;
    if ((rc = dev_upload(s, &d.materials, mats))) return rc;
  }
  if ((rc = dev_upload(s, &d.textures, s->textures))) return rc;
  if ((rc = dev_upload(s, &d.noise, s->noise_tables))) return rc;
  if ((rc = dev_upload(s, &d.texels, s->texels))) return rc;
  d.num_prims = n;
  d.num_nodes = n > 1 ? n - 1 : 1;
  d.has_instances = s->inst_range.size() > 1 ? 1u : 0u;
  d.num_insts = (uint32_t)s->inst_range.size();
  d.num_inst_ops = (uint32_t)s->inst_ops.size();
  d.has_tri_shade = s->tri_shade.empty() ? 0u : 1u;
  d.has_media = 0;
  for (uint32_t m : s->prim_meta)
    if ((m & 7u) >= PT_MEDIUM_SPHERE) { d.has_media = 1; break; }
  RTW_CUDA_TRY(cudaEventRecord(ev[1]));

  // ---- persistent outputs ---------------------------------------------------------------------
  float4 *d_enc, *d_geom, *d_nodes;
  int32_t* d_slot_prim;
  uint32_t* d_slot_meta;
  int2* d_slot_ms;
  if ((rc = dev_alloc(s, &d_slot_meta, n))) return rc;
  if ((rc = dev_alloc(s, &d_slot_ms, n))) return rc;
  if ((rc = dev_alloc(s, &d_enc, 3 * (size_t)n))) return rc;
  if ((rc = dev_alloc(s, &d_geom, 3 * (size_t)n))) return rc;
  if ((rc = dev_alloc(s, &d_nodes, 4 * (size_t)d.num_nodes))) return rc;
  if ((rc = dev_alloc(s, &d_slot_prim, n))) return rc;

  // ---- scratch ------------------------------------------------------------------------------------
  const uint32_t sort_blocks = std::min<uint32_t>(std::max<uint32_t>(1, (n + 2047) / 2048), 4u * (uint32_t)s->num_sms);
  const uint32_t chunk = ((n + sort_blocks - 1) / sort_blocks + SORT_THREADS - 1) / SORT_THREADS * SORT_THREADS;
  std::vector<void*> scratch;
  auto salloc = [&](void** p, size_t bytes) -> int {
    cudaError_t e = dev_malloc(p, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) { cudaGetLastError(); return set_error(RTW_ERR_NOMEM, "cudaMalloc (build scratch) failed"); }
    scratch.push_back(*p);
    return RTW_OK;
  };
  float4 *d_lo, *d_hi, *d_raw;
  if ((rc = salloc((void**)&d_raw, sizeof(float4) * 3 * (size_t)n))) return rc;
  RTW_CUDA_TRY(cudaMemcpy(d_raw, s->raw_geom.data(), sizeof(float4) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  uint64_t *d_k0, *d_k1;
  uint32_t *d_v0, *d_v1, *d_hist, *d_bounds, *d_nparent, *d_lparent, *d_flags, *d_heights, *d_info, *d_collapsed;
  if ((rc = salloc((void**)&d_collapsed, 4ull * n))) return rc;
  uint2* d_nrange;
  float *d_root, *d_ncost;
  if ((rc = salloc((void**)&d_nrange, 8ull * n))) return rc;
  if ((rc = salloc((void**)&d_ncost, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_lo, sizeof(float4) * n))) return rc;
  if ((rc = salloc((void**)&d_hi, sizeof(float4) * n))) return rc;
  if ((rc = salloc((void**)&d_k0, 8ull * n))) return rc;
  if ((rc = salloc((void**)&d_k1, 8ull * n))) return rc;
  if ((rc = salloc((void**)&d_v0, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_v1, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_hist, 4ull * 256 * sort_blocks))) return rc;
  if ((rc = salloc((void**)&d_bounds, 4 * 8))) return rc;
  if ((rc = salloc((void**)&d_nparent, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_lparent, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_flags, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_heights, 4ull * n))) return rc;
  if ((rc = salloc((void**)&d_root, 4 * 12))) return rc;  // box (6), height, root collapsed, rotations applied
  d_info = (uint32_t*)(d_root + 6);

  const uint32_t T = 256, G = (n + T - 1) / T;
  k_prim_setup<<<G, T>>>(n, d_raw, d.prim_meta, d.inst_range, d.inst_ops, time0, time1, d_enc, d_lo, d_hi);
  {
    uint32_t init[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u, 0u};
    RTW_CUDA_TRY(cudaMemcpy(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  k_bounds<<<std::min<uint32_t>(G, 4u * (uint32_t)s->num_sms), T>>>(n, d_lo, d_hi, d_bounds);
  k_morton<<<G, T>>>(n, d_lo, d_hi, d_bounds, d_k0, d_v0);
  for (int pass = 0; pass < 8; ++pass) {
    int shift = 8 * pass;
    k_sort_hist<<<sort_blocks, SORT_THREADS>>>(n, chunk, d_k0, shift, d_hist);
    k_sort_scan<<<1, 1024>>>(256 * sort_blocks, d_hist);
    k_sort_scatter<<<sort_blocks, SORT_THREADS>>>(n, chunk, d_k0, d_v0, d_k1, d_v1, shift, d_hist);
    std::swap(d_k0, d_k1);
    std::swap(d_v0, d_v1);
  }
  RTW_CUDA_TRY(cudaMemset(d_flags, 0, 4ull * n));
  RTW_CUDA_TRY(cudaMemset(d_heights, 0, 4ull * n));
  if (n > 1) k_karras<<<G, T>>>((int)n, d_k0, d_nparent, d_lparent, d_nrange);
  uint32_t h_bounds[8];
  RTW_CUDA_TRY(cudaMemcpy(h_bounds, d_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost));
  const float amax = ord2f(h_bounds[6]);
  RefitParams rp;
  rp.pad = amax * (1.0f / 1048576.0f);
  // SAH constants (measured on B200, profiles/r01_sah_sweep.txt): pair step = primitive test = 1 is
  // best for the mesh scenes; a scene of <= 32 primitives is kept FLAT (one leaf): every lane of a
  // warp then walks the same primitive list in the same order — no hierarchy beats that on a SIMT
  // machine (Cornell box: +11% over the best hierarchy).  Overridable for experiments.
  rp.c_pair = 0.5f;  // r01 final A/B (vote-terminated walk): 0.5 vs 1.0 = stress +4 %, jumpy +1.4 %, monument +1 %, cow +0.8 %
  rp.c_prim = 1.0f;
  rp.max_leaf = 32;
  uint32_t flat_max = 32;
  if (const char* e = getenv("RTW_SAH_PAIR_COST")) rp.c_pair = (float)atof(e);
  if (const char* e = getenv("RTW_MAX_LEAF")) rp.max_leaf = (uint32_t)std::min(std::max(1, atoi(e)), 32);  // 5 bits in a packed child reference
  if (const char* e = getenv("RTW_FLAT_SCENE_MAX")) flat_max = (uint32_t)std::max(0, atoi(e));
  rp.force_flat = (n <= flat_max && n <= rp.max_leaf) ? 1u : 0u;
  k_refit<<<G, T>>>(n, d_v0, d_lo, d_hi, rp, d_nparent, d_lparent, d_nrange, d_flags, d_heights, d_ncost, d_collapsed,
                    d_nodes, d_root, d_info);
  if (n > 1) k_sort_leaf_ranges<<<G, T>>>(n - 1, d_collapsed, d_nparent, d_nrange, d.prim_meta, d_v0);
  k_emit_leaves<<<G, T>>>(n, d_v0, d_enc, d.prim_meta, d.prim_mat, d.prim_shade, d_geom, d_slot_prim, d_slot_meta, d_slot_ms);
  // tree rotations: RTW_ROTATE = number of bottom-up passes (default 4, 0 = the plain Karras tree)
  int rotate_passes = 4;  // r02 A/B (traverse ms, cow / monument / stress): 0: 22.4 / 32.9 / 378.7, 1: 19.1 / 27.7 / 365.9, 2: 18.8 / 27.0 / 361.5, 4: 18.6 / 26.3 / 358.6; +0.1 ms of build per pass on 6 k primitives, +8 ms on 11 M
  if (const char* e = getenv("RTW_ROTATE")) rotate_passes = std::max(0, std::min(atoi(e), 8));
  uint32_t rotations = 0;
  if (n > 2 && rotate_passes > 0 && !rp.force_flat) {
    uint32_t h_info[4] = {0, 0, 0, 0};
    RTW_CUDA_TRY(cudaMemcpy(h_info, d_info, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));  // [0] height, [1] root collapsed
    if (!h_info[1]) {
      uint32_t* d_live = (uint32_t*)d_ncost;  // the SAH costs are no longer needed
      const uint32_t G2 = (2 * (n - 1) + T - 1) / T;
      k_rotate_live<<<G, T>>>(n - 1, d_collapsed, d_nparent, d_live);
      RTW_CUDA_TRY(cudaMemset(d_info + 2, 0, sizeof(uint32_t)));
      uint8_t* d_start = (uint8_t*)d_v1;  // sort scratch (4 n bytes), free since the last radix pass
      for (int pass = 0; pass < rotate_passes; ++pass) {
        RTW_CUDA_TRY(cudaMemset(d_flags, 0, 4ull * n));
        k_rotate_starts<<<G2, T>>>(n - 1, d_live, d_nodes, d_start);
        k_rotate<<<G2, T>>>(n - 1, d_start, d_nparent, d_flags, d_heights, d_nodes, d_info);
      }
      RTW_CUDA_TRY(cudaMemcpy(h_info, d_info, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      rotations = h_info[2];
      if (h_info[0] + 2 > RTW_STACK_SIZE) {  // (never seen) too deep for the traversal stack: back to the Karras tree
        RTW_CUDA_TRY(cudaMemset(d_flags, 0, 4ull * n));
        RTW_CUDA_TRY(cudaMemset(d_heights, 0, 4ull * n));
        k_karras<<<G, T>>>((int)n, d_k0, d_nparent, d_lparent, d_nrange);
        k_refit<<<G, T>>>(n, d_v0, d_lo, d_hi, rp, d_nparent, d_lparent, d_nrange, d_flags, d_heights, d_ncost, d_collapsed,
                          d_nodes, d_root, d_info);
        rotations = 0;
      }
    }
  }
  if (getenv("RTW_BUILD_VERBOSE")) fprintf(stderr, "[rtw_build] %u primitives, %d rotation passes, %u rotations applied\n", n, rotate_passes, rotations);
  d.nodes4 = nullptr;
  if (const char* e = getenv("RTW_WIDE")) {  // the 4-wide records exist only for the experiment that reads them
    if (atoi(e) != 0) {
      float4* d_nodes4;
      if ((rc = dev_alloc(s, &d_nodes4, 8 * (size_t)d.num_nodes))) return rc;
      k_build_wide<<<(d.num_nodes + T - 1) / T, T>>>(d.num_nodes, d_nodes, d_nodes4);
      d.nodes4 = d_nodes4;
    }
  }
  d.top_nodes = nullptr;
  d.top_count = 0;
#if RTW_TOP_TREE > 0
  {
    float4* d_top;
    uint32_t *d_top_src, *d_top_count;
    if ((rc = dev_alloc(s, &d_top, 4 * (size_t)RTW_TOP_TREE))) return rc;
    if ((rc = dev_alloc(s, &d_top_src, (size_t)RTW_TOP_TREE + 1))) return rc;
    d_top_count = d_top_src + RTW_TOP_TREE;
    k_top_tree<<<1, 32>>>(d_nodes, d.num_nodes, RTW_TOP_TREE, d_top, d_top_src, d_top_count);
    RTW_CUDA_TRY(cudaMemcpy(&d.top_count, d_top_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    d.top_nodes = d_top;
  }
#endif
  RTW_CUDA_TRY(cudaGetLastError());
  RTW_CUDA_TRY(cudaEventRecord(ev[2]));
  RTW_CUDA_TRY(cudaEventSynchronize(ev[2]));
  float h_root[8];
  RTW_CUDA_TRY(cudaMemcpy(h_root, d_root, sizeof(h_root), cudaMemcpyDeviceToHost));
  memcpy(s->root_box, h_root, 6 * sizeof(float));
  memcpy(&s->bvh_height, &h_root[6], 4);
  {
    uint32_t root_collapsed = 0;
    memcpy(&root_collapsed, &h_root[7], 4);
    d.flat_count = (root_collapsed && n <= 32) ? n : 0u;
  }
  cudaDeviceSynchronize();  // the blocks go back to the cache (rtw_mem.cu): nothing may still be using them
  for (void* p : scratch) mem_free(p);

  // compact pairs: only worth their decode instructions when the hierarchy lives in HBM (>= 2^20 primitives)
  d.nodes_c = nullptr;
  {
    bool want = n >= (1u << 20);
    if (const char* e = getenv("RTW_COMPACT")) want = atoi(e) != 0;
    if (want && n > 1 && n < (1u << 26)) {
      QuantGrid g;
      bool ok = true;
      for (int a = 0; a < 3; ++a) {
        const float lo = h_root[a], hi = h_root[3 + a];
        g.lo[a] = lo;
        g.step[a] = (hi > lo) ? (hi - lo) / 65535.0f * 1.000001f : 1.0f;
        ok = ok && std::isfinite(lo) && std::isfinite(hi) && std::isfinite(g.step[a]) && g.step[a] > 0.0f &&
             std::fma(65535.0f, g.step[a], lo) >= hi;
      }
      if (ok) {
        uint4* d_nc;
        uint32_t* d_bad;
        if ((rc = dev_alloc(s, &d_nc, 2 * (size_t)d.num_nodes))) return rc;
        if ((rc = dev_alloc(s, &d_bad, 1))) return rc;
        RTW_CUDA_TRY(cudaMemset(d_bad, 0, sizeof(uint32_t)));
        k_build_compact<<<(d.num_nodes + T - 1) / T, T>>>(d.num_nodes, d_nodes, g, d_nc, d_bad);
        uint32_t bad = 1;
        RTW_CUDA_TRY(cudaMemcpy(&bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (!bad) {
          d.nodes_c = d_nc;
          for (int a = 0; a < 3; ++a) { d.grid_lo[a] = g.lo[a]; d.grid_step[a] = g.step[a]; }
        }
      }
    }
  }

  d.nodes = d_nodes;
  d.geom = d_geom;
  d.raw_geom = d_enc;
  d.slot_prim = d_slot_prim;
  d.slot_meta = d_slot_meta;
  d.slot_ms = d_slot_ms;

  float ms_up = 0.f, ms_build = 0.f;
  cudaEventElapsedTime(&ms_up, ev[0], ev[1]);
  cudaEventElapsedTime(&ms_build, ev[1], ev[2]);
  for (auto& e : ev) cudaEventDestroy(e);
  if (s->bvh_height + 2 > RTW_STACK_SIZE)
    return set_error(RTW_ERR_UNSUPPORTED, "LBVH deeper than the traversal stack (" + std::to_string(s->bvh_height) + ")");
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->num_prims = n;
    stats->num_nodes = d.num_nodes;
    stats->max_depth = s->bvh_height;
    stats->num_instances = (uint32_t)s->inst_range.size();
    stats->ms_build = ms_build;
    stats->ms_upload = ms_up;
    stats->device_bytes = s->device_bytes;
  }
  return RTW_OK;
}

}  // namespace rtw
