// Closest-hit traversal of the pair-layout LBVH (rtw_bvh.cu): the GPU replacement of
//   world.hit(r, t_min, t_max)  =  list rule (hittable/mod.rs:57-69) over BvhNode::hit (bvh.rs:101-120)
//                                   over Aabb::hit (aabb.rs:23-48) over the primitives' hit().
//
// Result contract (exceptions 1-2): the primitive with the smallest accepted t;
// among primitives with bit-equal t the one with the HIGHEST canonical id — exactly what the
// reference's flat list in canonical order returns ("last tested wins", t == t_max is accepted).
// The result is therefore independent of the traversal order, which leaves the scheduling free:
//
// traverse_persistent(): persistent threads in the style of Aila & Laine 2009 ("while-while" with
// per-lane refill).  Every lane owns one ray.  Each trip of the outer loop (a) refills idle lanes
// with new rays from a device-side cursor once enough of the warp is idle, (b) lets every lane walk
// internal pairs until it holds a leaf, (c) tests the held leaves together.  Lanes that finish
// write their hit and go idle until the next refill, so a warp never waits for its slowest ray
// with more than REFILL_IDLE lanes parked.  (r01 baseline, one lane per ray and one if/else loop:
// 7.45 of 32 threads active per issued instruction — profiles/r01a_ncu_full_baseline.csv.)
#pragma once
#include "rtw_device.cuh"

namespace rtw {

#define RTW_LINK_DONE ((int32_t)0x80000000)  // ~slot never reaches it (slot < 2^28)

// ---- tuning switches.  The defaults are the measured best on B200; every alternative was A/B-tested on the same
// box and is kept compiled out for the record (numbers: profiles/r01_sweeps.txt).
#ifndef RTW_REFILL_IDLE
#define RTW_REFILL_IDLE 12      // refill once this many lanes of the warp are idle (1 / 4 / 8 / 12 within 2 %)
#endif
#ifndef RTW_NODE_EXIT_LANES
#define RTW_NODE_EXIT_LANES 8   // the node phase ends when fewer lanes than this still walk and a leaf waits (0 = never)
#endif
#ifndef RTW_SPECULATE
#define RTW_SPECULATE 0         // postpone the first leaf and keep walking (Aila-Laine): +8 registers, slower
#endif
#ifndef RTW_LEAF_PHASE_DRAIN
#define RTW_LEAF_PHASE_DRAIN 0  // leaf phase tests every leaf a lane pops, not at most two: no gain
#endif
#ifndef RTW_RECT_PERM_CACHE
#define RTW_RECT_PERM_CACHE 0   // permute the ray per rectangle run instead of one code path per orientation: slower
#endif
#ifndef RTW_CURSOR_CHUNKS
#define RTW_CURSOR_CHUNKS 0     // private 32-entry chunks with a prefetched cursor: 3-13 % slower
#endif
#ifndef RTW_PREFETCH_FAR
#define RTW_PREFETCH_FAR 0      // prefetch.global.L2 of the postponed child: 3-8 % slower
#endif
#ifndef RTW_TOP_TREE_GENERIC
#define RTW_TOP_TREE_GENERIC 0  // RTW_TOP_TREE > 0 only: generic loads instead of an LDS / LDG branch
#endif
// RTW_TOP_TREE (rtw_device.cuh): shared-memory copy of the top of the tree — 5-14 % slower, default 0.
#ifndef RTW_STACK_PACKED
#define RTW_STACK_PACKED 0      // 4-byte stack entries (leaf: 0x80000000 | (count - 1) << 26 | first slot) instead of int2: 2-4 % slower
#endif

// The traversal stack lives in local memory (96 entries per lane); an entry is a child reference (link, meta).
#if RTW_STACK_PACKED
typedef uint32_t StackEntry;  // needs first slot < 2^26 and leaves of <= 32 primitives
__device__ __forceinline__ StackEntry stack_pack(int32_t link, uint32_t meta) {
  return link >= 0 ? (uint32_t)link : (0x80000000u | ((meta - 1u) << 26) | (uint32_t)(~link));
}
__device__ __forceinline__ void stack_unpack(StackEntry e, int32_t& link, uint32_t& meta) {
  const bool leaf = (e & 0x80000000u) != 0u;
  link = leaf ? ~(int32_t)(e & 0x3FFFFFFu) : (int32_t)e;
  meta = leaf ? ((e >> 26) & 31u) + 1u : 0u;
}
#else
typedef int2 StackEntry;
__device__ __forceinline__ StackEntry stack_pack(int32_t link, uint32_t meta) { return make_int2(link, (int)meta); }
__device__ __forceinline__ void stack_unpack(StackEntry e, int32_t& link, uint32_t& meta) { link = e.x; meta = (uint32_t)e.y; }
#endif

// Slab test of one child record against the ray: aabb.rs:23-48 with (a) the reciprocal hoisted out
// of the node loop (1/d is the same value every time), (b) a NON-strict reject (the reference
// rejects t_max <= t_min; we keep t_max == t_min so that exact-t ties are still visited) and
// (c) the exit distance inflated by 4 ulp, so that rounding in the slab arithmetic can only make
// the test more conservative than the exact-arithmetic one.  NaNs from 0*inf are dropped by
// fminf/fmaxf exactly like Rust's f32::min/max do.
__device__ __forceinline__ bool slab(float4 lo, float4 hi, v3 o, v3 inv, float t_min, float t_max, float& t_near) {
  float t0 = (lo.x - o.x) * inv.x, t1 = (hi.x - o.x) * inv.x;
  float tn = inv.x < 0.0f ? t1 : t0, tf = inv.x < 0.0f ? t0 : t1;
  t_min = fmaxf(tn, t_min);
  t_max = fminf(tf, t_max);
  t0 = (lo.y - o.y) * inv.y; t1 = (hi.y - o.y) * inv.y;
  tn = inv.y < 0.0f ? t1 : t0; tf = inv.y < 0.0f ? t0 : t1;
  t_min = fmaxf(tn, t_min);
  t_max = fminf(tf, t_max);
  t0 = (lo.z - o.z) * inv.z; t1 = (hi.z - o.z) * inv.z;
  tn = inv.z < 0.0f ? t1 : t0; tf = inv.z < 0.0f ? t0 : t1;
  t_min = fmaxf(tn, t_min);
  t_max = fminf(tf, t_max);
  t_near = t_min;
  float t_far = __fmaf_rn(fabsf(t_max), 4.76837158e-7f, t_max);
  return t_min <= t_far;
}

// The same cull with one FFMA per plane: t = b * inv + (-o * inv).  This is NOT the reference's expression, so its
// result may differ from the exact-arithmetic slab distance T = (b - o) / d by
//     |t - T| <= |T| * 3u + |o * inv| * 1.5u        (u = 2^-24: roundings of inv, of o * inv and of the FFMA)
// — a relative part and an ABSOLUTE part (cancellation when the plane is near the origin).  A box may only be skipped
// when the exact interval is empty, so the far distance is pushed out by both: relative 2^-20 of |near| + |far|, absolute
// 2 * slack with slack = 2^-22 * max over the axes of |o * inv| (per ray, RaySlab::set).  On an axis with d = 0 the
// products are +-inf / NaN exactly as in the reference's form or NaN where that had +-inf: fminf / fmaxf drop NaNs,
// which can only keep a box, never lose one.  Half the instructions of slab(): 6 FFMA + 6 FMNMX + 2 FMNMX3 per box.
#ifndef RTW_SLAB_FMA
#define RTW_SLAB_FMA 0  // r02 A/B: cow 12.15 -> 12.06 ms, jumpy 4.54 -> 4.49, monument 18.9 -> 31.7 (the absolute slack of far-away origins un-culls the tree): off
#endif
struct RaySlab {
  v3 inv, nc;    // 1 / d (aabb.rs:29), -(o * inv)
  float slack2;  // 2 x the absolute slack
  __device__ __forceinline__ void set(v3 o, v3 d) {
    inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const v3 c = mk(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    nc = mk(-c.x, -c.y, -c.z);
    const float INF = __int_as_float(0x7f800000);
    // axes whose product is not finite (d = 0) contribute no slack: their distances are +-inf / NaN, not rounded values
    const float ax = fabsf(c.x) < INF ? fabsf(c.x) : 0.0f, ay = fabsf(c.y) < INF ? fabsf(c.y) : 0.0f,
                az = fabsf(c.z) < INF ? fabsf(c.z) : 0.0f;
    slack2 = fmaxf(fmaxf(ax, ay), az) * 4.76837158e-7f;  // 2 * 2^-22
  }
};
__device__ __forceinline__ bool slab_fma(float4 lo, float4 hi, const RaySlab& rs, float t_min, float t_max, float& t_near) {
  const float x0 = __fmaf_rn(lo.x, rs.inv.x, rs.nc.x), x1 = __fmaf_rn(hi.x, rs.inv.x, rs.nc.x);
  const float y0 = __fmaf_rn(lo.y, rs.inv.y, rs.nc.y), y1 = __fmaf_rn(hi.y, rs.inv.y, rs.nc.y);
  const float z0 = __fmaf_rn(lo.z, rs.inv.z, rs.nc.z), z1 = __fmaf_rn(hi.z, rs.inv.z, rs.nc.z);
  const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_min));
  const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_max));
  t_near = tn;
  const float far_out = __fmaf_rn(fabsf(tn) + fabsf(tf), 9.53674316e-7f, tf) + rs.slack2;
  return tn <= far_out;
}

// sm_100a: one 256-bit load (LDG.E.256) instead of two LDG.E.128 — a child-pair fetch is the walk's only divergent load
// and the L1TEX pipeline its busiest unit (r02 captures: 64-75 % busy), so the number of load instructions per step
// matters, not only the bytes.  p must be 32-byte aligned (pairs are 64-byte records, compact pairs 32-byte records).
#ifndef RTW_LDG256
#define RTW_LDG256 1
#endif
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
#if RTW_LDG256
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
#else
  a = __ldg(reinterpret_cast<const float4*>(p));
  b = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
#if RTW_LDG256
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
#else
  a = __ldg(reinterpret_cast<const uint4*>(p));
  b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct TraverseCounters {
  uint32_t pairs = 0;
  uint32_t prims = 0;
  uint32_t prim_bytes = 0;
};

// Stage the top of the tree (rtw_bvh.cu: k_top_tree) in shared memory; call with the whole block.
__device__ __forceinline__ void stage_top_tree(const SceneDev& sc, float4* top_smem) {
#if RTW_TOP_TREE > 0
  for (uint32_t i = threadIdx.x; i < 4u * sc.top_count; i += blockDim.x) top_smem[i] = sc.top_nodes[i];
  __syncthreads();
#endif
}

// IO policy of traverse_persistent:
//   bool load(uint32_t index, v3& o, v3& d, float& time, float& t_min, float& t_max, int32_t& slot0, bool& resumed)
//        fetch ray `index`; slot0 >= 0 / resumed: the ray was suspended with the hit (slot0, t_max)
//   void store(uint32_t index, v3 o, v3 d, float time, int32_t slot, float t, uint32_t meta)   publish its closest hit
//   void suspend(uint32_t index, int32_t slot, float t)   publish the closest hit so far, mark the ray pending
//   void rng_key(Rng& rng)   (pixel, sample, stage, seed) of the current ray — only called when a medium is tested
//   static constexpr int kSuspendLanes   drain cut-off, see below (0 = off, the default)
// MEDIA = the scene contains ConstantMedium primitives (compiled out otherwise: the keyed draw and the
// boundary tests cost registers and branches in the hottest loop).
//
// Drain cut-off (IO::kSuspendLanes > 0, measured slower and off by default): once the cursor has run dry a warp
// only drains what it holds, and a launch ends with every SM waiting for a handful of long rays (67 us of a 260 us
// launch on the cow scene).  A warp left with <= kSuspendLanes rays SUSPENDS them: the closest hit found so far is
// published with a "pending" mark, the next launch restarts the ray with that hit as its t_max (the result of a
// closest-hit query does not depend on how it is split — same tie rule), and the shade kernel leaves pending
// slots alone.  A restarted ray is never suspended again, so every ray finishes within two launches.
// NODES selects the node records the walk reads:
//   NODES_PAIR    64-byte fp32 child pairs (SceneDev::nodes) — the default; bit-identical boxes to the build
//   NODES_WIDE    128-byte records with the four grandchild boxes (SceneDev::nodes4): half the dependent fetches,
//                 more bytes — measured slower everywhere, experiment only (RTW_WIDE=1)
//   NODES_COMPACT 32-byte pairs with 16-bit boxes on the scene grid (SceneDev::nodes_c): half the bytes of a step for
//                 36 decode instructions — for hierarchies that live in HBM, where traversal runs at the memory
//                 system's random-gather bandwidth (config C5)
enum { NODES_PAIR = 0, NODES_WIDE = 1, NODES_COMPACT = 2 };

template <bool COUNT, bool MEDIA, int NODES = NODES_PAIR, class IO>
__device__ __forceinline__ void traverse_persistent(const SceneDev& sc, IO& io, uint32_t count, uint32_t* cursor,
                                                    TraverseCounters& cnt, const float4* top_smem = nullptr) {
  const uint32_t lane = threadIdx.x & 31;
#if RTW_TOP_TREE > 0
  const int32_t root_link = (NODES == NODES_PAIR && sc.top_count) ? (int32_t)RTW_LINK_TOP : 0;  // only the fp32-pair walk reads the staged copy
#else
  const int32_t root_link = 0;
#endif
  const uint32_t lane_lt = (1u << lane) - 1u;
  bool active = false;
  bool resumed = false;    // this ray was suspended by the previous launch: finish it
  bool exhausted = false;  // warp-uniform: the cursor ran past `count`
  // per-ray state
  uint32_t index = 0;
  v3 o = mk(0, 0, 0), d = mk(0, 0, 0), inv = mk(0, 0, 0), oi = o, di = d;
  RaySlab rs;
  rs.inv = inv; rs.nc = inv; rs.slack2 = 0.f;
#if RTW_SLAB_FMA
#define RTW_SLAB(lo, hi, tmin, tmax, tnear) slab_fma(lo, hi, rs, tmin, tmax, tnear)
#else
#define RTW_SLAB(lo, hi, tmin, tmax, tnear) slab(lo, hi, o, inv, tmin, tmax, tnear)
#endif
  float time = 0.f, t_min = 0.f, best_t = 0.f;
  int32_t best_slot = -1, best_id = -2, link = RTW_LINK_DONE, pl_link = 0;
  uint32_t best_meta = 0, meta = 0, pl_meta = 0, cur_inst = 0, cur_pm = 0xffffffffu;
#if RTW_RECT_PERM_CACHE
  float oA = 0.f, oB = 0.f, oK = 0.f, dA = 0.f, dB = 0.f, dK = 0.f;  // ray permuted for the current rectangle run
#endif
  StackEntry stack[RTW_STACK_SIZE];
  int sp = 0;
#if RTW_CURSOR_CHUNKS
  uint32_t chunk_pos = 0, chunk_end = 0;  // warp-uniform: unused entries of the warp's current chunk
  uint32_t next_base = 0;                 // lane 0: cursor value of the prefetched next chunk
  if (lane == 0) next_base = atomicAdd(cursor, 32u);
#endif

  for (;;) {
    // ---- (a) refill idle lanes -----------------------------------------------------------------
    const uint32_t idle = __ballot_sync(0xffffffffu, !active);
#if RTW_CURSOR_CHUNKS
    // The warp owns a chunk of 32 consecutive entries and has the NEXT chunk's cursor fetch already in flight
    // (issued when the current chunk was opened): a refill never waits for an atomic round trip — 10.6 % of the
    // kernel's stall samples on the flat Cornell scene, where all 32 lanes finish together (profiles/r01h).
    uint32_t base = 0, take = 0;
    if (idle != 0 && !exhausted && (__popc(idle) >= RTW_REFILL_IDLE || idle == 0xffffffffu)) {
      if (chunk_pos == chunk_end) {  // open the prefetched chunk, prefetch the one after
        chunk_pos = __shfl_sync(0xffffffffu, next_base, 0);
        if (chunk_pos >= count) {
          exhausted = true;
          chunk_end = chunk_pos;
        } else {
          chunk_end = min(chunk_pos + 32u, count);
          if (lane == 0) next_base = atomicAdd(cursor, 32u);
        }
      }
      take = min((uint32_t)__popc(idle), chunk_end - chunk_pos);
      base = chunk_pos;
      chunk_pos += take;
    }
    if (take != 0) {
      if (!active && (uint32_t)__popc(idle & lane_lt) >= take) {
        // no entry left in this chunk for this lane: next trip
      } else
#else
    if (idle != 0 && !exhausted && (__popc(idle) >= RTW_REFILL_IDLE || idle == 0xffffffffu)) {
      const int leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(idle));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (base + __popc(idle) >= count) exhausted = true;
#endif
      if (!active) {
        index = base + __popc(idle & lane_lt);
        float t_max;
        int32_t slot0 = -1;
        if (index < count && io.load(index, o, d, time, t_min, t_max, slot0, resumed)) {
#if RTW_SLAB_FMA
          rs.set(o, d);
#else
          inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // aabb.rs:29
#endif
          best_t = t_max; best_slot = slot0; best_id = -2;
          best_meta = slot0 >= 0 ? __ldg(sc.slot_meta + slot0) : 0u;
          oi = o; di = d; cur_inst = 0; cur_pm = 0xffffffffu;
          sp = 0; link = root_link; meta = 0; pl_meta = 0;
          active = true;
        }
      }
    }
    const uint32_t m_active = __ballot_sync(0xffffffffu, active);
    if (m_active == 0) {
      if (exhausted) break;
      continue;
    }
    if (IO::kSuspendLanes > 0 && exhausted && __popc(m_active) <= IO::kSuspendLanes) {
      if (active && !resumed) {
        io.suspend(index, best_slot, best_t);
        active = false;
      }
      if (__ballot_sync(0xffffffffu, active) == 0) break;
    }
    // ---- (b) node phase: walk internal pairs -----------------------------------------------------------
    // Plain while-while keeps every lane that has reached a leaf waiting for the slowest one (measured on the
    // cow scene: 8.3 of 32 lanes execute a node step, profiles/r01h).  Two remedies, after Aila & Laine:
    //  * speculation: the first leaf a lane reaches is POSTPONED (pl_link / pl_meta) and the lane keeps walking
    //    until it holds a second one; testing the postponed leaf later can only find hits the early test would
    //    have found too (the closest hit does not depend on the order of the tests);
    //  * the phase ends by vote: when no searching lane is empty-handed, or when fewer than
    //    RTW_NODE_EXIT_LANES lanes would take another step while others have leaves waiting.
    for (;;) {
      const bool searching = active && link >= 0;
      const uint32_t m_search = __ballot_sync(0xffffffffu, searching);
      if (m_search == 0) break;
#if RTW_SPECULATE
      if (__ballot_sync(0xffffffffu, searching && pl_meta == 0u) == 0) break;
#endif
#if RTW_NODE_EXIT_LANES > 0
      if (__popc(m_search) < RTW_NODE_EXIT_LANES &&
          __any_sync(0xffffffffu, active && (pl_meta != 0u || (link < 0 && link != RTW_LINK_DONE))))
        break;
#endif
      if (NODES == NODES_WIDE && searching) {
        const float4* __restrict__ n = sc.nodes4 + 8 * (size_t)link;
        if (COUNT) cnt.pairs += 2;
        const float INF = __int_as_float(0x7f800000);
        float tk[4];
        int32_t lk[4];
        uint32_t mk4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 lo = __ldg(n + 2 * c), hi = __ldg(n + 2 * c + 1);
          float t;
          const int32_t l = __float_as_int(lo.w);
          const bool h = (l != RTW_LINK_DONE) && RTW_SLAB(lo, hi, t_min, best_t, t);
          lk[c] = h ? l : RTW_LINK_DONE;
          tk[c] = h ? t : INF;
          mk4[c] = __float_as_uint(hi.w);
        }
        // sort the four by entry distance (misses sort last): 5-comparator network
#define RTW_CSWAP(a, b)                                                  \
  {                                                                      \
    const bool sw = tk[b] < tk[a];                                       \
    const float tt = sw ? tk[a] : tk[b]; tk[a] = sw ? tk[b] : tk[a]; tk[b] = tt;          \
    const int32_t ll = sw ? lk[a] : lk[b]; lk[a] = sw ? lk[b] : lk[a]; lk[b] = ll;        \
    const uint32_t mm = sw ? mk4[a] : mk4[b]; mk4[a] = sw ? mk4[b] : mk4[a]; mk4[b] = mm; \
  }
        RTW_CSWAP(0, 1) RTW_CSWAP(2, 3) RTW_CSWAP(0, 2) RTW_CSWAP(1, 3) RTW_CSWAP(1, 2)
#undef RTW_CSWAP
        // hits form a prefix; the farthest is pushed first so that the nearest of the rest is popped first
        if (lk[3] != RTW_LINK_DONE) stack[sp++] = stack_pack(lk[3], mk4[3]);
        if (lk[2] != RTW_LINK_DONE) stack[sp++] = stack_pack(lk[2], mk4[2]);
        if (lk[1] != RTW_LINK_DONE) stack[sp++] = stack_pack(lk[1], mk4[1]);
        if (lk[0] != RTW_LINK_DONE) {
          link = lk[0]; meta = mk4[0];
        } else if (sp > 0) {
          stack_unpack(stack[--sp], link, meta);
        } else {
          link = RTW_LINK_DONE;
        }
      }
      if (NODES == NODES_COMPACT && searching) {
        const uint4* __restrict__ nc = sc.nodes_c + 2 * (size_t)link;
        uint4 a, b;
        ldg256(nc, a, b);
        if (COUNT) cnt.pairs++;
        const float sx = sc.grid_step[0], sy = sc.grid_step[1], sz = sc.grid_step[2];
        const float gx = sc.grid_lo[0], gy = sc.grid_lo[1], gz = sc.grid_lo[2];
        // the same expression k_build_compact verified the containment with
        const float4 l0 = make_float4(__fmaf_rn((float)(a.x & 0xffffu), sx, gx), __fmaf_rn((float)(a.x >> 16), sy, gy),
                                      __fmaf_rn((float)(a.y & 0xffffu), sz, gz), 0.f);
        const float4 l1 = make_float4(__fmaf_rn((float)(a.y >> 16), sx, gx), __fmaf_rn((float)(a.z & 0xffffu), sy, gy),
                                      __fmaf_rn((float)(a.z >> 16), sz, gz), 0.f);
        const float4 r0 = make_float4(__fmaf_rn((float)(b.x & 0xffffu), sx, gx), __fmaf_rn((float)(b.x >> 16), sy, gy),
                                      __fmaf_rn((float)(b.y & 0xffffu), sz, gz), 0.f);
        const float4 r1 = make_float4(__fmaf_rn((float)(b.y >> 16), sx, gx), __fmaf_rn((float)(b.z & 0xffffu), sy, gy),
                                      __fmaf_rn((float)(b.z >> 16), sz, gz), 0.f);
        // child word -> (link, meta): internal pair index, or leaf 0x80000000 | (count - 1) << 26 | first slot
        const int32_t ll = (a.w & 0x80000000u) ? ~(int32_t)(a.w & 0x3FFFFFFu) : (int32_t)a.w;
        const int32_t rl = (b.w & 0x80000000u) ? ~(int32_t)(b.w & 0x3FFFFFFu) : (int32_t)b.w;
        const uint32_t lm = ((a.w >> 26) & 31u) + 1u, rm = ((b.w >> 26) & 31u) + 1u;
        float tl, tr;
        const bool hl = RTW_SLAB(l0, l1, t_min, best_t, tl);
        const bool hr = RTW_SLAB(r0, r1, t_min, best_t, tr);
        if (hl && hr) {
          const bool left_first = tl <= tr;
          stack[sp++] = left_first ? stack_pack(rl, rm) : stack_pack(ll, lm);
          link = left_first ? ll : rl;
          meta = left_first ? lm : rm;
        } else if (hl) {
          link = ll; meta = lm;
        } else if (hr) {
          link = rl; meta = rm;
        } else if (sp > 0) {
          stack_unpack(stack[--sp], link, meta);
        } else {
          link = RTW_LINK_DONE;
        }
      }
      if (NODES == NODES_PAIR && searching) {
#if RTW_TOP_TREE > 0 && RTW_TOP_TREE_GENERIC
        // generic loads: the pair lies in shared memory (top of the tree) or in global memory
        const float4* n = (link & RTW_LINK_TOP) ? top_smem + 4 * (size_t)(link & (RTW_LINK_TOP - 1))
                                                : sc.nodes + 4 * (size_t)link;
        const float4 l0 = n[0], l1 = n[1], r0 = n[2], r1 = n[3];
#elif RTW_TOP_TREE > 0
        float4 l0, l1, r0, r1;
        if (link & RTW_LINK_TOP) {  // LDS.128 x4
          const float4* n = top_smem + 4 * (size_t)(link & (RTW_LINK_TOP - 1));
          l0 = n[0]; l1 = n[1]; r0 = n[2]; r1 = n[3];
        } else {  // LDG.128 x4 through the read-only path
          const float4* __restrict__ n = sc.nodes + 4 * (size_t)link;
          l0 = __ldg(n); l1 = __ldg(n + 1); r0 = __ldg(n + 2); r1 = __ldg(n + 3);
        }
#else
        const float4* __restrict__ n = sc.nodes + 4 * (size_t)link;
        float4 l0, l1, r0, r1;
        ldg256(n, l0, l1);
        ldg256(n + 2, r0, r1);
#endif
        if (COUNT) cnt.pairs++;
        float tl, tr;
        const bool hl = RTW_SLAB(l0, l1, t_min, best_t, tl);
        const bool hr = RTW_SLAB(r0, r1, t_min, best_t, tr);
        if (hl && hr) {
          const bool left_first = tl <= tr;
          const int2 far = left_first ? make_int2(__float_as_int(r0.w), __float_as_int(r1.w))
                                      : make_int2(__float_as_int(l0.w), __float_as_int(l1.w));
          stack[sp++] = stack_pack(far.x, (uint32_t)far.y);
#if RTW_PREFETCH_FAR
          // the postponed child is the likeliest later visit: start its 64-byte pair (or its first primitive)
          // on the way to L2 now — only matters when the hierarchy does not fit the caches (config C5)
          if (far.x >= 0) prefetch_l2(sc.nodes + 4 * (size_t)far.x);
          else prefetch_l2(sc.geom + 3 * (size_t)(~far.x));
#endif
          link = left_first ? __float_as_int(l0.w) : __float_as_int(r0.w);
          meta = left_first ? __float_as_uint(l1.w) : __float_as_uint(r1.w);
        } else if (hl) {
          link = __float_as_int(l0.w); meta = __float_as_uint(l1.w);
        } else if (hr) {
          link = __float_as_int(r0.w); meta = __float_as_uint(r1.w);
        } else if (sp > 0) {
          stack_unpack(stack[--sp], link, meta);
        } else {
          link = RTW_LINK_DONE;
        }
#if RTW_SPECULATE
        if (link < 0 && link != RTW_LINK_DONE && pl_meta == 0u) {  // first leaf in hand: postpone it, keep walking
          pl_link = link; pl_meta = meta;
          if (sp > 0) {
            stack_unpack(stack[--sp], link, meta);
          } else {
            link = RTW_LINK_DONE;
          }
        }
#endif
      }
    }
    // ---- (c) leaf phase: test the primitives of the held leaves (contiguous slot ranges) ---------------------
    // Inside a leaf the build sorted the slots by (instance, type): the ray is re-transformed /
    // re-permuted only when that key changes (cur_pm caches it).
#if RTW_LEAF_PHASE_DRAIN
    while (active && (pl_meta != 0u || (link < 0 && link != RTW_LINK_DONE))) {
#else
    for (int rep = 0; rep < 2 && active && (pl_meta != 0u || (link < 0 && link != RTW_LINK_DONE)); ++rep) {  // the postponed leaf, then the held one
#endif
      uint32_t first, nprim;
      if (pl_meta != 0u) {
        first = (uint32_t)(~pl_link); nprim = pl_meta; pl_meta = 0u;
      } else {
        first = (uint32_t)(~link); nprim = meta;
        if (sp > 0) {
          stack_unpack(stack[--sp], link, meta);
        } else {
          link = RTW_LINK_DONE;
        }
      }
      for (uint32_t k = 0; k < nprim; ++k) {
        const uint32_t slot = first + k;
        const uint32_t pm = __ldg(sc.slot_meta + slot);
        const uint32_t type = pm & 7u;
#if RTW_RECT_PERM_CACHE
        if (pm != cur_pm) {
          const uint32_t inst = pm >> RTW_META_TYPE_BITS;
          if (inst != cur_inst) {
            oi = o; di = d;
            if (inst != 0) ray_to_instance(sc, inst, oi, di);
            cur_inst = inst;
          }
          if (type >= PT_RECT_YZ && type <= PT_RECT_XY) {
            const int axis = (int)type - (int)PT_RECT_YZ;
            const int A = (axis == 0) ? 1 : 0, B = (axis == 2) ? 1 : 2;
            oA = comp(oi, A); oB = comp(oi, B); oK = comp(oi, axis);
            dA = comp(di, A); dB = comp(di, B); dK = comp(di, axis);
          }
          cur_pm = pm;
        }
#else
        {
          const uint32_t inst = pm >> RTW_META_TYPE_BITS;
          if (inst != cur_inst) {
            oi = o; di = d;
            if (inst != 0) ray_to_instance(sc, inst, oi, di);
            cur_inst = inst;
          }
        }
#endif
        if (COUNT) {
          cnt.prims++;
          cnt.prim_bytes += 4u + ((type == PT_SPHERE) ? 16u : (((type >= PT_RECT_YZ && type <= PT_RECT_XY) || type >= PT_MEDIUM_SPHERE) ? 32u : 48u));
        }
        const float4* __restrict__ g = sc.geom + 3 * (size_t)slot;
        float t;
        bool hit;
        if (MEDIA && type >= PT_MEDIUM_SPHERE) {  // ConstantMedium: one keyed draw (volumes.rs:58)
          Rng rng;
          io.rng_key(rng);
          hit = medium_t(type, g, oi, di, d, t_min, best_t, rng.gen_f32_keyed((uint32_t)__ldg(sc.slot_prim + slot)), t);
        }
#if RTW_RECT_PERM_CACHE
        else if (type >= PT_RECT_YZ && type <= PT_RECT_XY)
          hit = rect_t_perm(oA, oB, oK, dA, dB, dK, t_min, best_t, __ldg(g), __ldg(reinterpret_cast<const float*>(g + 1)), t);
#else
        // one straight-line path per orientation: the components are picked at compile time (no per-run
        // permutation of the ray: 21 % of the instructions of the flat Cornell scene, profiles/r01_final), and in
        // a flat scene the whole warp takes the same path
        else if (type == PT_RECT_XZ)
          hit = rect_t_perm(oi.x, oi.z, oi.y, di.x, di.z, di.y, t_min, best_t, __ldg(g), __ldg(reinterpret_cast<const float*>(g + 1)), t);
        else if (type == PT_RECT_XY)
          hit = rect_t_perm(oi.x, oi.y, oi.z, di.x, di.y, di.z, t_min, best_t, __ldg(g), __ldg(reinterpret_cast<const float*>(g + 1)), t);
        else if (type == PT_RECT_YZ)
          hit = rect_t_perm(oi.y, oi.z, oi.x, di.y, di.z, di.x, t_min, best_t, __ldg(g), __ldg(reinterpret_cast<const float*>(g + 1)), t);
#endif
        else
          hit = prim_t(type, g, oi, di, time, t_min, best_t, t);
        if (hit) {
          if (best_slot < 0 || t < best_t) {
            best_t = t; best_slot = (int32_t)slot; best_meta = pm; best_id = -2;
          } else {  // t == best_t: the later primitive of the canonical order wins (hittable/mod.rs:61-66)
            if (best_id == -2) best_id = __ldg(sc.slot_prim + best_slot);
            const int32_t id = __ldg(sc.slot_prim + slot);
            if (id > best_id) { best_slot = (int32_t)slot; best_meta = pm; best_id = id; }
          }
        }
      }
    }
    __syncwarp();
    if (active && link == RTW_LINK_DONE && pl_meta == 0u) {
      io.store(index, o, d, time, best_slot, best_t, best_meta);
      active = false;
    }
  }
}

// ---- pooled leaf tests ----------------------------------------------------------------------------------------------
// traverse_persistent() above lets every lane test the leaves it reaches itself: measured on the cow scene (r02f capture,
// tools/ncu_regions.py) the primitive tests are 32 % of the kernel's warp instructions and run at 5.6 of 32 lanes, and
// the node steps (57 %) at 17 of 32 because lanes that hold a leaf wait for the walkers.  traverse_pooled() separates the
// two kinds of work inside a warp:
//   * a lane only WALKS.  When it reaches a leaf it appends (lane, primitive slot) entries to a per-warp FIFO in shared
//     memory — one primitive per trip — pops its stack and walks on, without waiting for the result;
//   * as soon as 32 entries are queued the whole warp DRAINS them: lane j tests entry j against its owner's ray (kept in
//     shared memory), so the primitive tests run at full width whatever leaf each walker is in;
//   * a hit is merged into its owner's record with one 64-bit shared-memory atomicMin over (ordered t, ~canonical id):
//     smallest t first, then the LATER primitive of the canonical order — the list rule of hittable/mod.rs:57-69, which
//     does not depend on the order of the tests.
// A walker therefore culls with a closest-hit distance that may lag a few steps behind (it can only be too large:
// culling stays conservative and the result is unchanged); a lane whose stack is empty parks until its last entry has
// been drained.  A partial batch is drained when RTW_POOL_WAIT_LANES lanes are parked or nothing else can progress.
#ifndef RTW_POOL_WAIT_LANES
#define RTW_POOL_WAIT_LANES 8
#endif
#ifndef RTW_POOL_SERVICE_LANES
#define RTW_POOL_SERVICE_LANES 4  // the node loop runs until this many lanes hold a leaf / have finished
#endif
#ifndef RTW_POOL_REFILL_IDLE
#define RTW_POOL_REFILL_IDLE 8
#endif
struct LeafPool {  // per warp
  float ray[8][32];             // world ray of lane's current query: o.xyz, d.xyz, time, t_min
  unsigned long long best[32];  // (ordered t << 32) | (0xFFFFFFFE - canonical id); low word 0xFFFFFFFF = no hit yet
  int32_t best_slot[32];
  uint32_t queue[64];           // (owner lane << 27) | primitive slot
};
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <bool COUNT, int NODES, class IO>
__device__ __forceinline__ void traverse_pooled(const SceneDev& sc, IO& io, uint32_t count, uint32_t* cursor,
                                                TraverseCounters& cnt, LeafPool& lp) {
  static_assert(NODES == NODES_PAIR || NODES == NODES_COMPACT, "pooled walk: fp32 pairs or compact pairs");
  const uint32_t FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lane_lt = (1u << lane) - 1u;
  bool active = false;
  bool parked = false;     // stack empty, entries still queued: waits for a drain
  bool exhausted = false;  // warp-uniform: the cursor ran past `count`
  uint32_t index = 0;
  v3 o = mk(0, 0, 0), d = mk(0, 0, 0), inv = mk(0, 0, 0);
  float t_min = 0.f, best_t = 0.f, time = 0.f;
  int32_t link = RTW_LINK_DONE;
  uint32_t meta = 0;
  StackEntry stack[RTW_STACK_SIZE];
  int sp = 0;
  uint32_t head = 0, tail = 0;  // warp-uniform FIFO counters (monotonic)
  uint32_t my_last = 0;         // FIFO position just past this lane's last entry

  for (;;) {
    // ---- (a) refill idle lanes ---------------------------------------------------------------------------------------
    const uint32_t idle = __ballot_sync(FULL, !active);
    if (idle != 0 && !exhausted && (__popc(idle) >= RTW_POOL_REFILL_IDLE || idle == FULL)) {
      const int leader = __ffs(idle) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(idle));
      base = __shfl_sync(FULL, base, leader);
      if (base + __popc(idle) >= count) exhausted = true;
      if (!active) {
        index = base + __popc(idle & lane_lt);
        float t_max;
        int32_t slot0 = -1;
        bool resumed = false;
        if (index < count && io.load(index, o, d, time, t_min, t_max, slot0, resumed)) {
          inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // aabb.rs:29
          best_t = t_max;
          lp.ray[0][lane] = o.x; lp.ray[1][lane] = o.y; lp.ray[2][lane] = o.z;
          lp.ray[3][lane] = d.x; lp.ray[4][lane] = d.y; lp.ray[5][lane] = d.z;
          lp.ray[6][lane] = time; lp.ray[7][lane] = t_min;
          const uint32_t low = slot0 >= 0 ? 0xFFFFFFFEu - (uint32_t)__ldg(sc.slot_prim + slot0) : 0xFFFFFFFFu;
          lp.best[lane] = ((unsigned long long)float_to_ordered(t_max) << 32) | low;
          lp.best_slot[lane] = slot0;
          sp = 0; link = 0; meta = 0;
          my_last = head;
          active = true;
        }
      }
      __syncwarp();
    }
    if (__ballot_sync(FULL, active) == 0) {
      if (exhausted) break;
      continue;
    }
    // ---- (b) node steps until RTW_POOL_SERVICE_LANES lanes need service (hold a leaf, or are finished) or nobody walks --
    const uint32_t m_busy = __ballot_sync(FULL, active && !parked);
    for (;;) {
      const bool walking = active && link >= 0;
      const uint32_t m_walk = __ballot_sync(FULL, walking);
      if (m_walk == 0u || __popc(m_busy & ~m_walk) >= RTW_POOL_SERVICE_LANES) break;
      if (walking) {
        float4 l0, l1, r0, r1;
        int32_t ll, rl;
        uint32_t lm, rm;
        if (NODES == NODES_COMPACT) {
          const uint4* __restrict__ nc = sc.nodes_c + 2 * (size_t)link;
          uint4 a, b;
          ldg256(nc, a, b);
          const float sx = sc.grid_step[0], sy = sc.grid_step[1], sz = sc.grid_step[2];
          const float gx = sc.grid_lo[0], gy = sc.grid_lo[1], gz = sc.grid_lo[2];
          // the same expression k_build_compact verified the containment with
          l0 = make_float4(__fmaf_rn((float)(a.x & 0xffffu), sx, gx), __fmaf_rn((float)(a.x >> 16), sy, gy),
                           __fmaf_rn((float)(a.y & 0xffffu), sz, gz), 0.f);
          l1 = make_float4(__fmaf_rn((float)(a.y >> 16), sx, gx), __fmaf_rn((float)(a.z & 0xffffu), sy, gy),
                           __fmaf_rn((float)(a.z >> 16), sz, gz), 0.f);
          r0 = make_float4(__fmaf_rn((float)(b.x & 0xffffu), sx, gx), __fmaf_rn((float)(b.x >> 16), sy, gy),
                           __fmaf_rn((float)(b.y & 0xffffu), sz, gz), 0.f);
          r1 = make_float4(__fmaf_rn((float)(b.y >> 16), sx, gx), __fmaf_rn((float)(b.z & 0xffffu), sy, gy),
                           __fmaf_rn((float)(b.z >> 16), sz, gz), 0.f);
          ll = (a.w & 0x80000000u) ? ~(int32_t)(a.w & 0x3FFFFFFu) : (int32_t)a.w;
          rl = (b.w & 0x80000000u) ? ~(int32_t)(b.w & 0x3FFFFFFu) : (int32_t)b.w;
          lm = ((a.w >> 26) & 31u) + 1u; rm = ((b.w >> 26) & 31u) + 1u;
        } else {
          const float4* __restrict__ n = sc.nodes + 4 * (size_t)link;
          ldg256(n, l0, l1);
          ldg256(n + 2, r0, r1);
          ll = __float_as_int(l0.w); rl = __float_as_int(r0.w);
          lm = __float_as_uint(l1.w); rm = __float_as_uint(r1.w);
        }
        if (COUNT) cnt.pairs++;
        float tl, tr;
        const bool hl = slab(l0, l1, o, inv, t_min, best_t, tl);
        const bool hr = slab(r0, r1, o, inv, t_min, best_t, tr);
        if (hl && hr) {
          const bool left_first = tl <= tr;
          stack[sp++] = left_first ? stack_pack(rl, rm) : stack_pack(ll, lm);
          link = left_first ? ll : rl;
          meta = left_first ? lm : rm;
        } else if (hl) {
          link = ll; meta = lm;
        } else if (hr) {
          link = rl; meta = rm;
        } else if (sp > 0) {
          stack_unpack(stack[--sp], link, meta);
        } else {
          link = RTW_LINK_DONE;
        }
      }
    }
    // ---- (c) every lane that holds a leaf queues its next primitive ------------------------------------------------------
    {
      const bool holder = active && link < 0 && link != RTW_LINK_DONE;
      const uint32_t m = __ballot_sync(FULL, holder);
      if (holder) {
        lp.queue[(tail + __popc(m & lane_lt)) & 63u] = (lane << 27) | (uint32_t)(~link);
        my_last = tail + __popc(m);
        link -= 1;  // ~link + 1: the next slot of the leaf
        meta -= 1;
        if (meta == 0u) {
          if (sp > 0) {
            stack_unpack(stack[--sp], link, meta);
          } else {
            link = RTW_LINK_DONE;
          }
        }
      }
      tail += __popc(m);
      __syncwarp();
    }
    // ---- (d) drain: full batches always; the rest when enough lanes are parked or nothing else can progress -------------
    {
      parked = active && link == RTW_LINK_DONE && (int32_t)(head - my_last) < 0;
      const uint32_t m_parked = __ballot_sync(FULL, parked);
      const uint32_t m_prog = __ballot_sync(FULL, active && link != RTW_LINK_DONE);
      const bool force = m_parked != 0u && (__popc(m_parked) >= RTW_POOL_WAIT_LANES || m_prog == 0u);
      while (tail - head >= 32u || (force && tail != head)) {
        const uint32_t n = min(tail - head, 32u);
        const bool v = lane < n;
        const uint32_t e = v ? lp.queue[(head + lane) & 63u] : 0u;
        head += n;
        const uint32_t owner = e >> 27, slot = e & 0x7FFFFFFu;
        bool hit = false;
        unsigned long long key = 0ull;
        if (v) {
          v3 ro = mk(lp.ray[0][owner], lp.ray[1][owner], lp.ray[2][owner]);
          v3 rd = mk(lp.ray[3][owner], lp.ray[4][owner], lp.ray[5][owner]);
          const float rtime = lp.ray[6][owner], rt_min = lp.ray[7][owner];
          const float bt = ordered_to_float((uint32_t)(lp.best[owner] >> 32));
          const uint32_t pm = __ldg(sc.slot_meta + slot);
          const uint32_t type = pm & 7u, inst = pm >> RTW_META_TYPE_BITS;
          if (inst != 0) ray_to_instance(sc, inst, ro, rd);
          if (COUNT) {
            cnt.prims++;
            cnt.prim_bytes += 4u + ((type == PT_SPHERE) ? 16u : ((type >= PT_RECT_YZ && type <= PT_RECT_XY) ? 32u : 48u));
          }
          const float4* __restrict__ g = sc.geom + 3 * (size_t)slot;
          float t;
          hit = prim_t(type, g, ro, rd, rtime, rt_min, bt, t);
          if (hit) {
            key = ((unsigned long long)float_to_ordered(t) << 32) | (0xFFFFFFFEu - (uint32_t)__ldg(sc.slot_prim + slot));
            atomicMin(&lp.best[owner], key);
          }
        }
        __syncwarp();
        if (hit && lp.best[owner] == key) lp.best_slot[owner] = (int32_t)slot;
        __syncwarp();
        if (active) best_t = ordered_to_float((uint32_t)(lp.best[lane] >> 32));
      }
    }
    // ---- (e) finished queries -------------------------------------------------------------------------------------------
    if (active && link == RTW_LINK_DONE && (int32_t)(head - my_last) >= 0) {
      const int32_t bs = lp.best_slot[lane];
      io.store(index, o, d, time, bs, ordered_to_float((uint32_t)(lp.best[lane] >> 32)), bs >= 0 ? __ldg(sc.slot_meta + bs) : 0u);
      active = false;
    }
    parked = active && link == RTW_LINK_DONE;  // still waiting for its last entries
  }
}

// Flat scenes (SceneDev::flat_count > 0: at most 32 primitives in one leaf, e.g. the Cornell box): every ray tests
// every primitive, in the same order — there is nothing to walk and nothing to balance.  The primitive records are
// staged in shared memory once per block (48 B geometry + meta + canonical id) together with their RUNS: the build
// sorted the slots of a leaf by (instance, type), so the list is a handful of runs of equal instance and type.  The
// loop over runs is warp-uniform: the ray is transformed once per instance, the primitive type is dispatched once per
// run, and inside a run of rectangles every test divides by the same direction component (SharedDivisor) and is
// branch free.  Same tests, same order of the float operations, same tie rule as traverse_persistent.
struct FlatRecords {
  float4 g[32][3];
  uint32_t meta[32];
  int32_t prim[32];
  uint4 run[32];     // first slot, one past the last, primitive type, instance
  uint32_t nruns;
  uint32_t k_safe;   // every rectangle plane k passes plane_div_safe: a run may skip the per-numerator window test
};

__device__ __forceinline__ void stage_flat(const SceneDev& sc, FlatRecords& fr) {
  for (uint32_t i = threadIdx.x; i < sc.flat_count; i += blockDim.x) {
    fr.g[i][0] = sc.geom[3 * (size_t)i]; fr.g[i][1] = sc.geom[3 * (size_t)i + 1]; fr.g[i][2] = sc.geom[3 * (size_t)i + 2];
    fr.meta[i] = sc.slot_meta[i];
    fr.prim[i] = sc.slot_prim[i];
  }
  if (threadIdx.x == 0) {
    uint32_t nruns = 0, first = 0, k_safe = 1;
    for (uint32_t i = 1; i <= sc.flat_count; ++i) {
      const uint32_t pm0 = sc.slot_meta[first];
      if (i == sc.flat_count || sc.slot_meta[i] != pm0) {
        fr.run[nruns++] = make_uint4(first, i, pm0 & 7u, pm0 >> RTW_META_TYPE_BITS);
        first = i;
      }
    }
    for (uint32_t i = 0; i < sc.flat_count; ++i) {
      const uint32_t type = sc.slot_meta[i] & 7u;
      if (type >= PT_RECT_YZ && type <= PT_RECT_XY) {
        const float k = sc.geom[3 * (size_t)i + 1].x;
        if (!plane_div_safe(k)) k_safe = 0;
      }
    }
    fr.nruns = nruns;
    fr.k_safe = k_safe;
  }
  __syncthreads();
}

// One run of rectangles of one orientation: (oA, oB, oK) / (dA, dB, dK) = the ray permuted to (in-plane A, in-plane B,
// constant axis K) — rectangular.rs:27-57 / 78-108 / 129-159.  Every test divides by dK: one exact reciprocal for the
// run when the operands allow it (SharedDivisor), the plain division otherwise.
#ifndef RTW_FLAT_UNROLL
#define RTW_FLAT_UNROLL 1  // the fused kernel's hot code must stay inside the 32 KB instruction cache (r02 A/B)
#endif
constexpr int kFlatUnroll = RTW_FLAT_UNROLL;
template <class Accept>
__device__ __forceinline__ void flat_rect_run(const FlatRecords& fr, uint32_t first, uint32_t last, float oA, float oB, float oK,
                                              float dA, float dB, float dK, float t_min, const float& best_t, Accept&& accept) {
  SharedDivisor dv;
  dv.set(dK);
  const bool run_fast = dv.fast && fr.k_safe && origin_div_safe(oK);  // true for every lane in practice
#pragma unroll kFlatUnroll
  for (uint32_t k = first; k < last; ++k) {
    const float num = fr.g[k][1].x - oK;
    float t;
    if (run_fast) t = dv.div_in_window(num);
    else t = div_exact(num, dK);
    accept(rect_accept(t, oA, oB, dA, dB, t_min, best_t, fr.g[k][0]), t, k);
  }
}

// The closest hit of one ray over the staged records: every lane of the warp walks the list together (call with the
// whole warp; `active` = this lane holds a ray).  best_t comes in as the ray's t_max.
template <class IV>
__device__ __forceinline__ void flat_closest(const IV& iv, const FlatRecords& fr, v3 o, v3 d, float time, float t_min,
                                             bool active, float& best_t, int32_t& best_slot, uint32_t& best_meta) {
  best_slot = -1;
  best_meta = 0;
  int32_t best_prim = -1;
  uint32_t cur_inst = 0;
  v3 oi = o, di = d;
  // closest-so-far update with the list rule (hittable/mod.rs:57-69), branch free: a hit at t <= best_t replaces the
  // record; among bit-equal t the later primitive of the CANONICAL order wins (the slots are not in canonical order)
  auto accept = [&](bool hit, float t, uint32_t k) {
    const int32_t prim_k = fr.prim[k];
    const bool h = hit && active;
    const bool closer = h && (best_slot < 0 || t < best_t);
    const bool take = closer || (h && prim_k > best_prim);
    best_t = closer ? t : best_t;
    best_slot = take ? (int32_t)k : best_slot;
    best_prim = take ? prim_k : best_prim;
  };
  const uint32_t nruns = fr.nruns;
  for (uint32_t r = 0; r < nruns; ++r) {  // warp-uniform
    const uint4 run = fr.run[r];
    const uint32_t first = run.x, last = run.y, type = run.z, inst = run.w;
    if (inst != cur_inst) {
      oi = o; di = d;
      if (inst != 0) ray_to_instance_iv(iv, inst, oi, di);
      cur_inst = inst;
    }
    if (type == PT_RECT_XZ) {
      flat_rect_run(fr, first, last, oi.x, oi.z, oi.y, di.x, di.z, di.y, t_min, best_t, accept);
    } else if (type == PT_RECT_XY) {
      flat_rect_run(fr, first, last, oi.x, oi.y, oi.z, di.x, di.y, di.z, t_min, best_t, accept);
    } else if (type == PT_RECT_YZ) {
      flat_rect_run(fr, first, last, oi.y, oi.z, oi.x, di.y, di.z, di.x, t_min, best_t, accept);
    } else if (type <= PT_MSPHERE) {
      for (uint32_t k = first; k < last; ++k) {
        const float4 g0 = fr.g[k][0];
        v3 center = mk(g0.x, g0.y, g0.z);
        if (type == PT_MSPHERE) center = moving_center(g0, fr.g[k][1], fr.g[k][2], time);
        float t = 0.f;
        const bool hit = sphere_t(oi, di, t_min, best_t, center, g0.w, t);
        accept(hit, t, k);
      }
    } else {
      for (uint32_t k = first; k < last; ++k) {
        float t = 0.f, a, b;
        const bool hit = tri_t(oi, di, t_min, best_t, fr.g[k][0], fr.g[k][1], fr.g[k][2], t, a, b);
        accept(hit, t, k);
      }
    }
  }
  if (best_slot >= 0) best_meta = fr.meta[best_slot];
}

template <class IO>
__device__ __forceinline__ void traverse_flat(const SceneDev& sc, const FlatRecords& fr, IO& io, uint32_t count, uint32_t* cursor) {
  const uint32_t lane = threadIdx.x & 31;
  const GlobalInst iv{sc.inst_range, sc.inst_ops};
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cursor, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= count) break;
    const uint32_t index = base + lane;
    v3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    float time = 0.f, t_min = 0.f, best_t = 0.f;
    int32_t slot0 = -1;
    bool resumed = false;
    const bool active = index < count && io.load(index, o, d, time, t_min, best_t, slot0, resumed);
    int32_t best_slot;
    uint32_t best_meta;
    flat_closest(iv, fr, o, d, time, t_min, active, best_t, best_slot, best_meta);
    if (active) io.store(index, o, d, time, best_slot, best_t, best_meta);
  }
}

// Every primitive in canonical order: hittable/mod.rs:57-69 literally (closest_so_far shrink,
// later primitive wins on equal t).  The debug / ground-truth path of rtw_trace_closest.
__device__ __forceinline__ void brute_closest(const SceneDev& sc, v3 o, v3 d, float time, float t_min, float t_max,
                                              const Rng& rng, int32_t& best_id, float& best_t, uint32_t& best_meta) {
  best_id = -1;
  best_t = t_max;
  best_meta = 0;
  v3 oi = o, di = d;
  uint32_t cur_inst = 0;
  for (uint32_t id = 0; id < sc.num_prims; ++id) {
    uint32_t meta = __ldg(sc.prim_meta + id);
    uint32_t type = meta & 7u, inst = meta >> RTW_META_TYPE_BITS;
    if (inst != cur_inst) {
      oi = o; di = d;
      if (inst != 0) ray_to_instance(sc, inst, oi, di);
      cur_inst = inst;
    }
    float t;
    const float4* __restrict__ g = sc.raw_geom + 3 * (size_t)id;
    const bool hit = (type >= PT_MEDIUM_SPHERE) ? medium_t(type, g, oi, di, d, t_min, best_t, rng.gen_f32_keyed(id), t)
                                                : prim_t(type, g, oi, di, time, t_min, best_t, t);
    if (hit) {
      best_t = t; best_id = (int32_t)id; best_meta = meta;
    }
  }
}

}  // namespace rtw
