// Wavefront path tracer: the GPU form of Raytracer::render (lib.rs:57-117).
//
//   pool of P path slots (SoA in HBM) ──► k_wave_traverse ──► k_wave_shade ──► next queue
//
// * A slot owns one WORK ITEM = (pixel, sample slice): it runs the samples of that slice one after
//   the other (lib.rs:83-88), adding each path's radiance to a running sum in sample order, writes
//   the slice sum when done and pulls the next item from a global cursor (path regeneration), so
//   the pool stays full until the frame runs dry.  Slices of a pixel are added in slice order by
//   k_wave_resolve.  Every float addition therefore happens in an order fixed by (spp, slices)
//   alone — the image is bit-reproducible for any pool size, GPU count or scheduling.
// * sample_ray's recursion (lib.rs:97-117) is run in its iterative form L += T*e; T *= a
//   (SURVEY.md §8 a3); one iteration of the wavefront = one path segment per live slot.
// * Both kernels are persistent: grid = SMs x resident blocks, warps pull 32 entries at a time from
//   a device-side cursor, so no launch parameter depends on the live count and the host only
//   synchronises every few iterations to learn whether anything is left.
// * While work items remain every slot is busy, so entry i of an iteration simply IS slot i
//   ("identity" mode: coalesced state accesses, no queue, no compaction atomics).  Once the item
//   cursor has run dry the shade kernel compacts the surviving slots into a queue each iteration
//   ("queue" mode), so the tail of the frame costs time proportional to the live paths.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rtw_scene.cuh"
#include "rtw_traverse.cuh"

namespace rtw {

namespace {

#define RTW_MAX_SUBPOOLS 4
#ifndef RTW_SHADE_STATIC
#define RTW_SHADE_STATIC 0  // static trips + L2 prefetch of the next trip: Cornell shade -5 %, cow +7 % (imbalance): off
#endif
#ifndef RTW_WIDE_HIT
#define RTW_WIDE_HIT 0  // A/B r01: 16-byte hit record (meta + material): shade +-0, traversal +1-2 % slower
#endif
// ray_d.w of a slot: 0 = a live ray, else
#define RTW_SLOT_DEAD 1.0f     // no work left for the slot (end of the frame, identity mode)
#define RTW_SLOT_PENDING 2.0f  // traversal suspended in the drain of the last launch (rtw_traverse.cuh)

struct WaveCtl {           // one per sub-pool (+ one extra whose item_cursor is the shared work-item cursor)
  unsigned long long item_cursor;
  unsigned long long segments;
  unsigned long long paths;
  unsigned long long pairs;
  unsigned long long prims;
  unsigned long long prim_bytes;
  uint32_t count[2];       // entries of the iteration with that parity (identity mode: slot_count)
  uint32_t cursor_traverse;
  uint32_t cursor_shade;
  uint32_t qmode[2];       // 0 = identity (entry i is slot slot_base + i), 1 = queue[parity] lists the live slots
  uint32_t exhausted;      // (shared ctl only) the item cursor ran past n_items
  uint32_t pad[3];
};

struct WaveDev {
  float4* ray_o;    // origin.xyz, time
  float4* ray_d;    // direction.xyz, -
#if RTW_WIDE_HIT
  int4* hit;        // primitive slot (or -1), t bits, slot meta (type | inst << 3), material: what the shade kernel
                    // would otherwise fetch through the slot in a dependent load
#else
  int2* hit;        // primitive slot (or -1), t bits
#endif
  float4* thr;      // throughput T.rgb
  float4* sum;      // running slice sum rgb
  uint4* state;     // pixel index (row*w+col), sample, sample_end, bounce | slice << 8
  uint32_t* queue[2];
  float4* partial;  // [slices][pix_per_slice] slice sums of the pixels this partition owns (slices > 1)
  WaveCtl* ctl;                      // this sub-pool's counters
  unsigned long long* item_cursor;   // shared by all sub-pools
  uint32_t* exhausted;               // shared by all sub-pools
  uint32_t slot_base, slot_count;    // this sub-pool's slots: [slot_base, slot_base + slot_count)
};

struct FrameDev {
  CameraDev cam;
  v3 background;
  uint32_t width, height;
  uint32_t max_depth;
  uint32_t sample_begin, nsamp, slices;
  uint32_t tile_size, tiles_x, tiles_total, part_rank, part_count;
  uint32_t seed_lo, seed_hi;
  unsigned long long pix_per_slice;  // owned tiles * tile_size^2 (includes out-of-image padding)
  unsigned long long n_items;
  uint32_t pool;
  uint32_t fit32;       // n_items, nsamp * slices < 2^32: decode_item stays in 32-bit arithmetic
  uint32_t tile_shift;  // log2(tile_size) when it is a power of two, else 0xffffffff
};

// the traversal kernel variants a render can launch
enum TravKind { TK_PAIR = 0, TK_MEDIA, TK_WIDE, TK_COMPACT, TK_COUNT, TK_COUNT_COMPACT, TK_FLAT, TK_N };

struct WaveHost {
  WaveDev dev{};
  std::vector<void*> allocs;
  uint32_t pool = 0;
  size_t partial_elems = 0;
  WaveCtl* pinned_ctl = nullptr;   // [RTW_MAX_SUBPOOLS][2]: ring for the lagging termination check
  cudaStream_t stream = nullptr;    // internal non-blocking stream (graph capture needs a non-legacy stream)
  cudaStream_t pool_stream[RTW_MAX_SUBPOOLS] = {};
  int blocks_trav[TK_N] = {};  // persistent grid per TravKind
  int blocks_shade = 0;
};

// ---- work items ---------------------------------------------------------------------------------
struct Item {
  uint32_t pixel;  // row*w + col   (row = bottom-up like Pixel.row, lib.rs:58)
  uint32_t sample, sample_end, slice;
};

__device__ __forceinline__ bool decode_item(const FrameDev& f, unsigned long long n, Item& it) {
  uint32_t k, tile_local, within, tx, ty, wx, wy;
  if (f.fit32) {  // warp-uniform; 64-bit divisions cost ~100 instructions each
    const uint32_t n32 = (uint32_t)n, pps = (uint32_t)f.pix_per_slice;
    k = n32 / pps;
    const uint32_t q = n32 - k * pps;
    if (f.tile_shift != 0xffffffffu) {
      tile_local = q >> (2u * f.tile_shift);
      within = q & ((1u << (2u * f.tile_shift)) - 1u);
      wx = within & (f.tile_size - 1u);
      wy = within >> f.tile_shift;
    } else {
      const uint32_t tsq = f.tile_size * f.tile_size;
      tile_local = q / tsq; within = q - tile_local * tsq;
      wy = within / f.tile_size; wx = within - wy * f.tile_size;
    }
    it.sample = f.sample_begin + (f.nsamp * k) / f.slices;
    it.sample_end = f.sample_begin + (f.nsamp * (k + 1u)) / f.slices;
  } else {
    k = (uint32_t)(n / f.pix_per_slice);
    const unsigned long long q = n % f.pix_per_slice;
    const uint32_t tsq = f.tile_size * f.tile_size;
    tile_local = (uint32_t)(q / tsq); within = (uint32_t)(q % tsq);
    wy = within / f.tile_size; wx = within % f.tile_size;
    it.sample = f.sample_begin + (uint32_t)(((unsigned long long)f.nsamp * k) / f.slices);
    it.sample_end = f.sample_begin + (uint32_t)(((unsigned long long)f.nsamp * (k + 1)) / f.slices);
  }
  const unsigned long long tile = (unsigned long long)tile_local * f.part_count + f.part_rank;
  if (tile >= f.tiles_total) return false;
  ty = (uint32_t)tile / f.tiles_x; tx = (uint32_t)tile - ty * f.tiles_x;
  const uint32_t x = tx * f.tile_size + wx;
  const uint32_t y_top = ty * f.tile_size + wy;
  if (x >= f.width || y_top >= f.height) return false;
  const uint32_t row = f.height - 1 - y_top;
  it.pixel = row * f.width + x;
  it.slice = k;
  return it.sample_end > it.sample;
}

// Index of an owned pixel inside a slice = the `q` decode_item splits (tile-major over the owned tiles): the slice
// sums are stored compactly, [slices][pix_per_slice], whatever share of the frame this partition renders.
__device__ __forceinline__ unsigned long long owned_index(const FrameDev& f, uint32_t x, uint32_t y_top) {
  uint32_t tx, ty, wx, wy;
  if (f.tile_shift != 0xffffffffu) {
    tx = x >> f.tile_shift; ty = y_top >> f.tile_shift;
    wx = x & (f.tile_size - 1u); wy = y_top & (f.tile_size - 1u);
  } else {
    tx = x / f.tile_size; ty = y_top / f.tile_size;
    wx = x - tx * f.tile_size; wy = y_top - ty * f.tile_size;
  }
  const uint32_t tile_local = (ty * f.tiles_x + tx) / f.part_count;
  return (unsigned long long)tile_local * (f.tile_size * f.tile_size) + wy * f.tile_size + wx;
}

// Warp-cooperative fetch: every lane with `need` gets a valid item or learns that none are left.
// Must be called by all 32 lanes.
__device__ __forceinline__ bool fetch_item(const FrameDev& f, const WaveDev& w, bool need, Item& it) {
  unsigned long long* item_cursor = w.item_cursor;
  const uint32_t lane = threadIdx.x & 31;
  bool got = false;
  for (;;) {
    uint32_t m = __ballot_sync(0xffffffffu, need && !got);
    if (m == 0) break;
    unsigned long long base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(item_cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (base + __popc(m) > f.n_items) {  // ran dry (the cursor may overshoot; harmless): later iterations compact
      if (lane == 0) *w.exhausted = 1u;
      if (base >= f.n_items) break;
    }
    if (need && !got) {
      unsigned long long n = base + __popc(m & ((1u << lane) - 1u));
      if (n < f.n_items) got = decode_item(f, n, it);
      else need = false;
    }
  }
  return got;
}

// lib.rs:84-86: start sample `it.sample` of the item's pixel.
__device__ __forceinline__ void start_path(const FrameDev& f, const WaveDev& w, uint32_t slot, const Item& it) {
  Rng rng;
  rng.begin(((uint64_t)f.seed_hi << 32) | f.seed_lo, it.pixel, it.sample, 0);
  uint32_t row = it.pixel / f.width, col = it.pixel % f.width;
  v3 o, d;
  float time;
  camera_ray(f.cam, f.width, f.height, row, col, rng, o, d, time);
  w.ray_o[slot] = make_float4(o.x, o.y, o.z, time);
  w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
  w.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
  w.state[slot] = make_uint4(it.pixel, it.sample, it.sample_end, it.slice << 8);
}

__device__ __forceinline__ void queue_push(uint32_t* queue, uint32_t* count, bool alive, uint32_t slot) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t m = __ballot_sync(0xffffffffu, alive);
  if (m == 0) return;
  uint32_t base = 0;
  if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (alive) queue[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

__global__ void k_wave_init(SceneDev sc, FrameDev f, WaveDev w) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;  // block = 128 threads: whole warps reach the ballots
  const uint32_t slot = w.slot_base + tid;
  if (tid == 0) {  // iteration 0 reads the slots in identity mode
    w.ctl->count[0] = w.slot_count;
    w.ctl->qmode[0] = 0;
  }
  Item it;
  bool got = fetch_item(f, w, tid < w.slot_count, it);
  if (got) {
    start_path(f, w, slot, it);
    w.sum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (tid < w.slot_count) {
    w.ray_d[slot] = make_float4(0.f, 0.f, 0.f, RTW_SLOT_DEAD);
  }
  uint32_t m = __ballot_sync(0xffffffffu, got);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.ctl->paths, (unsigned long long)__popc(m));
}

// ---- traversal -----------------------------------------------------------------------------------
struct WaveIO {
  const WaveDev& w;
  const uint32_t* __restrict__ queue;  // nullptr: identity mode
  uint32_t slot;
  uint32_t seed_lo, seed_hi;
  // key of the current path segment: (pixel, sample), stage = bounce + 1 — the stage the shade kernel uses
  __device__ __forceinline__ void rng_key(Rng& rng) {
    const uint4 st = w.state[slot];
    rng.begin(((uint64_t)seed_hi << 32) | seed_lo, st.x, st.y, (st.w & 0xffu) + 1u);
  }
#ifndef RTW_SUSPEND_LANES
#define RTW_SUSPEND_LANES 0  // A/B r01 (cow, monument): 4 / 8 / 12 cost 3-12 % — the restarts outweigh the shorter drain
#endif
  static constexpr int kSuspendLanes = RTW_SUSPEND_LANES;
  bool was_pending;
  const int2* __restrict__ slot_ms;  // SceneDev::slot_ms
  __device__ __forceinline__ bool load(uint32_t i, v3& o, v3& d, float& time, float& t_min, float& t_max, int32_t& slot0,
                                       bool& resumed) {
    slot = queue ? queue[i] : w.slot_base + i;
    const float4 o4 = w.ray_o[slot], d4 = w.ray_d[slot];
    if (d4.w == RTW_SLOT_DEAD) return false;  // identity mode at the end of the frame
    o = mk(o4.x, o4.y, o4.z); d = mk(d4.x, d4.y, d4.z); time = o4.w;
    t_min = 0.001f; t_max = __int_as_float(0x7f800000);  // lib.rs:102: world.hit(r, 0.001, f32::INFINITY)
    resumed = was_pending = d4.w == RTW_SLOT_PENDING;
    if (resumed) {  // suspended by the previous launch with this hit
      const auto h = w.hit[slot];
      slot0 = h.x; t_max = __int_as_float(h.y);
    }
    return true;
  }
  __device__ __forceinline__ void store(uint32_t, v3, v3 d, float, int32_t hslot, float t, uint32_t meta) {
#if RTW_WIDE_HIT
    w.hit[slot] = make_int4(hslot, __float_as_int(t), (int)meta, hslot >= 0 ? slot_ms[hslot].x : 0);
#else
    w.hit[slot] = make_int2(hslot, __float_as_int(t));
#endif
    if (was_pending) w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
  }
  __device__ __forceinline__ void suspend(uint32_t, int32_t hslot, float t) {
#if RTW_WIDE_HIT
    w.hit[slot] = make_int4(hslot, __float_as_int(t), 0, 0);
#else
    w.hit[slot] = make_int2(hslot, __float_as_int(t));
#endif
    const float4 d4 = w.ray_d[slot];
    w.ray_d[slot] = make_float4(d4.x, d4.y, d4.z, RTW_SLOT_PENDING);
  }
};

// 8 resident blocks (64 registers) for the product variant: A/B r01 +2..3 % over the compiler's own choice (72)
#ifndef RTW_TRAVERSE_MINBLOCKS
#define RTW_TRAVERSE_MINBLOCKS 8
#endif
// Prologue of every traversal kernel of the wavefront: this iteration's entry count and input mode, and (one thread)
// the bookkeeping for the shade kernel that follows.
__device__ __forceinline__ void traverse_prologue(const WaveDev& w, uint32_t parity, uint32_t& count, uint32_t& in_queue) {
  WaveCtl* ctl = w.ctl;
  count = ctl->count[parity];
  in_queue = ctl->qmode[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->cursor_shade = 0;  // consumed by the shade kernel that follows
    // What the shade kernel that follows writes for the next iteration: a queue once the work items have run
    // out (decided HERE, between launches, so that all of its blocks agree), else nothing (identity).
    const uint32_t out_queue = (in_queue || *w.exhausted) ? 1u : 0u;
    ctl->qmode[parity ^ 1] = out_queue;
    ctl->count[parity ^ 1] = out_queue ? 0u : w.slot_count;
  }
}

// Flat scenes (<= 32 primitives in one leaf): rtw_traverse.cuh, traverse_flat
__global__ void __launch_bounds__(128, 8) k_wave_traverse_flat(SceneDev sc, WaveDev w, uint32_t parity, uint32_t seed_lo,
                                                               uint32_t seed_hi) {
  __shared__ FlatRecords fr;
  uint32_t count, in_queue;
  traverse_prologue(w, parity, count, in_queue);
  stage_flat(sc, fr);
  WaveIO io{w, in_queue ? w.queue[parity] : nullptr, 0, seed_lo, seed_hi, false, sc.slot_ms};
  traverse_flat(sc, fr, io, count, &w.ctl->cursor_traverse);
}

template <bool COUNT, bool MEDIA, int NODES>
__global__ void __launch_bounds__(128, (COUNT || MEDIA || NODES == NODES_WIDE) ? 1 : RTW_TRAVERSE_MINBLOCKS)
    k_wave_traverse(SceneDev sc, WaveDev w, uint32_t parity, uint32_t seed_lo, uint32_t seed_hi) {
  WaveCtl* ctl = w.ctl;
  uint32_t count, in_queue;
  traverse_prologue(w, parity, count, in_queue);
  const uint32_t lane = threadIdx.x & 31;
  TraverseCounters cnt;
  WaveIO io{w, in_queue ? w.queue[parity] : nullptr, 0, seed_lo, seed_hi, false, sc.slot_ms};
#if RTW_TOP_TREE > 0
  __shared__ float4 top_smem[4 * RTW_TOP_TREE];
  stage_top_tree(sc, top_smem);
  traverse_persistent<COUNT, MEDIA, NODES>(sc, io, count, &ctl->cursor_traverse, cnt, top_smem);
#else
  traverse_persistent<COUNT, MEDIA, NODES>(sc, io, count, &ctl->cursor_traverse, cnt);
#endif
  if (COUNT) {
    uint32_t p = cnt.pairs, q = cnt.prims, r = cnt.prim_bytes;
    for (int off = 16; off > 0; off >>= 1) {
      p += __shfl_xor_sync(0xffffffffu, p, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
      r += __shfl_xor_sync(0xffffffffu, r, off);
    }
    if (lane == 0) {
      atomicAdd(&ctl->pairs, (unsigned long long)p);
      atomicAdd(&ctl->prims, (unsigned long long)q);
      atomicAdd(&ctl->prim_bytes, (unsigned long long)r);
    }
  }
}

// ---- shade + scatter + regenerate -----------------------------------------------------------------
// A warp shades 32 queue entries per trip.  Paths that end in a trip (15 % of the lanes on Cornell) are NOT
// restarted in place: that ran the whole regeneration code (work-item decode, Philox, camera ray) with 2.3 of
// 32 lanes active — a quarter of the kernel's issued instructions (profiles/r01f_shade_lines.txt).  They go
// on a per-warp backlog in shared memory instead; once 32 are waiting the warp restarts them together.
#define RTW_BACKLOG 64
#define RTW_NEED_ITEM 0x80000000u

struct ShadeBacklog {  // per warp; SoA so that lane i touches bank i
  uint32_t slot[RTW_BACKLOG], pixel[RTW_BACKLOG], sample[RTW_BACKLOG], sample_end[RTW_BACKLOG], slice[RTW_BACKLOG];
};

// restart the paths of backlog entries [first, first + 32) (those with valid == true): lib.rs:83-86
#ifdef RTW_NOINLINE_REGEN
static __device__ __noinline__ void regenerate(
#else
__device__ __forceinline__ void regenerate(
#endif
const FrameDev& f, const WaveDev& w, ShadeBacklog& bl, uint32_t idx, bool valid,
                                           uint32_t* next_queue, uint32_t* next_count, bool out_queue,
                                           uint32_t& new_paths) {
  Item it;
  uint32_t slot = 0;
  bool need_item = false, go = false;
  if (valid) {
    slot = bl.slot[idx];
    it.pixel = bl.pixel[idx]; it.sample = bl.sample[idx]; it.sample_end = bl.sample_end[idx]; it.slice = bl.slice[idx];
    need_item = (it.slice & RTW_NEED_ITEM) != 0;
    go = !need_item;
  }
  __syncwarp();
  if (fetch_item(f, w, need_item, it)) {  // the slot finished its work item: pull the next one
    w.sum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    go = true;
  }
  if (go) {
    start_path(f, w, slot, it);
    new_paths++;
  } else if (valid) {
    w.ray_d[slot] = make_float4(0.f, 0.f, 0.f, RTW_SLOT_DEAD);  // no work left for this slot
  }
  if (out_queue) queue_push(next_queue, next_count, go, slot);
}

#ifdef RTW_SHADE_MINBLOCKS  // A/B r01: 6 (80 registers) and 7 (72) are 4 % and 11 % slower than the default (94)
__global__ void __launch_bounds__(128, RTW_SHADE_MINBLOCKS) k_wave_shade(
#else
__global__ void __launch_bounds__(128) k_wave_shade(
#endif
    SceneDev sc, FrameDev f, WaveDev w, float* __restrict__ accum,
                                                    uint32_t parity) {
  __shared__ ShadeBacklog backlog[4];
  ShadeBacklog& bl = backlog[threadIdx.x >> 5];
  WaveCtl* ctl = w.ctl;
  const uint32_t count = ctl->count[parity];
  const uint32_t* __restrict__ queue = ctl->qmode[parity] ? w.queue[parity] : nullptr;  // nullptr: identity
  const bool out_queue = ctl->qmode[parity ^ 1] != 0;
  uint32_t* next_queue = w.queue[parity ^ 1];
  uint32_t* next_count = &ctl->count[parity ^ 1];
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->cursor_traverse = 0;  // for the next traversal launch
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lane_lt = (1u << lane) - 1u;
  const uint64_t seed = ((uint64_t)f.seed_hi << 32) | f.seed_lo;
  uint32_t new_paths = 0, nseg = 0;
  uint32_t nback = 0;  // warp-uniform: entries waiting on the backlog
#if RTW_SHADE_STATIC
  // Identity mode: trips are handed out statically (warp w takes trips w, w + W, ...), so the NEXT trip's slots are
  // known and their state can be started on its way into L2 while this trip is shaded.
  const bool static_trips = queue == nullptr;
  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  uint32_t trip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
#else
  const bool static_trips = false;
  uint32_t trip = 0;
  const uint32_t warps_total = 0;
#endif
  uint32_t grabbed = 0;
  if (!static_trips && lane == 0) grabbed = atomicAdd(&ctl->cursor_shade, 32u);
  for (;;) {
    const uint32_t base = static_trips ? trip * 32u : __shfl_sync(0xffffffffu, grabbed, 0);
    if (base >= count) break;
#if RTW_SHADE_STATIC
    if (static_trips) {
      trip += warps_total;
      const uint32_t nxt = trip * 32u + lane;
      if (nxt < count) {
        const uint32_t ns = w.slot_base + nxt;
        prefetch_l2(&w.hit[ns]); prefetch_l2(&w.ray_o[ns]); prefetch_l2(&w.ray_d[ns]);
        prefetch_l2(&w.thr[ns]); prefetch_l2(&w.state[ns]); prefetch_l2(&w.sum[ns]);
      }
    }
#endif
#ifdef RTW_CURSOR_PREFETCH
    if (lane == 0) grabbed = atomicAdd(&ctl->cursor_shade, 32u);
#endif
    const uint32_t i = base + lane;
    bool active = i < count;
    uint32_t slot = 0;
    bool alive = false;  // slot continues into the next iteration
    bool ended = false;  // the path ended: restart the slot (next sample of its item, or a new item)
    Item it;
    it.pixel = it.sample = it.sample_end = it.slice = 0;
    float4 o4, d4, T4, s4;
#if RTW_WIDE_HIT
    int4 h;
#else
    int2 h;
#endif
    uint4 st;
    if (active) {  // every load of the slot's state is issued before the first use
      slot = queue ? queue[i] : w.slot_base + i;
      h = w.hit[slot];
      o4 = w.ray_o[slot]; d4 = w.ray_d[slot];
      T4 = w.thr[slot];
      st = w.state[slot];
      s4 = w.sum[slot];
      alive = d4.w == RTW_SLOT_PENDING;  // traversal not finished: carried to the next iteration unshaded
      active = d4.w == 0.f;              // RTW_SLOT_DEAD: identity mode at the end of the frame
    }
    if (active) {
      nseg++;
      v3 T = mk(T4.x, T4.y, T4.z);
      uint32_t bounce = st.w & 0xffu;
      v3 L = mk(0.f, 0.f, 0.f);
      if (h.x < 0) {  // lib.rs:102-105: miss -> background
        L = T * f.background;
        ended = true;
      } else {
#if RTW_WIDE_HIT
        const uint32_t meta = (uint32_t)h.z;
        const MaterialRec m = sc.materials[h.w];
        // only a triangle with per-vertex normals / uvs needs its TriShade index (one more hop through the slot)
        int2 ms = make_int2(h.w, -1);
        if ((meta & 7u) == PT_TRI && sc.has_tri_shade) ms.y = sc.slot_ms[h.x].y;
#else
        const uint32_t meta = sc.slot_meta[h.x];
        const int2 ms = sc.slot_ms[h.x];
        const MaterialRec m = sc.materials[ms.x];
#endif
        const v3 o = mk(o4.x, o4.y, o4.z), d = mk(d4.x, d4.y, d4.z);
        const bool need_uv = !m.solid && (m.type == MT_LAMBERTIAN || m.type == MT_DIFFUSE_LIGHT) && texture_needs_uv(sc, m.tex);
        HitRec rec;
        finalize_hit(sc, meta & 7u, meta >> RTW_META_TYPE_BITS, sc.geom + 3 * (size_t)h.x, ms.y, o, d, o4.w,
                     __int_as_float(h.y), need_uv, rec);
        const v3 emitted = material_emitted(sc, m, rec);  // lib.rs:107-109
        Rng rng;
        rng.begin(seed, st.x, st.y, bounce + 1);
        v3 att, out_dir;
        if (!material_scatter(sc, m, d, rec, rng, att, out_dir)) {  // lib.rs:111-114
          L = T * emitted;
          ended = true;
        } else {
          // L += T*emitted with emitted == 0 for every scattering material: exact no-op
          T = T * att;  // lib.rs:116
          bounce += 1;
          ended = bounce >= f.max_depth;  // lib.rs:98-100: depth exhausted -> black
          if (!ended) {
            w.ray_o[slot] = make_float4(rec.p.x, rec.p.y, rec.p.z, o4.w);
            w.ray_d[slot] = make_float4(out_dir.x, out_dir.y, out_dir.z, 0.f);
            w.thr[slot] = make_float4(T.x, T.y, T.z, 0.f);
            w.state[slot] = make_uint4(st.x, st.y, st.z, (st.w & ~0xffu) | bounce);
            alive = true;
          }
        }
      }
      if (ended) {
        v3 sum = mk(s4.x, s4.y, s4.z) + L;  // lib.rs:87: pixel_color += sample_ray(..)
        it.pixel = st.x; it.sample = st.y + 1; it.sample_end = st.z; it.slice = st.w >> 8;
        if (st.y + 1 < st.z) {  // next sample of the same item
          w.sum[slot] = make_float4(sum.x, sum.y, sum.z, 0.f);
        } else {  // item done: publish the slice sum
          uint32_t row = st.x / f.width, col = st.x - row * f.width;
          size_t pix = (size_t)(f.height - 1 - row) * f.width + col;
          if (f.slices == 1) {
            accum[3 * pix] = sum.x; accum[3 * pix + 1] = sum.y; accum[3 * pix + 2] = sum.z;
          } else {
            w.partial[(size_t)(st.w >> 8) * f.pix_per_slice + owned_index(f, col, f.height - 1 - row)] =
                make_float4(sum.x, sum.y, sum.z, 0.f);
          }
          it.slice = RTW_NEED_ITEM;
        }
      }
    }
    if (out_queue) queue_push(next_queue, next_count, alive, slot);
    const uint32_t m_end = __ballot_sync(0xffffffffu, ended);
    if (ended) {
      const uint32_t pos = nback + __popc(m_end & lane_lt);
      bl.slot[pos] = slot; bl.pixel[pos] = it.pixel; bl.sample[pos] = it.sample; bl.sample_end[pos] = it.sample_end;
      bl.slice[pos] = it.slice;
    }
    nback += __popc(m_end);
    __syncwarp();
#ifndef RTW_REGEN_THRESHOLD
#define RTW_REGEN_THRESHOLD 32
#endif
#ifndef RTW_CURSOR_PREFETCH
    if (!static_trips && lane == 0) grabbed = atomicAdd(&ctl->cursor_shade, 32u);
#endif
    if (nback >= RTW_REGEN_THRESHOLD) {
      const uint32_t take = min(nback, 32u);
      nback -= take;
      regenerate(f, w, bl, nback + lane, lane < take, next_queue, next_count, out_queue, new_paths);
    }
  }
  if (nback > 0) regenerate(f, w, bl, lane, lane < nback, next_queue, next_count, out_queue, new_paths);
  for (int off = 16; off > 0; off >>= 1) {
    new_paths += __shfl_xor_sync(0xffffffffu, new_paths, off);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, off);
  }
  if (lane == 0 && new_paths) atomicAdd(&ctl->paths, (unsigned long long)new_paths);
  if (lane == 0 && nseg) atomicAdd(&ctl->segments, (unsigned long long)nseg);  // one world.hit per live entry (lib.rs:102)
}

// pixel_color = sum over slices, in slice order; pixels of other partitions = 0
__global__ void k_wave_resolve(FrameDev f, const float4* __restrict__ partial, float* __restrict__ accum) {
  size_t npix = (size_t)f.width * f.height;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)(pix % f.width), y_top = (uint32_t)(pix / f.width);
    uint32_t tile = (y_top / f.tile_size) * f.tiles_x + (x / f.tile_size);
    v3 acc = mk(0.f, 0.f, 0.f);
    if (tile % f.part_count == f.part_rank) {
      const unsigned long long q = owned_index(f, x, y_top);
      for (uint32_t k = 0; k < f.slices; ++k) {
        float4 s = partial[(size_t)k * f.pix_per_slice + q];
        acc = acc + mk(s.x, s.y, s.z);
      }
    }
    accum[3 * pix] = acc.x; accum[3 * pix + 1] = acc.y; accum[3 * pix + 2] = acc.z;
  }
}

// console_app/src/main.rs:73-86
__global__ void k_resolve_rgb8(const float* __restrict__ accum, size_t n, float scale, uint8_t* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float c = sqrtf(scale * accum[i]);
    float cl = c;
    if (cl < 0.0f) cl = 0.0f;
    if (cl > 0.999f) cl = 0.999f;
    float v = 255.999f * cl;
    uint32_t b = (v != v) ? 0u : __float2uint_rz(v);  // `as u8`: saturating, NaN -> 0
    out[i] = (uint8_t)min(b, 255u);
  }
}

template <class T>
int wave_alloc(WaveHost* wh, T** out, size_t count) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(RTW_ERR_NOMEM, std::string("cudaMalloc (wavefront state) failed: ") + cudaGetErrorString(e));
  }
  wh->allocs.push_back(p);
  *out = (T*)p;
  return RTW_OK;
}

void wave_release(WaveHost* wh) {
  for (void* p : wh->allocs) cudaFree(p);
  wh->allocs.clear();
  wh->pool = 0;
  wh->partial_elems = 0;
}

}  // namespace

void free_wave(rtw_scene* s) {
  WaveHost* wh = (WaveHost*)s->wave;
  if (!wh) return;
  wave_release(wh);
  if (wh->pinned_ctl) cudaFreeHost(wh->pinned_ctl);
  if (wh->stream) cudaStreamDestroy(wh->stream);
  for (auto& ps : wh->pool_stream)
    if (ps) cudaStreamDestroy(ps);
  delete wh;
  s->wave = nullptr;
}

int resolve_rgb8_device(const float* d_accum, size_t n, uint32_t spp, uint8_t* d_rgb8, cudaStream_t st) {
  if (n == 0) return RTW_OK;
  float scale = 1.0f / (float)spp;
  uint32_t blocks = (uint32_t)std::min<size_t>((n + 255) / 256, 65535);
  k_resolve_rgb8<<<blocks, 256, 0, st>>>(d_accum, n, scale, d_rgb8);
  RTW_CUDA_TRY(cudaGetLastError());
  return RTW_OK;
}

int render_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* d_accum, cudaStream_t st,
                  rtw_render_stats* stats) {
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  if (p->width < 2 || p->height < 2) return set_error(RTW_ERR_INVALID, "render: width and height must be >= 2");
  if ((uint64_t)p->width * p->height > 0xFFFFFFFFull) return set_error(RTW_ERR_INVALID, "render: image too large");
  uint32_t s0 = p->sample_begin, s1 = p->sample_end;
  if (s0 == 0 && s1 == 0) s1 = p->spp;
  if (s1 < s0) return set_error(RTW_ERR_INVALID, "render: sample_end < sample_begin");
  const uint32_t nsamp = s1 - s0;
  const size_t npix = (size_t)p->width * p->height;

  FrameDev f;
  memset(&f, 0, sizeof(f));
  auto V = [](const float* a) { return mk(a[0], a[1], a[2]); };
  f.cam.origin = V(cam->origin);
  f.cam.lower_left_corner = V(cam->lower_left_corner);
  f.cam.horizontal = V(cam->horizontal);
  f.cam.vertical = V(cam->vertical);
  f.cam.u = V(cam->u);
  f.cam.v = V(cam->v);
  f.cam.lens_radius = cam->lens_radius;
  f.cam.time0 = cam->time0;
  f.cam.time1 = cam->time1;
  f.background = V(p->background);
  f.width = p->width;
  f.height = p->height;
  f.max_depth = p->max_depth ? p->max_depth : 50;
  if (f.max_depth > 255) return set_error(RTW_ERR_INVALID, "render: max_depth must be <= 255");
  f.sample_begin = s0;
  f.nsamp = nsamp;
  f.tile_size = p->tile_size ? p->tile_size : 32;
  f.tiles_x = (p->width + f.tile_size - 1) / f.tile_size;
  const uint32_t tiles_y = (p->height + f.tile_size - 1) / f.tile_size;
  f.tiles_total = f.tiles_x * tiles_y;
  f.part_count = p->part_count ? p->part_count : 1;
  f.part_rank = p->part_rank;
  if (f.part_rank >= f.part_count) return set_error(RTW_ERR_INVALID, "render: part_rank >= part_count");
  f.seed_lo = (uint32_t)p->seed;
  f.seed_hi = (uint32_t)(p->seed >> 32);
  const uint32_t owned_tiles = f.tiles_total > f.part_rank ? (f.tiles_total - f.part_rank + f.part_count - 1) / f.part_count : 0;
  f.pix_per_slice = (unsigned long long)owned_tiles * f.tile_size * f.tile_size;

  // Default pool (measured, profiles/r01_sweeps.txt): every launch of the persistent traversal ends with a drain
  // in which the SMs wait for the longest rays (~13 us flat scene, ~67 us cow), so launches should be few and
  // large; the slot state is streamed coalesced (identity mode), it does not need to stay in L2.  2^22 slots
  // for a flat scene, 2^23 with a hierarchy (104 B per slot: 0.4 / 0.9 GB).
  uint32_t pool = p->pool_size ? p->pool_size : ((s->dev.num_prims <= 32) ? (1u << 22) : (1u << 23));
  pool = (pool + 31u) & ~31u;
  uint32_t slices = p->slices;
  if (slices == 0) {  // enough items that the last ones to finish are a small fraction of the frame
    unsigned long long want = 16ull * pool;
    unsigned long long per = std::max<unsigned long long>(f.pix_per_slice, 1);
    // as many slices as it takes for 16 items per slot — a slot runs the samples of its item one after the other,
    // so coarse items leave a large pool idle behind a few long chains (8-way partition of Cornell: 105 ms at 64
    // slices vs 72 ms ideal) — bounded by the slice-sum buffer (16 B per item, <= 4 GiB)
    const unsigned long long cap = std::max<unsigned long long>((256ull << 20) / per, 1ull);
    slices = (uint32_t)std::min<unsigned long long>((want + per - 1) / per, cap);
  }
  slices = std::max(1u, std::min(slices, std::max(nsamp, 1u)));
  if (slices > 0xFFFFFFu) return set_error(RTW_ERR_INVALID, "render: too many slices");
  f.slices = slices;
  f.n_items = (nsamp == 0) ? 0ull : f.pix_per_slice * slices;
  pool = (uint32_t)std::min<unsigned long long>(pool, std::max<unsigned long long>((f.n_items + 31ull) & ~31ull, 32ull));
  f.pool = pool;
  f.fit32 = (f.n_items < (1ull << 32) && (unsigned long long)nsamp * (slices + 1ull) < (1ull << 32)) ? 1u : 0u;
  f.tile_shift = 0xffffffffu;
  if ((f.tile_size & (f.tile_size - 1u)) == 0u) {
    f.tile_shift = 0;
    while ((1u << f.tile_shift) < f.tile_size) f.tile_shift++;
  }

  // ---- scratch (kept on the scene between calls) -------------------------------------------------
  WaveHost* wh = (WaveHost*)s->wave;
  if (!wh) {
    wh = new WaveHost();
    s->wave = wh;
    RTW_CUDA_TRY(cudaMallocHost((void**)&wh->pinned_ctl, 2 * (RTW_MAX_SUBPOOLS + 1) * sizeof(WaveCtl)));
    RTW_CUDA_TRY(cudaStreamCreateWithFlags(&wh->stream, cudaStreamNonBlocking));
    for (auto& ps : wh->pool_stream) RTW_CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    int nb = 0;
    const void* kernels[TK_N] = {(const void*)k_wave_traverse<false, false, NODES_PAIR>, (const void*)k_wave_traverse<false, true, NODES_PAIR>,
                                 (const void*)k_wave_traverse<false, false, NODES_WIDE>, (const void*)k_wave_traverse<false, false, NODES_COMPACT>,
                                 (const void*)k_wave_traverse<true, true, NODES_PAIR>, (const void*)k_wave_traverse<true, true, NODES_COMPACT>,
                                 (const void*)k_wave_traverse_flat};
    for (int k = 0; k < TK_N; ++k) {
      RTW_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernels[k], 128, 0));
      wh->blocks_trav[k] = std::max(nb, 1) * s->num_sms;
    }
    RTW_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wave_shade, 128, 0));
    wh->blocks_shade = std::max(nb, 1) * s->num_sms;
  }
  const size_t partial_elems = slices > 1 ? (size_t)slices * f.pix_per_slice : 0;
  if (wh->pool != pool || wh->partial_elems < partial_elems) {
    wave_release(wh);
    int rc;
    WaveDev& w = wh->dev;
    if ((rc = wave_alloc(wh, &w.ray_o, pool))) return rc;
    if ((rc = wave_alloc(wh, &w.ray_d, pool))) return rc;
    if ((rc = wave_alloc(wh, &w.hit, pool))) return rc;
    if ((rc = wave_alloc(wh, &w.thr, pool))) return rc;
    if ((rc = wave_alloc(wh, &w.sum, pool))) return rc;
    if ((rc = wave_alloc(wh, &w.state, pool))) return rc;
    if ((rc = wave_alloc(wh, &w.queue[0], pool))) return rc;
    if ((rc = wave_alloc(wh, &w.queue[1], pool))) return rc;
    if ((rc = wave_alloc(wh, &w.partial, partial_elems))) return rc;
    if ((rc = wave_alloc(wh, &w.ctl, RTW_MAX_SUBPOOLS + 1))) return rc;
    wh->pool = pool;
    wh->partial_elems = partial_elems;
  }
  WaveDev& w = wh->dev;

  // The 4-wide walk (SceneDev::nodes4) halves the dependent node fetches but moves MORE bytes (it fetches the boxes
  // below a child whose own box the ray misses).  Measured r01: no gain on C5 (257 vs 253 ms) — that traversal sits at
  // the random-gather bandwidth of the memory system (tools/gather_peak.cu: 1.3 TB/s beyond L2), not at its latency —
  // and 5-17 % slower on cache-resident scenes.  Off unless RTW_WIDE=1 (parity-tested).
  bool wide = false;
  if (const char* e = getenv("RTW_WIDE"))
    wide = atoi(e) != 0 && !s->dev.has_media && 3u * (s->bvh_height / 2u + 1u) + 2u <= RTW_STACK_SIZE;
  const bool count_trav = (p->flags & RTW_RENDER_COUNT_TRAVERSAL) != 0;
  const bool time_kernels = (p->flags & 2u) != 0;
  // compact pairs exist only when rtw_build found the hierarchy too large for the caches (rtw_bvh.cu)
  const bool compact = s->dev.nodes_c != nullptr && !s->dev.has_media;
  // a scene that is one leaf (<= 32 primitives, no media) is not walked at all (RTW_FLAT=0: use the general kernel)
  bool flat = s->dev.flat_count > 0 && !s->dev.has_media;
  if (const char* e = getenv("RTW_FLAT")) flat = flat && atoi(e) != 0;
  const TravKind kind = count_trav ? (compact ? TK_COUNT_COMPACT : TK_COUNT)
                                   : (s->dev.has_media ? TK_MEDIA : (flat ? TK_FLAT : (compact ? TK_COMPACT : (wide ? TK_WIDE : TK_PAIR))));
  auto launch_traverse = [&](cudaStream_t sk, const WaveDev& wd, uint32_t parity, int grid) {
    switch (kind) {
      case TK_PAIR: k_wave_traverse<false, false, NODES_PAIR><<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
      case TK_MEDIA: k_wave_traverse<false, true, NODES_PAIR><<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
      case TK_WIDE: k_wave_traverse<false, false, NODES_WIDE><<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
      case TK_COMPACT: k_wave_traverse<false, false, NODES_COMPACT><<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
      case TK_COUNT: k_wave_traverse<true, true, NODES_PAIR><<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
      case TK_FLAT: k_wave_traverse_flat<<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
      default: k_wave_traverse<true, true, NODES_COMPACT><<<grid, 128, 0, sk>>>(s->dev, wd, parity, f.seed_lo, f.seed_hi); break;
    }
  };
  // All work runs on an internal stream ordered after the caller's stream; the call returns only after
  // that stream has drained, so the caller's stream order is preserved on both sides.
  cudaStream_t user_stream = st;
  st = wh->stream;
  cudaEvent_t ev_in, ev_begin, ev_end;
  RTW_CUDA_TRY(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
  RTW_CUDA_TRY(cudaEventRecord(ev_in, user_stream));
  RTW_CUDA_TRY(cudaStreamWaitEvent(st, ev_in, 0));
  RTW_CUDA_TRY(cudaEventCreate(&ev_begin));
  RTW_CUDA_TRY(cudaEventCreate(&ev_end));
  std::vector<cudaEvent_t> kev;  // begin/end pairs: traverse, shade, traverse, shade ...
  uint32_t launches = 0, iterations = 0;

  // Sub-pools: the slot range is split into K independent halves/quarters, each with its own queues,
  // counters and stream.  Their kernels depend only on their own predecessor, so the ramp-up and the
  // tail of one sub-pool's persistent kernel (~15 us per launch, measured by the pool-size sweep in
  // profiles/) overlap the body of another's, and latency-bound shading overlaps issue-bound traversal.
  // (r01: K = 2 gained 4 % while every iteration compacted into queues; with identity slots and 2^22+ pools one
  // pool is fastest — K = 1 / 2 / 3: cow 3507 / 3454 / 3347 Mrays/s, Cornell 7124 / 7113 / 6840.  Kept for experiments.)
  uint32_t K = 1;
  if (!count_trav && !time_kernels && pool >= (1u << 16)) {
    if (const char* e = getenv("RTW_SUBPOOLS")) K = (uint32_t)std::min(std::max(atoi(e), 1), RTW_MAX_SUBPOOLS);
  }
  WaveDev wk[RTW_MAX_SUBPOOLS];
  {
    uint32_t per = ((pool / K) + 31u) & ~31u, base = 0;
    for (uint32_t k = 0; k < K; ++k) {
      wk[k] = w;
      wk[k].ctl = w.ctl + k;
      wk[k].item_cursor = &w.ctl[RTW_MAX_SUBPOOLS].item_cursor;
      wk[k].exhausted = &w.ctl[RTW_MAX_SUBPOOLS].exhausted;
      wk[k].slot_base = base;
      wk[k].slot_count = (k + 1 == K) ? pool - base : std::min(per, pool - base);
      wk[k].queue[0] = w.queue[0] + base;
      wk[k].queue[1] = w.queue[1] + base;
      base += wk[k].slot_count;
    }
  }
  RTW_CUDA_TRY(cudaEventRecord(ev_begin, st));
  RTW_CUDA_TRY(cudaMemsetAsync(w.ctl, 0, (RTW_MAX_SUBPOOLS + 1) * sizeof(WaveCtl), st));
  if (slices == 1) RTW_CUDA_TRY(cudaMemsetAsync(d_accum, 0, npix * 3 * sizeof(float), st));
  if (f.n_items > 0) {
    if (!count_trav && !time_kernels) {
      // ---- product path: per sub-pool a CUDA graph of BATCH iterations, launched back to back; the host
      // looks at the live count of round i-1 while round i is already running.
      // BATCH is even (the queue parity is back to 0 after every batch).  The termination check lags one round, so a
      // frame runs up to 2 x BATCH - 1 empty iterations at its end: 16 for long frames, 4 for short ones (fewer than
      // two work items per slot: a ~50-iteration frame of a few ms, where 31 empty launch pairs would be a third of it).
      const int BATCH = (f.n_items >= 2ull * pool) ? 16 : 4;
      // a small pool does not need the full persistent grid: fewer blocks launch (and drain) faster
      const int need_blocks = (int)((pool + 127u) / 128u);
      int grid_t = std::min(wh->blocks_trav[kind], std::max(need_blocks, 1));
      int grid_s = std::min(wh->blocks_shade, std::max(need_blocks, 1));
  
      cudaEvent_t ev_fork, ev_join[RTW_MAX_SUBPOOLS], ring_ev[RTW_MAX_SUBPOOLS][2];
      cudaGraph_t graph[RTW_MAX_SUBPOOLS];
      cudaGraphExec_t exec[RTW_MAX_SUBPOOLS];
      RTW_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      RTW_CUDA_TRY(cudaEventRecord(ev_fork, st));
      for (uint32_t k = 0; k < K; ++k) {
        cudaStream_t sk = wh->pool_stream[k];
        RTW_CUDA_TRY(cudaEventCreateWithFlags(&ev_join[k], cudaEventDisableTiming));
        for (auto& e : ring_ev[k]) RTW_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        RTW_CUDA_TRY(cudaStreamWaitEvent(sk, ev_fork, 0));
        k_wave_init<<<(wk[k].slot_count + 127) / 128, 128, 0, sk>>>(s->dev, f, wk[k]);
        launches++;
        RTW_CUDA_TRY(cudaStreamBeginCapture(sk, cudaStreamCaptureModeThreadLocal));
        for (int b = 0; b < BATCH; ++b) {
          launch_traverse(sk, wk[k], (uint32_t)(b & 1), grid_t);
          k_wave_shade<<<grid_s, 128, 0, sk>>>(s->dev, f, wk[k], d_accum, (uint32_t)(b & 1));
        }
        RTW_CUDA_TRY(cudaStreamEndCapture(sk, &graph[k]));
        RTW_CUDA_TRY(cudaGraphInstantiate(&exec[k], graph[k], 0));
      }
      bool done[RTW_MAX_SUBPOOLS] = {false, false, false, false};
      for (uint32_t i = 0;; ++i) {
        for (uint32_t k = 0; k < K; ++k) {
          if (done[k]) continue;
          cudaStream_t sk = wh->pool_stream[k];
          RTW_CUDA_TRY(cudaGraphLaunch(exec[k], sk));
          RTW_CUDA_TRY(cudaMemcpyAsync(&wh->pinned_ctl[2 * k + (i & 1)], w.ctl + k, sizeof(WaveCtl), cudaMemcpyDeviceToHost, sk));
          RTW_CUDA_TRY(cudaEventRecord(ring_ev[k][i & 1], sk));
          launches += 2 * BATCH;
        }
        iterations += BATCH;
        if (i >= 1) {
          bool all = true;
          for (uint32_t k = 0; k < K; ++k) {
            if (!done[k]) {
              RTW_CUDA_TRY(cudaEventSynchronize(ring_ev[k][(i - 1) & 1]));
              if (wh->pinned_ctl[2 * k + ((i - 1) & 1)].count[0] == 0) done[k] = true;  // absorbing: queue mode, no live path, no item left
            }
            all = all && done[k];
          }
          if (all) break;
        }
      }
      for (uint32_t k = 0; k < K; ++k) {
        RTW_CUDA_TRY(cudaEventRecord(ev_join[k], wh->pool_stream[k]));
        RTW_CUDA_TRY(cudaStreamWaitEvent(st, ev_join[k], 0));
        cudaEventDestroy(ev_join[k]);
        for (auto& e : ring_ev[k]) cudaEventDestroy(e);
        cudaGraphExecDestroy(exec[k]);
        cudaGraphDestroy(graph[k]);
      }
      cudaEventDestroy(ev_fork);
    } else {
      // ---- instrumented path (traversal counters / per-kernel CUDA events): one pool, plain launches
      k_wave_init<<<(pool + 127) / 128, 128, 0, st>>>(s->dev, f, wk[0]);
      launches++;
      uint32_t parity = 0;
      const int batch = 8;
      for (;;) {
        for (int b = 0; b < batch; ++b) {
          if (time_kernels) {
            cudaEvent_t e4[4];
            for (auto& e : e4) { RTW_CUDA_TRY(cudaEventCreate(&e)); kev.push_back(e); }
            RTW_CUDA_TRY(cudaEventRecord(e4[0], st));
          }
          launch_traverse(st, wk[0], parity, wh->blocks_trav[kind]);
          if (time_kernels) {
            RTW_CUDA_TRY(cudaEventRecord(kev[kev.size() - 3], st));
            RTW_CUDA_TRY(cudaEventRecord(kev[kev.size() - 2], st));
          }
          k_wave_shade<<<wh->blocks_shade, 128, 0, st>>>(s->dev, f, wk[0], d_accum, parity);
          if (time_kernels) RTW_CUDA_TRY(cudaEventRecord(kev[kev.size() - 1], st));
          parity ^= 1;
          launches += 2;
          iterations++;
        }
        RTW_CUDA_TRY(cudaMemcpyAsync(wh->pinned_ctl, w.ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, st));
        RTW_CUDA_TRY(cudaStreamSynchronize(st));
        if (wh->pinned_ctl->count[parity] == 0) break;
      }
    }
  }
  if (slices > 1 || f.n_items == 0) {
    if (f.n_items == 0 && slices > 1) RTW_CUDA_TRY(cudaMemsetAsync(w.partial, 0, partial_elems * sizeof(float4), st));
    if (slices > 1) {
      k_wave_resolve<<<std::min<uint32_t>((uint32_t)((npix + 255) / 256), 8u * (uint32_t)s->num_sms), 256, 0, st>>>(f, w.partial, d_accum);
      launches++;
    }
  }
  RTW_CUDA_TRY(cudaGetLastError());
  RTW_CUDA_TRY(cudaEventRecord(ev_end, st));
  RTW_CUDA_TRY(cudaMemcpyAsync(wh->pinned_ctl, w.ctl, RTW_MAX_SUBPOOLS * sizeof(WaveCtl), cudaMemcpyDeviceToHost, st));
  RTW_CUDA_TRY(cudaStreamSynchronize(st));
  WaveCtl total;
  memset(&total, 0, sizeof(total));
  for (uint32_t k = 0; k < K; ++k) {
    total.segments += wh->pinned_ctl[k].segments;
    total.paths += wh->pinned_ctl[k].paths;
    total.pairs += wh->pinned_ctl[k].pairs;
    total.prims += wh->pinned_ctl[k].prims;
    total.prim_bytes += wh->pinned_ctl[k].prim_bytes;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev_begin, ev_end);
  cudaEventDestroy(ev_in);
  cudaEventDestroy(ev_begin);
  cudaEventDestroy(ev_end);
  float ms_t = 0.f, ms_s = 0.f;
  for (size_t i = 0; i + 3 < kev.size(); i += 4) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, kev[i], kev[i + 1]);
    cudaEventElapsedTime(&b, kev[i + 2], kev[i + 3]);
    ms_t += a;
    ms_s += b;
  }
  for (auto& e : kev) cudaEventDestroy(e);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->segments = total.segments;
    stats->paths = total.paths;
    stats->node_visits = total.pairs;
    stats->prim_tests = total.prims;
    stats->prim_bytes = total.prim_bytes;
    stats->iterations = iterations;
    stats->launches = launches;
    stats->pool_size = pool;
    stats->slices = slices;
    stats->ms_render = ms;
    stats->ms_traverse = ms_t;
    stats->ms_shade = ms_s;
    stats->node_record_bytes = compact ? 32.f : (wide ? 128.f : 64.f);
  }
  return RTW_OK;
}

}  // namespace rtw
