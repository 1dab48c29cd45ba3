// The GPU form of Raytracer::render (lib.rs:57-117).  Two schedules over the same arithmetic:
//
//  (1) scenes with a hierarchy — WAVEFRONT:
//        pool of P path slots (SoA in HBM) ──► k_wave_traverse ──► k_wave_shade ──► next iteration
//  (2) one-leaf scenes (<= 32 primitives, e.g. the Cornell box) — FUSED persistent kernel k_mega_flat: there is no
//      tree to walk, every ray tests the same primitive list, so nothing is gained by splitting intersection from
//      shading; a lane keeps its path in registers from the camera to its end and the scene sits in shared memory.
//      No path state ever touches HBM; a frame is ONE launch.
//
// Common to both:
// * A WORK ITEM = (pixel, sample slice).  Whoever owns the item (a pool slot / a lane) runs the samples of that slice
//   one after the other (lib.rs:83-88), adding each path's radiance to a running sum in sample order, writes the slice
//   sum when done and pulls the next item from a global cursor (path regeneration).  Slices of a pixel are added in
//   slice order by k_wave_resolve.  Every float addition therefore happens in an order fixed by (spp, slices) alone —
//   the image is bit-reproducible for any pool size, GPU count, schedule (1) or (2), or scheduling.
// * The default slice count depends on (width, height, samples, scene class) only — never on the pool or on the tile
//   partition — so a frame rendered by 8 GPUs has the same bits as the frame rendered by one.
// * sample_ray's recursion (lib.rs:97-117) is run in its iterative form L += T*e; T *= a.
//
// Wavefront specifics:
// * Both kernels are persistent: grid = SMs x resident blocks, warps pull 32 entries at a time from a device-side
//   cursor, so no launch parameter depends on the live count and the host only looks at the live count every BATCH
//   iterations, one round late.
// * While work items remain every slot is busy, so entry i of an iteration simply IS slot i ("identity" mode:
//   coalesced state accesses, no queue, no compaction atomics).  Once the item cursor has run dry the shade kernel
//   compacts the surviving slots into a queue each iteration ("queue" mode).
// * The per-frame values (camera, seed, frame pointer ...) live in a FrameDev record in device memory, so the
//   instantiated CUDA graph of BATCH iterations, the events and the pool are created once per scene and reused by
//   every later frame (r01 re-captured and re-instantiated the graph and created ~10 events per frame).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rtw_scene.cuh"
#include "rtw_traverse.cuh"
#include "rtw_raysort.cuh"

namespace rtw {

// ---- handle accounting (tests/test_gpu_parity.py::test_failed_render_leaks_nothing) ----------------------------------
static std::atomic<int> g_live_handles{0};
int live_handles() { return g_live_handles.load(); }
void count_handle(int delta) { g_live_handles += delta; }

namespace {

cudaError_t new_event(cudaEvent_t* e, unsigned flags) {
  cudaError_t r = cudaEventCreateWithFlags(e, flags);
  if (r == cudaSuccess) count_handle(1);
  return r;
}
void drop_event(cudaEvent_t& e) {
  if (e) { cudaEventDestroy(e); count_handle(-1); e = nullptr; }
}
cudaError_t new_stream(cudaStream_t* s) {
  cudaError_t r = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
  if (r == cudaSuccess) count_handle(1);
  return r;
}
void drop_stream(cudaStream_t& s) {
  if (s) { cudaStreamDestroy(s); count_handle(-1); s = nullptr; }
}
struct EventBag {  // events of one call (per-kernel timing): destroyed on every way out
  std::vector<cudaEvent_t> v;
  ~EventBag() { for (auto& e : v) drop_event(e); }
  cudaError_t add(cudaEvent_t* out) {
    cudaError_t r = new_event(out, cudaEventDefault);
    if (r == cudaSuccess) v.push_back(*out);
    return r;
  }
};

// ray_d.w of a slot: 0 = a live ray, else
#define RTW_SLOT_DEAD 1.0f     // no work left for the slot (end of the frame, identity mode)
#define RTW_SLOT_PENDING 2.0f  // traversal suspended in the drain of the last launch (rtw_traverse.cuh; off by default)

struct WaveCtl {
  unsigned long long item_cursor;
  unsigned long long segments;
  unsigned long long paths;
  unsigned long long pairs;
  unsigned long long prims;
  unsigned long long prim_bytes;
  unsigned long long regen_cursor0;  // item cursor before the current iteration's restart demand (k_wave_regen_scan)
  uint32_t count[2];       // entries of the iteration with that parity (identity mode: slot_count)
  uint32_t cursor_traverse;
  uint32_t cursor_shade;
  uint32_t qmode[2];       // 0 = identity (entry i is slot i), 1 = queue[parity] lists the live slots
  uint32_t exhausted;      // the item cursor ran past n_items
  uint32_t pad[3];
};

struct WaveDev {
  float4* ray_o;    // origin.xyz, time
  float4* ray_d;    // direction.xyz, slot flag
  int2* hit;        // primitive slot (or -1), t bits
  float4* thr;      // throughput T.rgb
  float4* sum;      // running slice sum rgb
  uint4* state;     // pixel index (row*w+col), sample, sample_end, bounce | slice << 8
  uint32_t* queue[2];
  float4* partial;  // [slices][pix_per_slice] slice sums of the pixels this partition owns (slices > 1)
  WaveCtl* ctl;
  uint32_t slot_count;
  // ray reordering (rtw_raysort.cuh); both null when off
  uint32_t* restart;         // split shade / restart: per-warp segments of slots whose path ended this iteration
  uint2* restart_counts;     // per segment: slots that go on with their item (listed from the front), slots that need a new item (from the back)
  uint32_t* restart_item_base;  // per segment: first work item for its need-an-item slots (k_wave_regen_scan)
  uint32_t* sort_key;     // [slot] key of the ray the slot holds, written with the ray
  const uint32_t* order;  // the entries of the iteration in key order: what the traversal kernel walks
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
  uint32_t pad0;
  uint32_t tile_shift;    // log2(tile_size) when it is a power of two, else 0xffffffff
  uint32_t skip_unowned;  // k_wave_resolve leaves the pixels of other partitions untouched (multi-GPU: one shared frame)
  uint32_t pad;
  float* accum;           // the frame: width*height*3 sums (this device's memory, or a peer's over NVLink)
  RaySortGrid sort;       // scene box for the ray keys (rtw_raysort.cuh)
};

// the traversal kernel variants a wavefront render can launch
enum TravKind { TK_PAIR = 0, TK_MEDIA, TK_WIDE, TK_COMPACT, TK_COUNT, TK_COUNT_COMPACT, TK_FLAT,
                TK_POOL_PAIR, TK_POOL_COMPACT, TK_POOL_COUNT_PAIR, TK_POOL_COUNT_COMPACT, TK_N };

struct GraphKey {
  int kind = -1, batch = 0, grid_t = 0, grid_s = 0, sort = 0, split = 0;
  uint32_t pool = 0;
  bool operator==(const GraphKey& o) const {
    return kind == o.kind && batch == o.batch && grid_t == o.grid_t && grid_s == o.grid_s && pool == o.pool && sort == o.sort &&
           split == o.split;
  }
};

struct WaveHost {
  WaveDev dev{};
  RaySortDev rsort{};              // ray reordering scratch (allocated with the pool on first use)
  uint32_t rsort_pool = 0;
  int rsort_blocks = 0;
  std::vector<void*> pool_allocs;
  uint32_t pool = 0;               // slots the pool arrays hold
  float4* partial = nullptr;
  size_t partial_elems = 0;
  WaveCtl* d_ctl = nullptr;
  FrameDev* d_frame = nullptr;     // read by the wavefront kernels (they are baked into a cached graph)
  FrameDev* h_frame = nullptr;     // pinned staging of d_frame
  WaveCtl* pinned_ctl = nullptr;   // [2]: ring for the lagging termination check, [2] = final counters
  cudaStream_t stream = nullptr;   // internal non-blocking stream (graph capture needs a non-legacy stream)
  cudaEvent_t ev_in = nullptr, ev_begin = nullptr, ev_end = nullptr, ring_ev[2] = {nullptr, nullptr};
  GraphKey graph_key;
  cudaGraphExec_t graph_exec = nullptr;
  int blocks_trav[TK_N] = {};      // persistent grid per TravKind
  int blocks_shade = 0;
  int blocks_shade_split = 0, blocks_regen = 0;
  uint32_t restart_pool = 0;       // pool size the restart list was allocated for
  int blocks_mega[2] = {0, 0};  // [LEAN]
};

// ---- work items ---------------------------------------------------------------------------------
struct Item {
  uint32_t pixel;  // row*w + col   (row = bottom-up like Pixel.row, lib.rs:58)
  uint32_t sample, sample_end, slice;
};

// (render_device guarantees n_items < 2^32 and nsamp * (slices + 1) < 2^32: everything stays in 32-bit arithmetic —
// a 64-bit division costs ~100 instructions and r01 carried a second, 64-bit copy of this function in every kernel)
__device__ __forceinline__ bool decode_item(const FrameDev& f, uint32_t n, Item& it) {
  uint32_t tile_local, within, wx, wy;
  const uint32_t pps = (uint32_t)f.pix_per_slice;
  const uint32_t k = n / pps;
  const uint32_t q = n - k * pps;
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
  const unsigned long long tile = (unsigned long long)tile_local * f.part_count + f.part_rank;
  if (tile >= f.tiles_total) return false;
  const uint32_t ty = (uint32_t)tile / f.tiles_x, tx = (uint32_t)tile - ty * f.tiles_x;
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
__device__ __forceinline__ bool fetch_item(const FrameDev& f, WaveCtl* ctl, bool need, Item& it) {
  const uint32_t lane = threadIdx.x & 31;
  bool got = false;
  for (;;) {
    uint32_t m = __ballot_sync(0xffffffffu, need && !got);
    if (m == 0) break;
    unsigned long long base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&ctl->item_cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (base + __popc(m) > f.n_items) {  // ran dry (the cursor may overshoot; harmless): later iterations compact
      if (lane == 0) ctl->exhausted = 1u;
      if (base >= f.n_items) break;
    }
    if (need && !got) {
      unsigned long long n = base + __popc(m & ((1u << lane) - 1u));
      if (n < f.n_items) got = decode_item(f, (uint32_t)n, it);
      else need = false;
    }
  }
  return got;
}

// lib.rs:84-86: the camera ray of sample `it.sample` of the item's pixel (stage 0 of the path's stream).
__device__ __forceinline__ void camera_path(const FrameDev& f, const Item& it, v3& o, v3& d, float& time) {
  uint32_t row = it.pixel / f.width, col = it.pixel - row * f.width;
  camera_ray_fresh(f.cam, f.width, f.height, row, col, ((uint64_t)f.seed_hi << 32) | f.seed_lo, it.pixel, it.sample, o, d, time);
}

__device__ __forceinline__ void start_path(const FrameDev& f, const WaveDev& w, uint32_t slot, const Item& it) {
  v3 o, d;
  float time;
  camera_path(f, it, o, d, time);
  w.ray_o[slot] = make_float4(o.x, o.y, o.z, time);
  w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
  w.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
  w.state[slot] = make_uint4(it.pixel, it.sample, it.sample_end, it.slice << 8);
  if (w.sort_key) w.sort_key[slot] = ray_sort_key(f.sort, o, d);
}

// the slice sum of a finished work item: straight into the frame (one slice) or into the slice-sum buffer
__device__ __forceinline__ void publish_item(const FrameDev& f, float4* __restrict__ partial, uint32_t pixel, uint32_t slice,
                                             v3 sum) {
  uint32_t row = pixel / f.width, col = pixel - row * f.width;
  if (f.slices == 1) {
    size_t pix = (size_t)(f.height - 1 - row) * f.width + col;
    f.accum[3 * pix] = sum.x; f.accum[3 * pix + 1] = sum.y; f.accum[3 * pix + 2] = sum.z;
  } else {
    partial[(size_t)slice * f.pix_per_slice + owned_index(f, col, f.height - 1 - row)] = make_float4(sum.x, sum.y, sum.z, 0.f);
  }
}

// One path segment after its closest-hit query (lib.rs:102-116): the ray (o, d, time) hit primitive type/inst with
// geometry words g0..g2 at distance t.  Returns true when the path ENDS with this segment (L = the radiance it
// contributes, throughput applied); otherwise T, bounce and the ray are advanced to the scattered ray.
// Shared by the wavefront shade kernel and the fused flat-scene kernel: one copy of the parity-critical arithmetic.
// LEAN: the scene's materials are all Lambertian / DiffuseLight over solid colours (SceneDev::all_diffuse_solid): the
// metal / glass / texture code is compiled out (fewer registers, a third of the code).
template <bool LEAN = false, class IV>
__device__ __forceinline__ bool shade_hit(const SceneDev& sc, const IV& iv, uint32_t max_depth, uint32_t meta,
                                          const MaterialRec& m, int32_t shade_idx, float4 g0, float4 g1, float4 g2, float t,
                                          uint64_t seed, uint32_t pixel, uint32_t sample, v3& o, v3& d, float time, v3& T,
                                          uint32_t& bounce, v3& L) {
  const bool need_uv = !LEAN && !m.solid && (m.type == MT_LAMBERTIAN || m.type == MT_DIFFUSE_LIGHT) && texture_needs_uv(sc, m.tex);
  HitRec rec;
  finalize_hit_iv(sc, iv, meta & 7u, meta >> RTW_META_TYPE_BITS, g0, g1, g2, shade_idx, o, d, time, t, need_uv, rec);
  Rng rng;
  rng.begin(seed, pixel, sample, bounce + 1);
  v3 att, out_dir;
  if (LEAN) {
    if (m.type != MT_LAMBERTIAN) {  // DiffuseLight: emits its solid colour, never scatters (light_source.rs:17-23)
      L = T * mk(m.r, m.g, m.b);
      return true;
    }
    out_dir = rec.normal + unit_vector(random_in_unit_sphere_fresh(rng));  // material.rs:42-56, as in material_scatter
    const float S = 1e-8f;
    if ((fabsf(out_dir.x) < S) && (fabsf(out_dir.y) < S) && (fabsf(out_dir.z) < S)) out_dir = rec.normal;
    att = mk(m.r, m.g, m.b);
  } else {
    v3 emitted;
    if (!material_shade(sc, m, d, rec, rng, emitted, att, out_dir)) {  // lib.rs:107-114
      L = T * emitted;
      return true;
    }
  }
  // L += T*emitted with emitted == 0 for every scattering material: exact no-op
  T = T * att;  // lib.rs:116
  bounce += 1;
  if (bounce >= max_depth) {  // lib.rs:98-100: depth exhausted -> black
    L = mk(0.f, 0.f, 0.f);
    return true;
  }
  o = rec.p;
  d = out_dir;
  return false;
}

__device__ __forceinline__ void load_frame(FrameDev& fs, const FrameDev* __restrict__ fp) {
  const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(fp);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&fs);
  for (uint32_t i = threadIdx.x; i < sizeof(FrameDev) / 4u; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
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

__global__ void __launch_bounds__(128) k_wave_init(SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w) {
  __shared__ FrameDev f;
  load_frame(f, fp);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;  // block = 128 threads: whole warps reach the ballots
  const uint32_t slot = tid;
  if (tid == 0) {  // iteration 0 reads the slots in identity mode
    w.ctl->count[0] = w.slot_count;
    w.ctl->qmode[0] = 0;
  }
  Item it;
  bool got = fetch_item(f, w.ctl, tid < w.slot_count, it);
  if (got) {
    start_path(f, w, slot, it);
    w.sum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (tid < w.slot_count) {
    w.ray_d[slot] = make_float4(0.f, 0.f, 0.f, RTW_SLOT_DEAD);
    if (w.sort_key) w.sort_key[slot] = RTW_RAYSORT_DEAD_KEY;
  }
  uint32_t m = __ballot_sync(0xffffffffu, got);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&w.ctl->paths, (unsigned long long)__popc(m));
}

// ---- traversal -----------------------------------------------------------------------------------
struct WaveIO {
  const WaveDev& w;
  const uint32_t* __restrict__ queue;  // nullptr: identity mode
  uint32_t slot;
  const FrameDev* __restrict__ fp;
  // key of the current path segment: (pixel, sample), stage = bounce + 1 — the stage the shade kernel uses
  __device__ __forceinline__ void rng_key(Rng& rng) {
    const uint4 st = w.state[slot];
    rng.begin(((uint64_t)fp->seed_hi << 32) | fp->seed_lo, st.x, st.y, (st.w & 0xffu) + 1u);
  }
#ifndef RTW_SUSPEND_LANES
#define RTW_SUSPEND_LANES 0  // A/B r01 (cow, monument): 4 / 8 / 12 cost 3-12 % — the restarts outweigh the shorter drain
#endif
  static constexpr int kSuspendLanes = RTW_SUSPEND_LANES;
  bool was_pending;
  __device__ __forceinline__ bool load(uint32_t i, v3& o, v3& d, float& time, float& t_min, float& t_max, int32_t& slot0,
                                       bool& resumed) {
    slot = queue ? queue[i] : i;
    const float4 o4 = w.ray_o[slot], d4 = w.ray_d[slot];
    if (d4.w == RTW_SLOT_DEAD) return false;  // identity mode at the end of the frame
    o = mk(o4.x, o4.y, o4.z); d = mk(d4.x, d4.y, d4.z); time = o4.w;
    t_min = 0.001f; t_max = __int_as_float(0x7f800000);  // lib.rs:102: world.hit(r, 0.001, f32::INFINITY)
    resumed = was_pending = d4.w == RTW_SLOT_PENDING;
    if (resumed) {  // suspended by the previous launch with this hit
      const int2 h = w.hit[slot];
      slot0 = h.x; t_max = __int_as_float(h.y);
    }
    return true;
  }
  __device__ __forceinline__ void store(uint32_t, v3, v3 d, float, int32_t hslot, float t, uint32_t) {
    w.hit[slot] = make_int2(hslot, __float_as_int(t));
    if (was_pending) w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
  }
  __device__ __forceinline__ void suspend(uint32_t, int32_t hslot, float t) {
    w.hit[slot] = make_int2(hslot, __float_as_int(t));
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
    const uint32_t out_queue = (in_queue || ctl->exhausted) ? 1u : 0u;
    ctl->qmode[parity ^ 1] = out_queue;
    ctl->count[parity ^ 1] = out_queue ? 0u : w.slot_count;
  }
}

// Flat scenes through the wavefront (instrumented runs, RTW_MEGA=0): rtw_traverse.cuh, traverse_flat
__global__ void __launch_bounds__(128, 8) k_wave_traverse_flat(SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w,
                                                               uint32_t parity) {
  __shared__ FlatRecords fr;
  uint32_t count, in_queue;
  traverse_prologue(w, parity, count, in_queue);
  stage_flat(sc, fr);
  WaveIO io{w, w.order ? w.order : (in_queue ? w.queue[parity] : nullptr), 0, fp, false};
  traverse_flat(sc, fr, io, count, &w.ctl->cursor_traverse);
}

template <bool COUNT, bool MEDIA, int NODES>
__global__ void __launch_bounds__(128, (COUNT || MEDIA || NODES == NODES_WIDE) ? 1 : RTW_TRAVERSE_MINBLOCKS)
    k_wave_traverse(SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w, uint32_t parity) {
  WaveCtl* ctl = w.ctl;
  uint32_t count, in_queue;
  traverse_prologue(w, parity, count, in_queue);
  const uint32_t lane = threadIdx.x & 31;
  TraverseCounters cnt;
  WaveIO io{w, w.order ? w.order : (in_queue ? w.queue[parity] : nullptr), 0, fp, false};
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

// The walk with pooled leaf tests (rtw_traverse.cuh: traverse_pooled): scenes without media
template <bool COUNT, int NODES>
__global__ void __launch_bounds__(128, COUNT ? 1 : RTW_TRAVERSE_MINBLOCKS)
    k_wave_traverse_pooled(SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w, uint32_t parity) {
  __shared__ LeafPool pools[4];
  WaveCtl* ctl = w.ctl;
  uint32_t count, in_queue;
  traverse_prologue(w, parity, count, in_queue);
  const uint32_t lane = threadIdx.x & 31;
  TraverseCounters cnt;
  WaveIO io{w, w.order ? w.order : (in_queue ? w.queue[parity] : nullptr), 0, fp, false};
  traverse_pooled<COUNT, NODES>(sc, io, count, &ctl->cursor_traverse, cnt, pools[threadIdx.x >> 5]);
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
// 32 lanes active.  They go on a per-warp backlog in shared memory instead; once 32 are waiting the warp restarts
// them together.
#define RTW_BACKLOG 64
#define RTW_NEED_ITEM 0x80000000u

struct ShadeBacklog {  // per warp; SoA so that lane i touches bank i
  uint32_t slot[RTW_BACKLOG], pixel[RTW_BACKLOG], sample[RTW_BACKLOG], sample_end[RTW_BACKLOG], slice[RTW_BACKLOG];
};

// restart the paths of backlog entries [first, first + 32) (those with valid == true): lib.rs:83-86
__device__ __forceinline__ void regenerate(const FrameDev& f, const WaveDev& w, ShadeBacklog& bl, uint32_t idx, bool valid,
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
  if (fetch_item(f, w.ctl, need_item, it)) {  // the slot finished its work item: pull the next one
    w.sum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    go = true;
  }
  if (go) {
    start_path(f, w, slot, it);
    new_paths++;
  } else if (valid) {
    w.ray_d[slot] = make_float4(0.f, 0.f, 0.f, RTW_SLOT_DEAD);  // no work left for this slot
    if (w.sort_key) w.sort_key[slot] = RTW_RAYSORT_DEAD_KEY;
  }
  if (out_queue) queue_push(next_queue, next_count, go, slot);
}

// ---- sm_100a: the state of a trip arrives through the bulk-copy engine ------------------------------------------------
// In identity mode (entry i is slot i — all of a frame but its tail) a trip's state is six CONTIGUOUS runs of the pool
// arrays (32 x 8 B hit records + 5 x 32 x 16 B).  One elected lane asks the copy engine for them (cp.async.bulk,
// global -> shared, completion counted in bytes on an mbarrier: SASS UBLKCP + SYNCS) as soon as the previous trip's state
// has been read into registers, so the fetch of trip k+1 overlaps the shading of trip k without holding a single
// register; the warp then waits on the mbarrier instead of on six scoreboards.  Trips are assigned statically
// (warp g takes trips g, g + warps, ...): the address of the next trip is known a whole trip ahead, and the per-trip
// cursor atomic (10 % of the kernel's stall samples in the r02f capture) disappears.  Queue mode (the tail of a frame:
// scattered slots, few of them) keeps the plain loads and the cursor.
#ifndef RTW_SHADE_BULK
#define RTW_SHADE_BULK 1
#endif
struct ShadeStage {  // the state of one trip (32 slots), filled by the copy engine
  float4 ray_o[32], ray_d[32], thr[32], sum[32];
  uint4 state[32];
  int2 hit[32];
};
constexpr uint32_t kShadeStageBytes = 5u * 512u + 256u;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RTW_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RTW_MBAR_DONE;\n"
      "bra RTW_MBAR_WAIT;\n"
      "RTW_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one lane: fetch the state of the 32 slots from `base` on
__device__ __forceinline__ void shade_stage_fetch(const WaveDev& w, uint32_t base, ShadeStage& stg, uint64_t* bar) {
  mbar_expect_tx(bar, kShadeStageBytes);
  bulk_g2s(stg.ray_o, w.ray_o + base, 512u, bar);
  bulk_g2s(stg.ray_d, w.ray_d + base, 512u, bar);
  bulk_g2s(stg.thr, w.thr + base, 512u, bar);
  bulk_g2s(stg.sum, w.sum + base, 512u, bar);
  bulk_g2s(stg.state, w.state + base, 512u, bar);
  bulk_g2s(stg.hit, w.hit + base, 256u, bar);
}

#ifndef RTW_SHADE_BULK_SINGLE
#define RTW_SHADE_BULK_SINGLE 0  // the bulk-copy staging inside the single shade kernel: r02 A/B cow 13.07 -> 13.76 ms, off
#endif
#ifdef RTW_SHADE_MINBLOCKS  // A/B r01: 6 (80 registers) and 7 (72) are 4 % and 11 % slower than the compiler's choice
__global__ void __launch_bounds__(128, RTW_SHADE_MINBLOCKS) k_wave_shade(
#else
__global__ void __launch_bounds__(128) k_wave_shade(
#endif
    SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w, uint32_t parity) {
  __shared__ ShadeBacklog backlog[4];
  __shared__ FrameDev f;
#if RTW_SHADE_BULK_SINGLE
  __shared__ __align__(128) ShadeStage stages[4];
  __shared__ __align__(8) uint64_t stage_bar[4];
  ShadeStage& stg = stages[threadIdx.x >> 5];
  uint64_t* bar = &stage_bar[threadIdx.x >> 5];
#endif
  load_frame(f, fp);
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
  const uint32_t max_depth = f.max_depth;
  const GlobalInst iv{sc.inst_range, sc.inst_ops};
  uint32_t new_paths = 0, nseg = 0;
  uint32_t nback = 0;  // warp-uniform: entries waiting on the backlog
  uint32_t grabbed = 0;
#if RTW_SHADE_BULK_SINGLE
  const bool ident = queue == nullptr && (count & 31u) == 0u;  // kernel-uniform: whole trips of consecutive slots
  uint32_t trip = blockIdx.x * 4u + (threadIdx.x >> 5);         // static assignment: trips trip, trip + warps, ...
  const uint32_t trip_stride = gridDim.x * 4u;
  const uint32_t ntrips = count >> 5;
  uint32_t phase = 0;
  if (ident) {
    if (lane == 0) {
      mbar_init(bar, 1u);
      if (trip < ntrips) shade_stage_fetch(w, trip << 5, stg, bar);
    }
    __syncwarp();
  } else
#else
  const bool ident = false;
  uint32_t trip = 0;
  const uint32_t ntrips = 0;
#endif
  if (lane == 0) grabbed = atomicAdd(&ctl->cursor_shade, 32u);
  for (;;) {
    const uint32_t base = ident ? (trip < ntrips ? trip << 5 : count) : __shfl_sync(0xffffffffu, grabbed, 0);
    const bool more = base < count;  // warp-uniform: another trip of 32 entries
    if (!more && nback == 0) break;
    const uint32_t i = base + lane;
    bool active = more && i < count;
    uint32_t slot = 0;
    bool alive = false;  // slot continues into the next iteration
    bool ended = false;  // the path ended: restart the slot (next sample of its item, or a new item)
    Item it;
    it.pixel = it.sample = it.sample_end = it.slice = 0;
    float4 o4, d4, T4, s4;
    int2 h;
    uint4 st;
#if RTW_SHADE_BULK_SINGLE
    if (ident && more) {  // the copy engine has (or will have) put this trip's state in shared memory
      mbar_wait(bar, phase);
      phase ^= 1u;
      slot = i;
      h = stg.hit[lane];
      o4 = stg.ray_o[lane]; d4 = stg.ray_d[lane];
      T4 = stg.thr[lane];
      st = stg.state[lane];
      s4 = stg.sum[lane];
      __syncwarp();  // every lane has read the stage: it may be overwritten
      trip += trip_stride;
      if (lane == 0 && trip < ntrips) shade_stage_fetch(w, trip << 5, stg, bar);  // overlaps the shading below
      alive = d4.w == RTW_SLOT_PENDING;
      active = d4.w == 0.f;
    } else
#endif
    if (active) {  // every load of the slot's state is issued before the first use
      slot = queue ? queue[i] : i;
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
        const uint32_t meta = sc.slot_meta[h.x];
        const int2 ms = sc.slot_ms[h.x];
        const MaterialRec m = sc.materials[ms.x];
        const uint32_t type = meta & 7u;
        const float4* __restrict__ g = sc.geom + 3 * (size_t)h.x;
        const float4 g0 = __ldg(g);
        float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g2 = g1;
        if (type == PT_MSPHERE || type == PT_TRI) { g1 = __ldg(g + 1); g2 = __ldg(g + 2); }
        v3 o = mk(o4.x, o4.y, o4.z), d = mk(d4.x, d4.y, d4.z);
        ended = shade_hit(sc, iv, max_depth, meta, m, ms.y, g0, g1, g2, __int_as_float(h.y), seed, st.x, st.y, o, d, o4.w, T,
                          bounce, L);
        if (!ended) {
          w.ray_o[slot] = make_float4(o.x, o.y, o.z, o4.w);
          w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
          w.thr[slot] = make_float4(T.x, T.y, T.z, 0.f);
          w.state[slot] = make_uint4(st.x, st.y, st.z, (st.w & ~0xffu) | bounce);
          if (w.sort_key) w.sort_key[slot] = ray_sort_key(f.sort, o, d);
          alive = true;
        }
      }
      if (ended) {
        v3 sum = mk(s4.x, s4.y, s4.z) + L;  // lib.rs:87: pixel_color += sample_ray(..)
        it.pixel = st.x; it.sample = st.y + 1; it.sample_end = st.z; it.slice = st.w >> 8;
        if (st.y + 1 < st.z) {  // next sample of the same item
          w.sum[slot] = make_float4(sum.x, sum.y, sum.z, 0.f);
        } else {  // item done: publish the slice sum
          publish_item(f, w.partial, st.x, st.w >> 8, sum);
          it.slice = RTW_NEED_ITEM;
        }
      }
    }
    if (out_queue && more) queue_push(next_queue, next_count, alive, slot);
    const uint32_t m_end = __ballot_sync(0xffffffffu, ended);
    if (ended) {
      const uint32_t pos = nback + __popc(m_end & lane_lt);
      bl.slot[pos] = slot; bl.pixel[pos] = it.pixel; bl.sample[pos] = it.sample; bl.sample_end[pos] = it.sample_end;
      bl.slice[pos] = it.slice;
    }
    nback += __popc(m_end);
    __syncwarp();
    if (!ident && more && lane == 0) grabbed = atomicAdd(&ctl->cursor_shade, 32u);
    // restart 32 waiting paths together — or, once the entries have run out, whatever is left (ONE call site: the
    // restart code is a fifth of the kernel)
    if (nback >= 32u || !more) {
      const uint32_t take = min(nback, 32u);
      nback -= take;
      regenerate(f, w, bl, nback + lane, lane < take, next_queue, next_count, out_queue, new_paths);
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    new_paths += __shfl_xor_sync(0xffffffffu, new_paths, off);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, off);
  }
  if (lane == 0 && new_paths) atomicAdd(&ctl->paths, (unsigned long long)new_paths);
  if (lane == 0 && nseg) atomicAdd(&ctl->segments, (unsigned long long)nseg);  // one world.hit per live entry (lib.rs:102)
}

// ---- shade and restart as TWO kernels (experiment, RTW_SHADE_SPLIT=1) ---------------------------------------------------------------
// k_wave_shade above is 65 KB of SASS; the r02 captures show its warps waiting for INSTRUCTIONS (no_instruction 1.5-5.4
// warps per issue: 20 warps per SM in different corners of a kernel twice the size of the instruction cache) as much as
// for data.  A fifth of that code restarts ended paths (work-item decode, Philox, camera ray) and runs for the 15-25 % of
// the lanes whose path just ended.  Here it is a kernel of its own:
//   k_wave_shade_split   shades the iteration's entries.  A path that ends adds its radiance to the slot's running sum
//                        (or publishes the finished item) and appends its slot to the warp's PRIVATE segment of the
//                        restart list — no atomics, the count of a segment is written once when the warp leaves.  Trips
//                        are assigned statically (warp g: trips g, g + warps, ...), so a segment cannot overflow and, in
//                        identity mode, the copy engine can fetch the next trip's state while this one is shaded.
//   k_wave_regen         warp g walks segment g 32 entries at a time, all lanes busy: next sample of the slot's item or a
//                        new item from the cursor (lib.rs:78-88), camera ray (camera.rs:66-74), fresh throughput.
// Same functions, same per-slot order of the float additions: the frame keeps its bits (RTW_SHADE_SPLIT=0: A/B and
// parity against the single kernel).
__device__ __forceinline__ uint32_t restart_segment_capacity(uint32_t slot_count, uint32_t warps) {
  const uint32_t ntrips = (slot_count + 31u) >> 5;
  return ((ntrips + warps - 1u) / warps) << 5;
}

#ifdef RTW_SHADE_MINBLOCKS
__global__ void __launch_bounds__(128, RTW_SHADE_MINBLOCKS) k_wave_shade_split(
#else
__global__ void __launch_bounds__(128) k_wave_shade_split(
#endif
    SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w, uint32_t parity) {
  __shared__ FrameDev f;
#if RTW_SHADE_BULK
  __shared__ __align__(128) ShadeStage stages[4];
  __shared__ __align__(8) uint64_t stage_bar[4];
  ShadeStage& stg = stages[threadIdx.x >> 5];
  uint64_t* bar = &stage_bar[threadIdx.x >> 5];
#endif
  load_frame(f, fp);
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
  const uint32_t max_depth = f.max_depth;
  const GlobalInst iv{sc.inst_range, sc.inst_ops};
  const uint32_t warps = gridDim.x * 4u, warp_g = blockIdx.x * 4u + (threadIdx.x >> 5);
  uint32_t* __restrict__ segment = w.restart + (size_t)warp_g * restart_segment_capacity(w.slot_count, warps);
  const uint32_t seg_cap = restart_segment_capacity(w.slot_count, warps);
  uint32_t n_same = 0, n_need = 0;  // warp-uniform: entries of this warp's segment, from the front / from the back
  uint32_t nseg = 0;
  const uint32_t ntrips = (count + 31u) >> 5;
  uint32_t trip = warp_g;
#if RTW_SHADE_BULK
  const bool ident = queue == nullptr && (count & 31u) == 0u;  // kernel-uniform: whole trips of consecutive slots
  uint32_t phase = 0;
  if (ident) {
    if (lane == 0) {
      mbar_init(bar, 1u);
      if (trip < ntrips) shade_stage_fetch(w, trip << 5, stg, bar);
    }
    __syncwarp();
  }
#else
  const bool ident = false;
#endif
  for (; trip < ntrips; trip += warps) {
    const uint32_t i = (trip << 5) + lane;
    bool active = i < count;
    uint32_t slot = 0;
    bool alive = false;  // slot continues into the next iteration
    bool ended = false;  // the path ended: the slot goes on the restart list
    float4 o4, d4, T4, s4;
    int2 h;
    uint4 st;
#if RTW_SHADE_BULK
    if (ident) {  // the copy engine has (or will have) put this trip's state in shared memory
      mbar_wait(bar, phase);
      phase ^= 1u;
      slot = i;
      h = stg.hit[lane];
      o4 = stg.ray_o[lane]; d4 = stg.ray_d[lane];
      T4 = stg.thr[lane];
      st = stg.state[lane];
      s4 = stg.sum[lane];
      __syncwarp();  // every lane has read the stage: it may be overwritten
      if (lane == 0 && trip + warps < ntrips) shade_stage_fetch(w, (trip + warps) << 5, stg, bar);  // overlaps the shading below
      alive = d4.w == RTW_SLOT_PENDING;
      active = d4.w == 0.f;
    } else
#endif
    if (active) {  // every load of the slot's state is issued before the first use
      slot = queue ? queue[i] : i;
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
        const uint32_t meta = sc.slot_meta[h.x];
        const int2 ms = sc.slot_ms[h.x];
        const MaterialRec m = sc.materials[ms.x];
        const uint32_t type = meta & 7u;
        const float4* __restrict__ g = sc.geom + 3 * (size_t)h.x;
        const float4 g0 = __ldg(g);
        float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g2 = g1;
        if (type == PT_MSPHERE || type == PT_TRI) { g1 = __ldg(g + 1); g2 = __ldg(g + 2); }
        v3 o = mk(o4.x, o4.y, o4.z), d = mk(d4.x, d4.y, d4.z);
        ended = shade_hit(sc, iv, max_depth, meta, m, ms.y, g0, g1, g2, __int_as_float(h.y), seed, st.x, st.y, o, d, o4.w, T,
                          bounce, L);
        if (!ended) {
          w.ray_o[slot] = make_float4(o.x, o.y, o.z, o4.w);
          w.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
          w.thr[slot] = make_float4(T.x, T.y, T.z, 0.f);
          w.state[slot] = make_uint4(st.x, st.y, st.z, (st.w & ~0xffu) | bounce);
          if (w.sort_key) w.sort_key[slot] = ray_sort_key(f.sort, o, d);
          alive = true;
        }
      }
      if (ended) {
        const v3 sum = mk(s4.x, s4.y, s4.z) + L;  // lib.rs:87: pixel_color += sample_ray(..)
        if (st.y + 1 < st.z) w.sum[slot] = make_float4(sum.x, sum.y, sum.z, 0.f);  // the item has samples left
        else publish_item(f, w.partial, st.x, st.w >> 8, sum);                      // item done: its slice sum
      }
    }
    if (out_queue) queue_push(next_queue, next_count, alive, slot);
    // the slot restarts: with the next sample of its item (front of the segment) or with a new item (back)
    const bool need = ended && !(st.y + 1 < st.z);
    const uint32_t m_same = __ballot_sync(0xffffffffu, ended && !need), m_need = __ballot_sync(0xffffffffu, need);
    if (ended) {
      if (need) segment[seg_cap - 1u - (n_need + __popc(m_need & lane_lt))] = slot;
      else segment[n_same + __popc(m_same & lane_lt)] = slot;
    }
    n_same += __popc(m_same);
    n_need += __popc(m_need);
  }
  if (lane == 0) w.restart_counts[warp_g] = make_uint2(n_same, n_need);
  for (int off = 16; off > 0; off >>= 1) nseg += __shfl_xor_sync(0xffffffffu, nseg, off);
  if (lane == 0 && nseg) atomicAdd(&ctl->segments, (unsigned long long)nseg);  // one world.hit per live entry (lib.rs:102)
}

// Work items for the slots that finished theirs: one block turns the per-segment demand into item ranges (exclusive
// scan from the item cursor) — the restart kernel then needs no atomic at all.  (r02: 52 k warp-level atomicAdds on the one
// 64-bit cursor per iteration were the floor of the restart work: ~5 ns each at the L2, 0.26 ms.)
__global__ void __launch_bounds__(1024) k_wave_regen_scan(const FrameDev* __restrict__ fp, WaveDev w, uint32_t shade_warps) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t per = (shade_warps + 1023u) / 1024u;
  const uint32_t begin = min(shade_warps, threadIdx.x * per), end = min(shade_warps, begin + per);
  uint32_t sum = 0;
  for (uint32_t g = begin; g < end; ++g) sum += w.restart_counts[g].y;
  uint32_t x = sum;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= (uint32_t)off) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = warp_sums[lane];
    uint32_t ws = v;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
      if (lane >= (uint32_t)off) ws += y;
    }
    warp_sums[lane] = ws - v;
    if (lane == 31) {  // the block's total: advance the cursor (no other kernel touches it while this one runs)
      const unsigned long long cur = w.ctl->item_cursor;
      if (cur + ws > fp->n_items) w.ctl->exhausted = 1u;  // ran dry (the cursor may overshoot; harmless): later iterations compact
      w.ctl->regen_cursor0 = cur;
      w.ctl->item_cursor = cur + ws;
    }
  }
  __syncthreads();
  uint32_t run = warp_sums[warp] + (x - sum);
  for (uint32_t g = begin; g < end; ++g) {
    w.restart_item_base[g] = run;  // relative to regen_cursor0
    run += w.restart_counts[g].y;
  }
}

// restart the slots the shade kernel listed (lib.rs:83-88).  Work unit = (segment g, batch b of 32 entries), b-major: the
// non-empty batches of all segments spread evenly over every warp of the grid.
__global__ void __launch_bounds__(128) k_wave_regen(SceneDev sc, const FrameDev* __restrict__ fp, WaveDev w, uint32_t parity,
                                                     uint32_t shade_warps) {
  __shared__ FrameDev f;
  load_frame(f, fp);
  WaveCtl* ctl = w.ctl;
  const bool out_queue = ctl->qmode[parity ^ 1] != 0;
  uint32_t* next_queue = w.queue[parity ^ 1];
  uint32_t* next_count = &ctl->count[parity ^ 1];
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * 4u;
  const uint32_t cap = restart_segment_capacity(w.slot_count, shade_warps);
  const unsigned long long cursor0 = ctl->regen_cursor0;  // the item cursor before this iteration's demand was added
  uint32_t new_paths = 0;
  const uint32_t units = shade_warps * (cap >> 5);
  for (uint32_t u = blockIdx.x * 4u + (threadIdx.x >> 5); u < units; u += warps) {
    const uint32_t g = u % shade_warps, base = (u / shade_warps) << 5;
    const uint2 n = w.restart_counts[g];
    if (base >= n.x && base + 32u <= cap - n.y) continue;  // this batch holds no entry
    const uint32_t e = base + lane;
    const bool same = e < n.x, need = e >= cap - n.y;
    uint32_t slot = 0;
    Item it;
    it.pixel = it.sample = it.sample_end = it.slice = 0;
    bool go = false;
    if (same || need) slot = (w.restart + (size_t)g * cap)[e];
    if (same) {
      const uint4 st = w.state[slot];
      it.pixel = st.x; it.sample = st.y + 1; it.sample_end = st.z; it.slice = st.w >> 8;
      go = true;
    }
    bool got = false, retry = false;
    if (need) {
      const unsigned long long item = cursor0 + w.restart_item_base[g] + (cap - 1u - e);
      if (item < f.n_items) {
        got = decode_item(f, (uint32_t)item, it);
        retry = !got;  // an item of the tile padding (outside the image): nothing to render, take another one
      }
    }
    if (__any_sync(0xffffffffu, retry)) got = fetch_item(f, ctl, retry, it) || got;  // rare: through the cursor
    if (got) {
      w.sum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
      go = true;
    }
    if (go) {
      start_path(f, w, slot, it);
      new_paths++;
    } else if (need) {
      w.ray_d[slot] = make_float4(0.f, 0.f, 0.f, RTW_SLOT_DEAD);  // no work left for this slot
      if (w.sort_key) w.sort_key[slot] = RTW_RAYSORT_DEAD_KEY;
    }
    if (out_queue) queue_push(next_queue, next_count, go, slot);
  }
  for (int off = 16; off > 0; off >>= 1) new_paths += __shfl_xor_sync(0xffffffffu, new_paths, off);
  if (lane == 0 && new_paths) atomicAdd(&ctl->paths, (unsigned long long)new_paths);
}

// ---- the fused kernel of one-leaf scenes ------------------------------------------------------------------------------
// Persistent: grid = SMs x resident blocks.  A lane owns one path at a time and keeps it in registers: camera ray
// (lib.rs:84-86) -> closest hit over the staged primitive list (hittable/mod.rs:57-69, rtw_traverse.cuh::flat_closest)
// -> shade / scatter (shade_hit) -> ... until the path ends; then the next sample of its work item, or the next item from
// the global cursor.  The primitive records, the instance chains and the material of every primitive slot are staged
// in shared memory once per block; the only global traffic of a frame is the slice sums (16 B per work item).
// A lane whose path ended restarts at the top of the next trip (all lanes of a warp run the list walk together again);
// a warp leaves when the cursor is dry and none of its lanes holds a path.
#ifndef RTW_MEGA_MINBLOCKS
#define RTW_MEGA_MINBLOCKS 7  // r02 A/B (Cornell, final code): 5 -> 11256, 6 -> 12407, 7 -> 12942 Mrays/s (72 registers, 28 warps per SM)
#endif
#define RTW_MEGA_MAX_OPS 264  // 33 chains x RTW_MAX_CHAIN

struct MegaShared {
  FlatRecords fr;
  MaterialRec mat[32];   // material of primitive SLOT k
  int32_t shade[32];     // its TriShade index or -1
  uint2 inst_range[40];
  InstOp inst_ops[RTW_MEGA_MAX_OPS];
};

#ifndef RTW_MEGA_REGEN_MIN
#define RTW_MEGA_REGEN_MIN 4  // r02 A/B: 1 -> 10318, 4 -> 10737, 8 -> 10620 Mrays/s. lanes that must be waiting before a warp runs the restart code (or no lane holds a path)
#endif
template <bool LEAN>
__global__ void __launch_bounds__(128, RTW_MEGA_MINBLOCKS) k_mega_flat(SceneDev sc, FrameDev f, float4* __restrict__ partial,
                                                                      WaveCtl* __restrict__ ctl) {
  __shared__ MegaShared sh;
  for (uint32_t i = threadIdx.x; i < sc.flat_count; i += blockDim.x) {
    const int2 ms = sc.slot_ms[i];
    sh.mat[i] = sc.materials[ms.x];
    sh.shade[i] = ms.y;
  }
  for (uint32_t i = threadIdx.x; i < sc.num_insts; i += blockDim.x) sh.inst_range[i] = sc.inst_range[i];
  for (uint32_t i = threadIdx.x; i < sc.num_inst_ops; i += blockDim.x) sh.inst_ops[i] = sc.inst_ops[i];
  stage_flat(sc, sh.fr);  // ends with __syncthreads()
  const SharedInst iv{sh.inst_range, sh.inst_ops};
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t seed = ((uint64_t)f.seed_hi << 32) | f.seed_lo;
  uint32_t nseg = 0, npaths = 0;
  // the lane's path
  v3 o = mk(0.f, 0.f, 0.f), d = mk(0.f, 0.f, 1.f), T = mk(1.f, 1.f, 1.f), sum = mk(0.f, 0.f, 0.f);
  float time = 0.f;
  uint32_t bounce = 0;
  Item it;
  it.pixel = it.sample = it.sample_end = it.slice = 0;
  bool have = false;       // the lane holds a live path
  bool need_item = true;   // its work item is finished (or it never had one)
  bool done = false;       // the cursor is dry for this lane
  for (;;) {
    // ---- (re)start paths: next sample of the lane's item, or a new item ---------------------------------------------
    const bool want = !have && !done;
    const uint32_t m_want = __ballot_sync(0xffffffffu, want);
    if (m_want != 0u && ((uint32_t)__popc(m_want) >= RTW_MEGA_REGEN_MIN || !__any_sync(0xffffffffu, have))) {
      bool go = want && !need_item;
      Item nit = it;
      if (fetch_item(f, ctl, want && need_item, nit)) {
        it = nit;
        sum = mk(0.f, 0.f, 0.f);
        need_item = false;
        go = true;
      } else if (want && need_item) {
        done = true;
      }
      if (go) {
        camera_path(f, it, o, d, time);
        T = mk(1.f, 1.f, 1.f);
        bounce = 0;
        have = true;
        npaths++;
      }
    }
    if (!__any_sync(0xffffffffu, have)) break;  // every lane is done
    // ---- closest hit: lib.rs:102 world.hit(r, 0.001, INFINITY) ---------------------------------------------------------
    float best_t = __int_as_float(0x7f800000);
    int32_t best_slot;
    uint32_t best_meta;
    flat_closest(iv, sh.fr, o, d, time, 0.001f, have, best_t, best_slot, best_meta);
    // ---- shade ---------------------------------------------------------------------------------------------------------
    if (have) {
      nseg++;
      v3 L;
      bool ended;
      if (best_slot < 0) {  // lib.rs:102-105: miss -> background
        L = T * f.background;
        ended = true;
      } else {
        ended = shade_hit<LEAN>(sc, iv, f.max_depth, best_meta, sh.mat[best_slot], sh.shade[best_slot], sh.fr.g[best_slot][0],
                          sh.fr.g[best_slot][1], sh.fr.g[best_slot][2], best_t, seed, it.pixel, it.sample, o, d, time, T,
                          bounce, L);
      }
      if (ended) {
        sum = sum + L;  // lib.rs:87: pixel_color += sample_ray(..)
        it.sample += 1;
        have = false;
        if (it.sample >= it.sample_end) {  // item done: publish the slice sum
          publish_item(f, partial, it.pixel, it.slice, sum);
          need_item = true;
        }
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    npaths += __shfl_xor_sync(0xffffffffu, npaths, off);
    nseg += __shfl_xor_sync(0xffffffffu, nseg, off);
  }
  if (lane == 0 && npaths) atomicAdd(&ctl->paths, (unsigned long long)npaths);
  if (lane == 0 && nseg) atomicAdd(&ctl->segments, (unsigned long long)nseg);
}

// pixel_color = sum over slices, in slice order; pixels of other partitions = 0 (or untouched: skip_unowned)
__global__ void k_wave_resolve(FrameDev f, const float4* __restrict__ partial) {
  float* __restrict__ accum = f.accum;
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
    } else if (f.skip_unowned) {
      continue;
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

// debug (RTW_RAYSORT_CHECK=1, instrumented path): inversions of order[] and entries that are not a permutation
__global__ void k_raysort_check(const uint32_t* __restrict__ count, const uint32_t* __restrict__ key, const uint32_t* __restrict__ order,
                                unsigned long long* __restrict__ out) {
  const uint32_t n = *count;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x) {
    const uint32_t mask = (1u << RTW_RAYSORT_KEY_BITS) - 1u;  // the passes sort these bits (a dead slot's key is all ones)
    if ((key[order[i]] & mask) > (key[order[i + 1]] & mask)) atomicAdd(&out[0], 1ull);
    if ((key[order[i]] & mask) != (key[order[i + 1]] & mask)) atomicAdd(&out[1], 1ull);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&out[2], (unsigned long long)n);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <class T>
int wave_alloc(std::vector<void*>* bag, T** out, size_t count) {
  void* p = nullptr;
  cudaError_t e = dev_malloc(&p, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(RTW_ERR_NOMEM, std::string("cudaMalloc (wavefront state) failed: ") + cudaGetErrorString(e));
  }
  if (bag) bag->push_back(p);
  *out = (T*)p;
  return RTW_OK;
}

void drop_graph(WaveHost* wh) {
  if (wh->graph_exec) {
    cudaGraphExecDestroy(wh->graph_exec);
    count_handle(-1);
    wh->graph_exec = nullptr;
  }
  wh->graph_key = GraphKey();
}

void release_pool(WaveHost* wh) {
  drop_graph(wh);  // the graph bakes the pool pointers
  for (void* p : wh->pool_allocs) mem_free(p);
  wh->pool_allocs.clear();
  wh->pool = 0;
  wh->rsort = RaySortDev{};
  wh->rsort_pool = 0;
  wh->restart_pool = 0;
  float4* keep = wh->dev.partial;
  WaveCtl* ctl = wh->dev.ctl;
  wh->dev = WaveDev{};
  wh->dev.partial = keep;
  wh->dev.ctl = ctl;
}

void destroy_wave(WaveHost* wh) {
  if (!wh) return;
  release_pool(wh);
  if (wh->partial) mem_free(wh->partial);
  if (wh->d_ctl) mem_free(wh->d_ctl);
  if (wh->d_frame) mem_free(wh->d_frame);
  if (wh->h_frame) mem_free(wh->h_frame);
  if (wh->pinned_ctl) mem_free(wh->pinned_ctl);
  drop_event(wh->ev_in); drop_event(wh->ev_begin); drop_event(wh->ev_end);
  drop_event(wh->ring_ev[0]); drop_event(wh->ring_ev[1]);
  drop_stream(wh->stream);
  delete wh;
}

// streams, events, control blocks, occupancy-derived grids: once per scene.  The scene gets the WaveHost only when all
// of it exists (a failure half way frees what was made).
int create_wave(rtw_scene* s, WaveHost** out) {
  WaveHost* wh = new WaveHost();
  auto fail = [&](int rc) { destroy_wave(wh); return rc; };
#define RTW_WAVE_TRY(expr)                                         \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return fail(cuda_fail(_e, #expr));       \
  } while (0)
  RTW_WAVE_TRY(pinned_malloc((void**)&wh->pinned_ctl, 3 * sizeof(WaveCtl)));
  RTW_WAVE_TRY(pinned_malloc((void**)&wh->h_frame, sizeof(FrameDev)));
  RTW_WAVE_TRY(dev_malloc((void**)&wh->d_frame, sizeof(FrameDev)));
  RTW_WAVE_TRY(dev_malloc((void**)&wh->d_ctl, sizeof(WaveCtl)));
  wh->dev.ctl = wh->d_ctl;
  RTW_WAVE_TRY(new_stream(&wh->stream));
  RTW_WAVE_TRY(new_event(&wh->ev_in, cudaEventDisableTiming));
  RTW_WAVE_TRY(new_event(&wh->ev_begin, cudaEventDefault));
  RTW_WAVE_TRY(new_event(&wh->ev_end, cudaEventDefault));
  RTW_WAVE_TRY(new_event(&wh->ring_ev[0], cudaEventDisableTiming));
  RTW_WAVE_TRY(new_event(&wh->ring_ev[1], cudaEventDisableTiming));
  int nb = 0;
  const void* kernels[TK_N] = {(const void*)k_wave_traverse<false, false, NODES_PAIR>, (const void*)k_wave_traverse<false, true, NODES_PAIR>,
                               (const void*)k_wave_traverse<false, false, NODES_WIDE>, (const void*)k_wave_traverse<false, false, NODES_COMPACT>,
                               (const void*)k_wave_traverse<true, true, NODES_PAIR>, (const void*)k_wave_traverse<true, true, NODES_COMPACT>,
                               (const void*)k_wave_traverse_flat,
                               (const void*)k_wave_traverse_pooled<false, NODES_PAIR>, (const void*)k_wave_traverse_pooled<false, NODES_COMPACT>,
                               (const void*)k_wave_traverse_pooled<true, NODES_PAIR>, (const void*)k_wave_traverse_pooled<true, NODES_COMPACT>};
  for (int k = 0; k < TK_N; ++k) {
    RTW_WAVE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernels[k], 128, 0));
    wh->blocks_trav[k] = std::max(nb, 1) * s->num_sms;
  }
  RTW_WAVE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wave_shade, 128, 0));
  wh->blocks_shade = std::max(nb, 1) * s->num_sms;
  RTW_WAVE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wave_shade_split, 128, 0));
  wh->blocks_shade_split = std::max(nb, 1) * s->num_sms;
  RTW_WAVE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_wave_regen, 128, 0));
  wh->blocks_regen = std::max(nb, 1) * s->num_sms;
  RTW_WAVE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mega_flat<false>, 128, 0));
  wh->blocks_mega[0] = std::max(nb, 1) * s->num_sms;
  RTW_WAVE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mega_flat<true>, 128, 0));
  wh->blocks_mega[1] = std::max(nb, 1) * s->num_sms;
#undef RTW_WAVE_TRY
  *out = wh;
  return RTW_OK;
}

}  // namespace

void free_wave(rtw_scene* s) {
  destroy_wave((WaveHost*)s->wave);
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

int render_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* d_accum, cudaStream_t user_stream,
                  rtw_render_stats* stats, bool skip_unowned) {
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
  f.skip_unowned = (skip_unowned && f.part_count > 1) ? 1u : 0u;
  f.accum = d_accum;
  const uint32_t owned_tiles = f.tiles_total > f.part_rank ? (f.tiles_total - f.part_rank + f.part_count - 1) / f.part_count : 0;
  f.pix_per_slice = (unsigned long long)owned_tiles * f.tile_size * f.tile_size;

  const bool count_trav = (p->flags & RTW_RENDER_COUNT_TRAVERSAL) != 0;
  const bool time_kernels = (p->flags & RTW_RENDER_TIME_KERNELS) != 0;
  // a scene that is one leaf (<= 32 primitives, no media) is not walked at all
  const bool flat_class = s->dev.flat_count > 0 && !s->dev.has_media;
  bool flat = flat_class;
  if (const char* e = getenv("RTW_FLAT")) flat = flat && atoi(e) != 0;  // 0: one-leaf scenes through the general walk
  // ... and runs fused (k_mega_flat) unless the call wants per-kernel times / traversal counters, which only the
  // wavefront kernels produce (RTW_MEGA=0: A/B against the wavefront)
  bool mega = flat && !count_trav && !time_kernels && s->dev.num_insts <= 40 && s->dev.num_inst_ops <= RTW_MEGA_MAX_OPS;
  if (const char* e = getenv("RTW_MEGA")) mega = mega && atoi(e) != 0;

  // ---- slices: from (width, height, samples, scene class) only ---------------------------------------------------------
  // Enough work items that the last ones to finish are a small fraction of the frame, for one GPU and for an 8-way
  // tile partition alike.  One-leaf scenes: ~75 k lanes consume items, 16 samples per item are plenty (slice sums
  // <= 1 GiB).  Hierarchies: the wavefront keeps 2^23 slots busy, a slot runs the samples of its item one after the
  // other, so coarse items leave the pool idle behind a few long chains (r01: 8-way partition of Cornell, 105 ms at
  // 64 slices vs 75 ms at 830, ideal 70): ~32 items per slot on one GPU = ~4 under an 8-way partition (slice sums <= 4 GiB).
  // NEVER derived from the pool size or the partition: the slice count fixes the order of the float additions, and
  // the frame must have the same bits on 1 or 8 GPUs.
  uint32_t slices = p->slices;
  if (slices == 0) {
    const unsigned long long full = std::max<unsigned long long>(npix, 1);
    unsigned long long want, cap;
    if (flat_class) {
      want = (nsamp + 15ull) / 16ull;
      cap = std::max<unsigned long long>((64ull << 20) / full, 1ull);
    } else {
      want = (32ull * (1ull << 23) + full - 1) / full;
      cap = std::max<unsigned long long>((256ull << 20) / full, 1ull);
    }
    slices = (uint32_t)std::min(want, cap);
  }
  slices = std::max(1u, std::min(slices, std::max(nsamp, 1u)));
  if (slices > 0xFFFFFFu) return set_error(RTW_ERR_INVALID, "render: too many slices");
  f.slices = slices;
  f.n_items = (nsamp == 0) ? 0ull : f.pix_per_slice * slices;
  if (f.n_items >= (1ull << 32) || (unsigned long long)nsamp * (slices + 1ull) >= (1ull << 32))
    return set_error(RTW_ERR_INVALID, "render: more than 2^32 work items (pixels x slices): use fewer slices");
  f.tile_shift = 0xffffffffu;
  if ((f.tile_size & (f.tile_size - 1u)) == 0u) {
    f.tile_shift = 0;
    while ((1u << f.tile_shift) < f.tile_size) f.tile_shift++;
  }

  // ---- scratch (kept on the scene between calls) -------------------------------------------------
  WaveHost* wh = (WaveHost*)s->wave;
  if (!wh) {
    int rc = create_wave(s, &wh);
    if (rc != RTW_OK) return rc;
    s->wave = wh;
  }
  cudaStream_t st = wh->stream;

  // Wavefront pool (measured, profiles/r01_sweeps.txt): every launch of the persistent traversal ends with a drain in
  // which the SMs wait for the longest rays (~13 us flat scene, ~67 us cow), so launches should be few and large; the
  // slot state is streamed coalesced, it does not need to stay in L2.  2^22 slots for a flat scene, 2^23 with a
  // hierarchy (88 B per slot + two queues: 0.4 / 0.8 GB).  The fused kernel has no pool.
  uint32_t pool = 0;
  if (!mega) {
    pool = p->pool_size ? p->pool_size : (flat_class ? (1u << 22) : (1u << 23));
    pool = (pool + 31u) & ~31u;
    pool = (uint32_t)std::min<unsigned long long>(pool, std::max<unsigned long long>((f.n_items + 31ull) & ~31ull, 32ull));
    if (wh->pool < pool) {  // grow only: a smaller frame reuses the larger arrays
      release_pool(wh);
      int rc;
      WaveDev& w = wh->dev;
      if ((rc = wave_alloc(&wh->pool_allocs, &w.ray_o, pool)) || (rc = wave_alloc(&wh->pool_allocs, &w.ray_d, pool)) ||
          (rc = wave_alloc(&wh->pool_allocs, &w.hit, pool)) || (rc = wave_alloc(&wh->pool_allocs, &w.thr, pool)) ||
          (rc = wave_alloc(&wh->pool_allocs, &w.sum, pool)) || (rc = wave_alloc(&wh->pool_allocs, &w.state, pool)) ||
          (rc = wave_alloc(&wh->pool_allocs, &w.queue[0], pool)) || (rc = wave_alloc(&wh->pool_allocs, &w.queue[1], pool))) {
        release_pool(wh);
        return rc;
      }
      wh->pool = pool;
    }
  }
  // Ray reordering between iterations (rtw_raysort.cuh): on for hierarchies that live in HBM (compact pairs exist),
  // RTW_RAYSORT=0/1 overrides.
  bool raysort = !mega && !flat && s->dev.nodes_c != nullptr && !s->dev.has_media;
  if (const char* e = getenv("RTW_RAYSORT")) raysort = !mega && !flat && atoi(e) != 0;
  if (raysort && wh->rsort_pool < wh->pool) {
    drop_graph(wh);
    RaySortDev& r = wh->rsort;
    wh->rsort_blocks = std::min(RTW_RAYSORT_MAX_BLOCKS, 8 * s->num_sms);
    int rc;
    if ((rc = wave_alloc(&wh->pool_allocs, &r.key, wh->pool)) || (rc = wave_alloc(&wh->pool_allocs, &r.keys[0], wh->pool)) ||
        (rc = wave_alloc(&wh->pool_allocs, &r.keys[1], wh->pool)) || (rc = wave_alloc(&wh->pool_allocs, &r.vals[0], wh->pool)) ||
        (rc = wave_alloc(&wh->pool_allocs, &r.vals[1], wh->pool)) ||
        (rc = wave_alloc(&wh->pool_allocs, &r.hist, (size_t)256 * RTW_RAYSORT_MAX_BLOCKS)) ||
        (rc = wave_alloc(&wh->pool_allocs, &r.totals, (size_t)256))) {
      release_pool(wh);
      return rc;
    }
    wh->rsort_pool = wh->pool;
  }
  // shade and restart as two kernels (k_wave_shade_split + k_wave_regen); RTW_SHADE_SPLIT=0: the single kernel
  // (r02 A/B, profiles/r02_sweeps.txt: the split kernels lose 8-20 % against the single kernel — the restart code is 40 % of the
  // shading instructions and gains nothing from running apart — so the single kernel stays the product path)
  bool split = false;
  if (const char* e = getenv("RTW_SHADE_SPLIT")) split = !mega && atoi(e) != 0;
  if (split && wh->restart_pool < wh->pool) {
    drop_graph(wh);
    int rc;
    if ((rc = wave_alloc(&wh->pool_allocs, &wh->dev.restart, (size_t)wh->pool + 32u * 4u * (size_t)wh->blocks_shade_split + 32u)) ||
        (rc = wave_alloc(&wh->pool_allocs, &wh->dev.restart_counts, 4u * (size_t)wh->blocks_shade_split)) ||
        (rc = wave_alloc(&wh->pool_allocs, &wh->dev.restart_item_base, 4u * (size_t)wh->blocks_shade_split))) {
      release_pool(wh);
      return rc;
    }
    wh->restart_pool = wh->pool;
  }
  wh->dev.sort_key = raysort ? wh->rsort.key : nullptr;
  wh->dev.order = raysort ? wh->rsort.vals[(RTW_RAYSORT_PASSES - 1) & 1] : nullptr;
  f.sort.enabled = raysort ? 1u : 0u;
  for (int a = 0; a < 3; ++a) {
    const float lo = s->root_box[a], hi = s->root_box[3 + a];
    f.sort.lo[a] = lo;
    f.sort.hi[a] = hi;
    f.sort.scale[a] = (hi > lo) ? 1024.0f / (hi - lo) : 0.0f;
  }
  const size_t partial_elems = slices > 1 ? (size_t)slices * f.pix_per_slice : 0;
  if (wh->partial_elems < partial_elems) {
    drop_graph(wh);
    if (wh->partial) mem_free(wh->partial);
    wh->partial = nullptr;
    wh->partial_elems = 0;
    int rc = wave_alloc((std::vector<void*>*)nullptr, &wh->partial, partial_elems);
    if (rc) return rc;
    wh->partial_elems = partial_elems;
  }
  wh->dev.partial = wh->partial;
  wh->dev.ctl = wh->d_ctl;
  wh->dev.slot_count = pool;
  const WaveDev w = wh->dev;

  // The 4-wide walk (SceneDev::nodes4) halves the dependent node fetches but moves MORE bytes (it fetches the boxes
  // below a child whose own box the ray misses).  Measured r01: no gain on C5 — that traversal sits at the
  // random-gather bandwidth of the memory system, not at its latency — and 5-17 % slower on cache-resident scenes.
  // Built and used only with RTW_WIDE=1 (parity-tested).
  bool wide = false;
  if (const char* e = getenv("RTW_WIDE"))
    wide = atoi(e) != 0 && s->dev.nodes4 != nullptr && !s->dev.has_media && 3u * (s->bvh_height / 2u + 1u) + 2u <= RTW_STACK_SIZE;
  // compact pairs exist only when rtw_build found the hierarchy too large for the caches (rtw_bvh.cu)
  const bool compact = s->dev.nodes_c != nullptr && !s->dev.has_media;
  // leaf tests pooled per warp (traverse_pooled): every hierarchy without media; RTW_POOLED=0 selects the per-lane walk
  bool pooled = !s->dev.has_media && !flat && !wide && s->dev.num_prims < (1u << 27);
  if (const char* e = getenv("RTW_POOLED")) pooled = pooled && atoi(e) != 0;
  else pooled = false;  // measured slower (profiles/r02_sweeps.txt): experiment only
  const TravKind kind = pooled ? (count_trav ? (compact ? TK_POOL_COUNT_COMPACT : TK_POOL_COUNT_PAIR) : (compact ? TK_POOL_COMPACT : TK_POOL_PAIR))
                               : (count_trav ? (compact ? TK_COUNT_COMPACT : TK_COUNT)
                                             : (s->dev.has_media ? TK_MEDIA : (flat ? TK_FLAT : (compact ? TK_COMPACT : (wide ? TK_WIDE : TK_PAIR)))));
  const FrameDev* dfp = wh->d_frame;
  auto launch_traverse = [&](uint32_t parity, int grid) {
    switch (kind) {
      case TK_PAIR: k_wave_traverse<false, false, NODES_PAIR><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_MEDIA: k_wave_traverse<false, true, NODES_PAIR><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_WIDE: k_wave_traverse<false, false, NODES_WIDE><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_COMPACT: k_wave_traverse<false, false, NODES_COMPACT><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_COUNT: k_wave_traverse<true, true, NODES_PAIR><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_FLAT: k_wave_traverse_flat<<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_POOL_PAIR: k_wave_traverse_pooled<false, NODES_PAIR><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_POOL_COMPACT: k_wave_traverse_pooled<false, NODES_COMPACT><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_POOL_COUNT_PAIR: k_wave_traverse_pooled<true, NODES_PAIR><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      case TK_POOL_COUNT_COMPACT: k_wave_traverse_pooled<true, NODES_COMPACT><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
      default: k_wave_traverse<true, true, NODES_COMPACT><<<grid, 128, 0, st>>>(s->dev, dfp, w, parity); break;
    }
  };

  // order[] for the iteration of this parity: LSD radix passes over (key[slot], slot) of its entry list
  auto launch_sort = [&](uint32_t parity) {
    const RaySortDev& r = wh->rsort;
    const uint32_t* cnt = &wh->d_ctl->count[parity];
    const uint32_t* qm = &wh->d_ctl->qmode[parity];
    const int g = wh->rsort_blocks;
    for (int ps = 0; ps < RTW_RAYSORT_PASSES; ++ps) {
      const int shift = 8 * ps;
      const bool last = ps == RTW_RAYSORT_PASSES - 1;
      const uint32_t* kin = ps == 0 ? r.key : r.keys[(ps - 1) & 1];
      const uint32_t* vin = ps == 0 ? nullptr : r.vals[(ps - 1) & 1];
      if (ps == 0) k_raysort_hist<true><<<g, RTW_RAYSORT_THREADS, 0, st>>>(cnt, qm, w.queue[parity], kin, shift, r.hist);
      else k_raysort_hist<false><<<g, RTW_RAYSORT_THREADS, 0, st>>>(cnt, qm, nullptr, kin, shift, r.hist);
      k_raysort_scan<<<256, 1024, 0, st>>>((uint32_t)g, r.hist, r.totals);
      if (ps == 0 && last) k_raysort_scatter<true, true><<<g, RTW_RAYSORT_THREADS, 0, st>>>(cnt, qm, w.queue[parity], kin, vin, r.keys[ps & 1], r.vals[ps & 1], shift, r.hist, r.totals);
      else if (ps == 0) k_raysort_scatter<true, false><<<g, RTW_RAYSORT_THREADS, 0, st>>>(cnt, qm, w.queue[parity], kin, vin, r.keys[ps & 1], r.vals[ps & 1], shift, r.hist, r.totals);
      else if (last) k_raysort_scatter<false, true><<<g, RTW_RAYSORT_THREADS, 0, st>>>(cnt, qm, nullptr, kin, vin, r.keys[ps & 1], r.vals[ps & 1], shift, r.hist, r.totals);
      else k_raysort_scatter<false, false><<<g, RTW_RAYSORT_THREADS, 0, st>>>(cnt, qm, nullptr, kin, vin, r.keys[ps & 1], r.vals[ps & 1], shift, r.hist, r.totals);
    }
  };
  const uint32_t sort_launches = (raysort ? 3u * RTW_RAYSORT_PASSES : 0u) + (split ? 2u : 0u);  // per iteration, on top of traverse + shade
  auto launch_shade = [&](uint32_t parity, int grid) {
    if (split) {
      const int gs = std::min(grid, wh->blocks_shade_split);
      k_wave_shade_split<<<gs, 128, 0, st>>>(s->dev, dfp, w, parity);
      k_wave_regen_scan<<<1, 1024, 0, st>>>(dfp, w, (uint32_t)gs * 4u);
      k_wave_regen<<<wh->blocks_regen, 128, 0, st>>>(s->dev, dfp, w, parity, (uint32_t)gs * 4u);
    } else {
      k_wave_shade<<<std::min(grid, wh->blocks_shade), 128, 0, st>>>(s->dev, dfp, w, parity);
    }
  };

  // All work runs on an internal stream ordered after the caller's stream; the call returns only after that stream has
  // drained, so the caller's stream order is preserved on both sides.
  RTW_CUDA_TRY(cudaEventRecord(wh->ev_in, user_stream));
  RTW_CUDA_TRY(cudaStreamWaitEvent(st, wh->ev_in, 0));
  unsigned long long* d_check = nullptr;
  EventBag sev;  // begin/end pairs around the ray-reordering passes of the instrumented path
  EventBag kev;  // begin/end pairs of the instrumented path: traverse, shade, traverse, shade ...
  uint32_t launches = 0, iterations = 0;
  const char* fault = getenv("RTW_FAULT_INJECT");  // tests: "capture" fails inside the graph capture, "launch" launches an invalid grid

  RTW_CUDA_TRY(cudaEventRecord(wh->ev_begin, st));
  RTW_CUDA_TRY(cudaMemsetAsync(wh->d_ctl, 0, sizeof(WaveCtl), st));
  if (slices == 1 && !f.skip_unowned) RTW_CUDA_TRY(cudaMemsetAsync(d_accum, 0, npix * 3 * sizeof(float), st));
  if (f.n_items > 0 && mega) {
    // ---- one-leaf scene: the whole frame is one launch ---------------------------------------------------------------
    const unsigned long long need_blocks = (f.n_items + 127ull) / 128ull;
    bool lean = s->dev.all_diffuse_solid != 0;
    if (const char* e = getenv("RTW_MEGA_LEAN")) lean = lean && atoi(e) != 0;  // 0: A/B against the full kernel
    int grid = (int)std::min<unsigned long long>((unsigned long long)wh->blocks_mega[lean ? 1 : 0], std::max<unsigned long long>(need_blocks, 1ull));
    if (fault && !strcmp(fault, "launch")) grid = -1;
    if (lean) k_mega_flat<true><<<grid, 128, 0, st>>>(s->dev, f, wh->partial, wh->d_ctl);
    else k_mega_flat<false><<<grid, 128, 0, st>>>(s->dev, f, wh->partial, wh->d_ctl);
    RTW_CUDA_TRY(cudaGetLastError());
    launches++;
    iterations = 1;
    pool = (uint32_t)grid * 128u;
  } else if (f.n_items > 0) {
    *wh->h_frame = f;
    RTW_CUDA_TRY(cudaMemcpyAsync(wh->d_frame, wh->h_frame, sizeof(FrameDev), cudaMemcpyHostToDevice, st));
    if (!count_trav && !time_kernels) {
      // ---- product path: a CUDA graph of BATCH iterations, launched back to back; the host looks at the live count
      // of round i-1 while round i is already running.  BATCH is even (the queue parity is back to 0 after every
      // batch).  The termination check lags one round, so a frame runs up to 2 x BATCH - 1 empty iterations at its
      // end: 16 for long frames, 4 for short ones (fewer than two work items per slot).
      GraphKey key;
      key.kind = (int)kind;
      key.batch = (f.n_items >= 2ull * pool) ? 16 : 4;
      // a small pool does not need the full persistent grid: fewer blocks launch (and drain) faster
      const int need_blocks = (int)((pool + 127u) / 128u);
      key.grid_t = std::min(wh->blocks_trav[kind], std::max(need_blocks, 1));
      key.grid_s = std::min(split ? wh->blocks_shade_split : wh->blocks_shade, std::max(need_blocks, 1));
      key.split = split ? 1 : 0;
      key.pool = pool;
      key.sort = raysort ? 1 : 0;
      if (!(wh->graph_exec && wh->graph_key == key)) {
        drop_graph(wh);
        cudaGraph_t graph = nullptr;
        RTW_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int b = 0; b < key.batch; ++b) {
          if (raysort) launch_sort((uint32_t)(b & 1));
          launch_traverse((uint32_t)(b & 1), key.grid_t);
          launch_shade((uint32_t)(b & 1), key.grid_s);
        }
        cudaError_t ce = cudaStreamEndCapture(st, &graph);  // always ends the capture, whatever happened inside
        if (ce == cudaSuccess && fault && !strcmp(fault, "capture")) ce = cudaErrorUnknown;
        cudaGraphExec_t exec = nullptr;
        if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return cuda_fail(ce, "graph capture of the wavefront batch");
        count_handle(1);
        wh->graph_exec = exec;
        wh->graph_key = key;
      }
      k_wave_init<<<(pool + 127) / 128, 128, 0, st>>>(s->dev, dfp, w);
      launches++;
      for (uint32_t i = 0;; ++i) {
        RTW_CUDA_TRY(cudaGraphLaunch(wh->graph_exec, st));
        RTW_CUDA_TRY(cudaMemcpyAsync(&wh->pinned_ctl[i & 1], wh->d_ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, st));
        RTW_CUDA_TRY(cudaEventRecord(wh->ring_ev[i & 1], st));
        launches += (2 + sort_launches) * key.batch;
        iterations += key.batch;
        if (i >= 1) {
          RTW_CUDA_TRY(cudaEventSynchronize(wh->ring_ev[(i - 1) & 1]));
          if (wh->pinned_ctl[(i - 1) & 1].count[0] == 0) break;  // absorbing: queue mode, no live path, no item left
        }
      }
    } else {
      // ---- instrumented path (traversal counters / per-kernel CUDA events): plain launches
      if (raysort && getenv("RTW_RAYSORT_CHECK")) {
        dev_malloc((void**)&d_check, 3 * sizeof(unsigned long long));
        cudaMemsetAsync(d_check, 0, 3 * sizeof(unsigned long long), st);
      }
      k_wave_init<<<(pool + 127) / 128, 128, 0, st>>>(s->dev, dfp, w);
      launches++;
      uint32_t parity = 0;
      const int batch = 8;
      for (;;) {
        for (int b = 0; b < batch; ++b) {
          cudaEvent_t e4[4] = {nullptr, nullptr, nullptr, nullptr};
          if (raysort) {
            cudaEvent_t es[2] = {nullptr, nullptr};
            if (time_kernels) {
              for (auto& e : es) RTW_CUDA_TRY(sev.add(&e));
              RTW_CUDA_TRY(cudaEventRecord(es[0], st));
            }
            launch_sort(parity);
            if (time_kernels) RTW_CUDA_TRY(cudaEventRecord(es[1], st));
          }
          if (time_kernels) {
            for (auto& e : e4) RTW_CUDA_TRY(kev.add(&e));
            RTW_CUDA_TRY(cudaEventRecord(e4[0], st));
          }
          if (raysort && d_check) k_raysort_check<<<1184, 256, 0, st>>>(&wh->d_ctl->count[parity], wh->rsort.key, w.order, d_check);
          launch_traverse(parity, wh->blocks_trav[kind]);
          if (time_kernels) {
            RTW_CUDA_TRY(cudaEventRecord(e4[1], st));
            RTW_CUDA_TRY(cudaEventRecord(e4[2], st));
          }
          launch_shade(parity, split ? wh->blocks_shade_split : wh->blocks_shade);
          if (time_kernels) RTW_CUDA_TRY(cudaEventRecord(e4[3], st));
          parity ^= 1;
          launches += 2 + sort_launches;
          iterations++;
        }
        RTW_CUDA_TRY(cudaMemcpyAsync(wh->pinned_ctl, wh->d_ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, st));
        RTW_CUDA_TRY(cudaStreamSynchronize(st));
        if (wh->pinned_ctl->count[parity] == 0) break;
      }
    }
  }
  if (slices > 1) {
    if (f.n_items == 0 && partial_elems) RTW_CUDA_TRY(cudaMemsetAsync(wh->partial, 0, partial_elems * sizeof(float4), st));
    k_wave_resolve<<<std::min<uint32_t>((uint32_t)((npix + 255) / 256), 8u * (uint32_t)s->num_sms), 256, 0, st>>>(f, wh->partial);
    launches++;
  }
  RTW_CUDA_TRY(cudaGetLastError());
  RTW_CUDA_TRY(cudaEventRecord(wh->ev_end, st));
  RTW_CUDA_TRY(cudaMemcpyAsync(&wh->pinned_ctl[2], wh->d_ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, st));
  RTW_CUDA_TRY(cudaStreamSynchronize(st));
  if (d_check) {
    unsigned long long hc[3] = {0, 0, 0};
    cudaMemcpy(hc, d_check, sizeof(hc), cudaMemcpyDeviceToHost);
    mem_free(d_check);
    fprintf(stdout, "[raysort check] inversions %llu, key changes %llu over %llu sorted entries\n", hc[0], hc[1], hc[2]);
  }
  const WaveCtl total = wh->pinned_ctl[2];
  float ms = 0.f;
  cudaEventElapsedTime(&ms, wh->ev_begin, wh->ev_end);
  float ms_t = 0.f, ms_s = 0.f;
  const bool trace_iterations = getenv("RTW_TRACE_ITERATIONS") != nullptr;  // debug: per-iteration kernel times of the instrumented path
  for (size_t i = 0; i + 3 < kev.v.size(); i += 4) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, kev.v[i], kev.v[i + 1]);
    cudaEventElapsedTime(&b, kev.v[i + 2], kev.v[i + 3]);
    ms_t += a;
    ms_s += b;
    if (trace_iterations) fprintf(stdout, "[iteration %zu] traverse %.3f ms  shade %.3f ms\n", i / 4, a, b);
  }
  float ms_sort = 0.f;
  for (size_t i = 0; i + 1 < sev.v.size(); i += 2) {
    float a = 0.f;
    cudaEventElapsedTime(&a, sev.v[i], sev.v[i + 1]);
    ms_sort += a;
  }
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
    stats->node_record_bytes = mega ? 0.f : (compact ? 32.f : (wide ? 128.f : 64.f));
    stats->fused = mega ? 1u : 0u;
    stats->ms_sort = ms_sort;
    stats->ray_sort = raysort ? 1u : 0u;
  }
  return RTW_OK;
}

}  // namespace rtw
