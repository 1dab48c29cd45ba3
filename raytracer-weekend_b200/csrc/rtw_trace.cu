// rtw_trace_closest: closest hit of a ray batch — the parity entry point.
// One thread per ray; AoS rtw_ray in, AoS rtw_hit out (the full HitRecord of hittable/mod.rs:22-29).
#include <algorithm>
#include <cstdlib>

#include "rtw_scene.cuh"
#include "rtw_traverse.cuh"

namespace rtw {

namespace {

__device__ __forceinline__ void write_hit(const SceneDev& sc, rtw_hit* __restrict__ hits, uint64_t i, v3 o, v3 d, float time,
                                          int32_t id, const float4* g, float best_t, uint32_t meta) {
  rtw_hit out;
  out.prim_id = -1; out.material_id = -1; out.t = 0.f;
  out.p[0] = out.p[1] = out.p[2] = 0.f;
  out.normal[0] = out.normal[1] = out.normal[2] = 0.f;
  out.u = out.v = 0.f; out.front_face = 0;
  if (id >= 0) {
    HitRec rec;
    finalize_hit(sc, meta & 7u, meta >> RTW_META_TYPE_BITS, g, sc.prim_shade[id], o, d, time, best_t, true, rec);
    out.prim_id = id;
    out.material_id = (int32_t)sc.prim_mat[id];
    out.t = rec.t;
    out.p[0] = rec.p.x; out.p[1] = rec.p.y; out.p[2] = rec.p.z;
    out.normal[0] = rec.normal.x; out.normal[1] = rec.normal.y; out.normal[2] = rec.normal.z;
    out.u = rec.u; out.v = rec.v;
    out.front_face = rec.front ? 1 : 0;
  }
  hits[i] = out;
}

struct BatchIO {
  const SceneDev& sc;
  const rtw_ray* __restrict__ rays;
  rtw_hit* __restrict__ hits;
  static constexpr int kSuspendLanes = 0;  // a ray batch is answered in one launch
  __device__ __forceinline__ void suspend(uint32_t, int32_t, float) {}
  __device__ __forceinline__ bool load(uint32_t i, v3& o, v3& d, float& time, float& t_min, float& t_max, int32_t&, bool&) {
    const float* rp = reinterpret_cast<const float*>(rays + i);
    o = mk(rp[0], rp[1], rp[2]); d = mk(rp[3], rp[4], rp[5]);
    time = rp[6]; t_min = rp[7]; t_max = rp[8];
    return true;
  }
  // a bare ray batch has no pixel / sample: medium draws use the all-zero key (the oracle does the same)
  __device__ __forceinline__ void rng_key(Rng& rng) { rng.begin(0, 0, 0, 0); }
  __device__ __forceinline__ void store(uint32_t i, v3 o, v3 d, float time, int32_t slot, float t, uint32_t meta) {
    int32_t id = slot >= 0 ? sc.slot_prim[slot] : -1;
    write_hit(sc, hits, i, o, d, time, id, sc.geom + 3 * (size_t)(slot >= 0 ? slot : 0), t, meta);
  }
};

// the product path: the same persistent traversal the wavefront renderer runs
template <bool MEDIA, int NODES>
__global__ void __launch_bounds__(128) k_trace_bvh(SceneDev sc, const rtw_ray* __restrict__ rays, uint32_t n,
                                                   rtw_hit* __restrict__ hits, uint32_t* cursor) {
  BatchIO io{sc, rays, hits};
  TraverseCounters cnt;
#if RTW_TOP_TREE > 0
  __shared__ float4 top_smem[4 * RTW_TOP_TREE];
  stage_top_tree(sc, top_smem);
  traverse_persistent<false, MEDIA, NODES>(sc, io, n, cursor, cnt, top_smem);
#else
  traverse_persistent<false, MEDIA, NODES>(sc, io, n, cursor, cnt);
#endif
}

// flat scenes (one leaf of <= 32 primitives): the same kernel body the wavefront renderer runs for them
__global__ void __launch_bounds__(128) k_trace_flat(SceneDev sc, const rtw_ray* __restrict__ rays, uint32_t n,
                                                    rtw_hit* __restrict__ hits, uint32_t* cursor) {
  __shared__ FlatRecords fr;
  stage_flat(sc, fr);
  BatchIO io{sc, rays, hits};
  traverse_flat(sc, fr, io, n, cursor);
}

__global__ void __launch_bounds__(128) k_trace_brute(SceneDev sc, const rtw_ray* __restrict__ rays, uint64_t n,
                                                     rtw_hit* __restrict__ hits) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* rp = reinterpret_cast<const float*>(rays + i);
  v3 o = mk(rp[0], rp[1], rp[2]), d = mk(rp[3], rp[4], rp[5]);
  float time = rp[6], t_min = rp[7], t_max = rp[8];
  float best_t;
  uint32_t meta;
  int32_t id;
  Rng rng;
  rng.begin(0, 0, 0, 0);
  brute_closest(sc, o, d, time, t_min, t_max, rng, id, best_t, meta);
  write_hit(sc, hits, i, o, d, time, id, sc.raw_geom + 3 * (size_t)(id >= 0 ? id : 0), best_t, meta);
}

}  // namespace

int trace_closest_device(rtw_scene* s, const rtw_ray* d_rays, uint64_t n, rtw_hit* d_hits, int mode, cudaStream_t st) {
  if (n == 0) return RTW_OK;
  const uint32_t T = 128;
  if (mode == RTW_TRACE_BVH) {
    // persistent grid; batches above 2^31 rays are split
    static thread_local int blocks_per_sm = 0;
    if (!blocks_per_sm) {
      RTW_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_trace_bvh<true, NODES_PAIR>, (int)T, 0));
      blocks_per_sm = blocks_per_sm > 0 ? blocks_per_sm : 1;
    }
    // same switch as the renderer (rtw_render.cu): the 4-wide walk is an experiment, off unless RTW_WIDE=1
    bool wide = false;
    if (const char* e = getenv("RTW_WIDE"))
      wide = atoi(e) != 0 && s->dev.nodes4 != nullptr && !s->dev.has_media && 3u * (s->bvh_height / 2u + 1u) + 2u <= RTW_STACK_SIZE;
    bool flat = s->dev.flat_count > 0 && !s->dev.has_media;
    if (const char* e = getenv("RTW_FLAT")) flat = flat && atoi(e) != 0;
    uint32_t* cursor = nullptr;
    RTW_CUDA_TRY(cudaMallocAsync((void**)&cursor, sizeof(uint32_t), st));
    for (uint64_t done = 0; done < n;) {
      uint32_t chunk = (uint32_t)std::min<uint64_t>(n - done, 1ull << 30);
      RTW_CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(uint32_t), st));
      uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)blocks_per_sm * s->num_sms, (chunk + T - 1) / T);
      if (s->dev.has_media)
        k_trace_bvh<true, NODES_PAIR><<<grid, T, 0, st>>>(s->dev, d_rays + done, chunk, d_hits + done, cursor);
      else if (flat)
        k_trace_flat<<<grid, T, 0, st>>>(s->dev, d_rays + done, chunk, d_hits + done, cursor);
      else if (s->dev.nodes_c)  // hierarchy beyond the caches: compact pairs (rtw_bvh.cu)
        k_trace_bvh<false, NODES_COMPACT><<<grid, T, 0, st>>>(s->dev, d_rays + done, chunk, d_hits + done, cursor);
      else if (wide)
        k_trace_bvh<false, NODES_WIDE><<<grid, T, 0, st>>>(s->dev, d_rays + done, chunk, d_hits + done, cursor);
      else
        k_trace_bvh<false, NODES_PAIR><<<grid, T, 0, st>>>(s->dev, d_rays + done, chunk, d_hits + done, cursor);
      done += chunk;
    }
    RTW_CUDA_TRY(cudaFreeAsync(cursor, st));
  } else if (mode == RTW_TRACE_BRUTE) {
    uint64_t blocks = (n + T - 1) / T;
    if (blocks > 0x7FFFFFFFull) return set_error(RTW_ERR_INVALID, "trace: batch too large");
    k_trace_brute<<<(uint32_t)blocks, T, 0, st>>>(s->dev, d_rays, n, d_hits);
  } else {
    return set_error(RTW_ERR_INVALID, "trace: unknown mode");
  }
  RTW_CUDA_TRY(cudaGetLastError());
  return RTW_OK;
}

}  // namespace rtw
