// rtw_trace_closest: closest hit of a ray batch — the parity entry point.
// One thread per ray; AoS rtw_ray in, AoS rtw_hit out (the full HitRecord of hittable/mod.rs:22-29).
#include "rtw_scene.cuh"
#include "rtw_traverse.cuh"

namespace rtw {

namespace {

template <int MODE>
__global__ void __launch_bounds__(128) k_trace_batch(SceneDev sc, const rtw_ray* __restrict__ rays, uint64_t n,
                                                     rtw_hit* __restrict__ hits) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* rp = reinterpret_cast<const float*>(rays + i);
  v3 o = mk(rp[0], rp[1], rp[2]), d = mk(rp[3], rp[4], rp[5]);
  float time = rp[6], t_min = rp[7], t_max = rp[8];
  rtw_hit out;
  out.prim_id = -1; out.material_id = -1; out.t = 0.f;
  out.p[0] = out.p[1] = out.p[2] = 0.f;
  out.normal[0] = out.normal[1] = out.normal[2] = 0.f;
  out.u = out.v = 0.f; out.front_face = 0;
  float best_t;
  uint32_t meta;
  int32_t id = -1;
  const float4* g = nullptr;
  if (MODE == RTW_TRACE_BVH) {
    int32_t slot;
    TraverseCounters cnt;
    traverse_closest<false>(sc, o, d, time, t_min, t_max, slot, best_t, meta, cnt);
    if (slot >= 0) { id = sc.slot_prim[slot]; g = sc.geom + 3 * (size_t)slot; }
  } else {
    brute_closest(sc, o, d, time, t_min, t_max, id, best_t, meta);
    if (id >= 0) g = sc.raw_geom + 3 * (size_t)id;
  }
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

}  // namespace

int trace_closest_device(rtw_scene* s, const rtw_ray* d_rays, uint64_t n, rtw_hit* d_hits, int mode, cudaStream_t st) {
  if (n == 0) return RTW_OK;
  const uint32_t T = 128;
  uint64_t blocks = (n + T - 1) / T;
  if (blocks > 0x7FFFFFFFull) return set_error(RTW_ERR_INVALID, "trace: batch too large");
  if (mode == RTW_TRACE_BVH)
    k_trace_batch<RTW_TRACE_BVH><<<(uint32_t)blocks, T, 0, st>>>(s->dev, d_rays, n, d_hits);
  else if (mode == RTW_TRACE_BRUTE)
    k_trace_batch<RTW_TRACE_BRUTE><<<(uint32_t)blocks, T, 0, st>>>(s->dev, d_rays, n, d_hits);
  else
    return set_error(RTW_ERR_INVALID, "trace: unknown mode");
  RTW_CUDA_TRY(cudaGetLastError());
  return RTW_OK;
}

}  // namespace rtw
