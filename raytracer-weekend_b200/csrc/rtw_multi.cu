// Multi-GPU behind the C ABI (SURVEY.md §8b `gpus`, §8e): ONE process drives N devices of the box.
//
//   rtw_scene_clone   the built scene — primitives, LBVH, materials, textures — is copied device to device
//                     (cudaMemcpyPeer: NVLink / NVSwitch between peers), never re-flattened or rebuilt: every
//                     replica holds bit-identical data ("build on GPU 0 and broadcast").
//   render_multi_device   pixels are independent (lib.rs:63): device i renders the 32x32 tiles k with k % N == i
//                     (rtw_render_params part_rank / part_count), one host thread per device, no exchange while
//                     rendering.  The merge is fused into the render: every replica's kernels store the pixels they
//                     finish straight into the ONE frame on the scene's device through peer memory (P2P stores over
//                     NVLink), so when the last device is done the frame is complete — no reduce, no gather, no
//                     zero-padded buffers.  Devices that are not peers render into a local frame that is copied
//                     across and merged by k_merge_owned (RTW_NO_PEER=1 forces that path: it is tested).
//   The random stream is keyed by (pixel, sample) and the slice count does not depend on the partition, so the
//   frame has the same bits for every N.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rtw_scene.cuh"

namespace rtw {

namespace {

// dst[pix] = src[pix] for the pixels of the tiles owned by `rank`
__global__ void k_merge_owned(const float* __restrict__ src, float* __restrict__ dst, uint32_t width, uint32_t height,
                              uint32_t tile, uint32_t tiles_x, uint32_t rank, uint32_t count) {
  const size_t npix = (size_t)width * height;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)(pix % width), y = (uint32_t)(pix / width);
    const uint32_t t = (y / tile) * tiles_x + (x / tile);
    if (t % count != rank) continue;
    dst[3 * pix] = src[3 * pix]; dst[3 * pix + 1] = src[3 * pix + 1]; dst[3 * pix + 2] = src[3 * pix + 2];
  }
}

template <class P>
void rebase(const rtw_scene* src, const rtw_scene* dst, P& p) {  // P = some `const T* __restrict__`
  if (!p) return;
  const char* c = reinterpret_cast<const char*>(p);
  for (size_t i = 0; i < src->allocations.size(); ++i) {
    const char* b = reinterpret_cast<const char*>(src->allocations[i]);
    if (c >= b && c < b + src->allocation_bytes[i]) {
      p = reinterpret_cast<P>(reinterpret_cast<char*>(dst->allocations[i]) + (c - b));
      return;
    }
  }
  p = nullptr;  // not one of the scene's buffers: cannot happen for a built scene
}

}  // namespace

int clone_scene(const rtw_scene* src, int device, rtw_scene** out) {
  int ndev = 0;
  RTW_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return set_error(RTW_ERR_INVALID, "clone: device index out of range");
  rtw_scene* d = new rtw_scene();
  d->device = device;
  d->built = true;
  d->is_replica = true;
  // host-side view of the flattening (rtw_scene_prim_info & co. work on a replica too)
  d->textures = src->textures;
  d->materials = src->materials;
  d->noise_tables = src->noise_tables;
  d->inst_range = src->inst_range;
  d->inst_ops = src->inst_ops;
  d->prim_meta = src->prim_meta;
  d->prim_mat = src->prim_mat;
  d->prim_shade = src->prim_shade;
  memcpy(d->root_box, src->root_box, sizeof(d->root_box));
  d->bvh_height = src->bvh_height;
  auto fail = [&](int rc) {
    cudaSetDevice(device);
    free_scene_device(d);
    delete d;
    return rc;
  };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(cuda_fail(e, "cudaSetDevice (clone)"));
  e = cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return fail(cuda_fail(e, "cudaDeviceGetAttribute (clone)"));
  for (size_t i = 0; i < src->allocations.size(); ++i) {
    void* p = nullptr;
    e = dev_malloc(&p, src->allocation_bytes[i]);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(set_error(RTW_ERR_NOMEM, std::string("clone: cudaMalloc failed: ") + cudaGetErrorString(e)));
    }
    d->allocations.push_back(p);
    d->allocation_bytes.push_back(src->allocation_bytes[i]);
    d->device_bytes += src->allocation_bytes[i];
    e = cudaMemcpyPeer(p, device, src->allocations[i], src->device, src->allocation_bytes[i]);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMemcpyPeer (clone)"));
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(cuda_fail(e, "cudaDeviceSynchronize (clone)"));
  SceneDev v = src->dev;
  rebase(src, d, v.nodes); rebase(src, d, v.geom); rebase(src, d, v.slot_prim); rebase(src, d, v.slot_meta);
  rebase(src, d, v.slot_ms); rebase(src, d, v.nodes_c); rebase(src, d, v.nodes4); rebase(src, d, v.top_nodes);
  rebase(src, d, v.prim_mat); rebase(src, d, v.prim_shade); rebase(src, d, v.prim_meta); rebase(src, d, v.raw_geom);
  rebase(src, d, v.tri_shade); rebase(src, d, v.inst_range); rebase(src, d, v.inst_ops); rebase(src, d, v.materials);
  rebase(src, d, v.textures); rebase(src, d, v.noise); rebase(src, d, v.texels);
  d->dev = v;
  *out = d;
  return RTW_OK;
}

void free_replicas(rtw_scene* s) {
  for (rtw_scene* r : s->replicas) {
    if (!r) continue;
    cudaSetDevice(r->device);
    free_wave(r);
    free_scene_device(r);
    if (r->io_frame) mem_free(r->io_frame);
    delete r;
  }
  s->replicas.clear();
  cudaSetDevice(s->device);
  for (float* p : s->staging)
    if (p) mem_free(p);
  s->staging.clear();
  s->staging_bytes = 0;
}

int render_multi_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* d_frame, rtw_render_stats* stats) {
  int ndev = 0;
  RTW_CUDA_TRY(cudaGetDeviceCount(&ndev));
  const uint32_t n = p->gpus;
  if (n < 2) return render_device(s, cam, p, d_frame, 0, stats);
  if ((int)n > ndev)
    return set_error(RTW_ERR_INVALID, "render: gpus = " + std::to_string(n) + " but the box has " + std::to_string(ndev) + " device(s)");
  if (p->part_count > 1) return set_error(RTW_ERR_INVALID, "render: gpus > 1 partitions the frame itself (part_count must be 0)");
  if (p->flags & (RTW_RENDER_COUNT_TRAVERSAL | RTW_RENDER_TIME_KERNELS))
    return set_error(RTW_ERR_INVALID, "render: the instrumented paths run on one device (gpus must be 0 or 1)");
  const size_t bytes = std::max<size_t>((size_t)p->width * p->height * 3 * sizeof(float), 4);
  const bool no_peer = getenv("RTW_NO_PEER") && atoi(getenv("RTW_NO_PEER")) != 0;

  // ---- replicas, peer access -----------------------------------------------------------------------------------------
  if (s->replicas.size() < n - 1) s->replicas.resize(n - 1, nullptr);
  if (s->staging.size() < n - 1) s->staging.resize(n - 1, nullptr);
  std::vector<float*> target(n, d_frame);
  std::vector<bool> staged(n, false);
  for (uint32_t i = 1; i < n; ++i) {
    const int dev = (s->device + (int)i) % ndev;
    if (!s->replicas[i - 1]) {
      int rc = clone_scene(s, dev, &s->replicas[i - 1]);
      if (rc != RTW_OK) return rc;
    }
    rtw_scene* r = s->replicas[i - 1];
    int can = 0;
    if (!no_peer) RTW_CUDA_TRY(cudaDeviceCanAccessPeer(&can, r->device, s->device));
    if (can) {
      RTW_CUDA_TRY(cudaSetDevice(r->device));
      cudaError_t e = cudaDeviceEnablePeerAccess(s->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) { cudaGetLastError(); can = 0; }
    }
    if (!can) {  // render into the replica's own frame, copy it across, merge the owned tiles
      RTW_CUDA_TRY(cudaSetDevice(r->device));
      if (r->io_bytes < bytes) {
        if (r->io_frame) mem_free(r->io_frame);
        r->io_frame = nullptr;
        r->io_bytes = 0;
        RTW_CUDA_TRY(dev_malloc((void**)&r->io_frame, bytes));
        r->io_bytes = bytes;
      }
      RTW_CUDA_TRY(cudaSetDevice(s->device));
      if (s->staging_bytes < bytes) {
        for (float*& q : s->staging) { if (q) mem_free(q); q = nullptr; }
        s->staging_bytes = bytes;
      }
      if (!s->staging[i - 1]) RTW_CUDA_TRY(dev_malloc((void**)&s->staging[i - 1], s->staging_bytes));
      target[i] = r->io_frame;
      staged[i] = true;
    }
  }
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  // peers store only the pixels they own; if any replica is staged, or nothing is rendered at all, start from zeros
  const uint32_t s0 = p->sample_begin, s1 = (p->sample_begin == 0 && p->sample_end == 0) ? p->spp : p->sample_end;
  if (s1 <= s0) {
    RTW_CUDA_TRY(cudaMemset(d_frame, 0, bytes));
    if (s1 < s0) return set_error(RTW_ERR_INVALID, "render: sample_end < sample_begin");
  }

  // ---- one host thread per device ---------------------------------------------------------------------------------------
  std::vector<rtw_render_stats> st(n);
  std::vector<int> rc(n, RTW_OK);
  std::vector<std::string> err(n);
  auto work = [&](uint32_t i) {
    rtw_scene* r = i == 0 ? s : s->replicas[i - 1];
    rtw_render_params q = *p;
    q.gpus = 0;
    q.tile_size = p->tile_size ? p->tile_size : 32;
    q.part_rank = i;
    q.part_count = n;
    rc[i] = render_device(r, cam, &q, target[i], 0, &st[i], /*skip_unowned=*/!staged[i]);
    if (rc[i] != RTW_OK) err[i] = rtw_last_error();
  };
  std::vector<std::thread> threads;
  for (uint32_t i = 1; i < n; ++i) threads.emplace_back(work, i);
  work(0);
  for (auto& t : threads) t.join();
  for (uint32_t i = 0; i < n; ++i)
    if (rc[i] != RTW_OK) return set_error(rc[i], "device " + std::to_string(i) + ": " + err[i]);
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  const uint32_t tile = p->tile_size ? p->tile_size : 32;
  const uint32_t tiles_x = (p->width + tile - 1) / tile;
  for (uint32_t i = 1; i < n; ++i) {
    if (!staged[i]) continue;
    rtw_scene* r = s->replicas[i - 1];
    RTW_CUDA_TRY(cudaMemcpyPeer(s->staging[i - 1], s->device, r->io_frame, r->device, bytes));
    const size_t npix = (size_t)p->width * p->height;
    k_merge_owned<<<(uint32_t)std::min<size_t>((npix + 255) / 256, 4096), 256>>>(s->staging[i - 1], d_frame, p->width, p->height,
                                                                                tile, tiles_x, i, n);
    RTW_CUDA_TRY(cudaGetLastError());
  }
  RTW_CUDA_TRY(cudaDeviceSynchronize());
  if (stats) {
    *stats = st[0];
    for (uint32_t i = 1; i < n; ++i) {
      stats->segments += st[i].segments;
      stats->paths += st[i].paths;
      stats->launches += st[i].launches;
      stats->iterations = std::max(stats->iterations, st[i].iterations);
      stats->ms_render = std::max(stats->ms_render, st[i].ms_render);  // devices run concurrently
    }
    stats->gpus = n;
  }
  return RTW_OK;
}

}  // namespace rtw
