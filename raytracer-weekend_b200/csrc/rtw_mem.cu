// Device / pinned memory for the scene, build scratch and render scratch, with a process-wide cache of freed blocks.
//
// Why.  A front end that renders frame after frame from host buffers (console_app per camera, bench.py's e2e: flatten ->
// rtw_build -> rtw_render -> destroy per step) creates and destroys a scene per frame.  cudaMalloc / cudaFree of the
// ~40 buffers involved — among them the 0.8 GB wavefront pool — cost 30-900 ms per scene on this pool's boxes (measured:
// tools/e2e_probe.py; cudaFree of a large block unmaps it and synchronises the device), against 8-35 ms of rendering for
// the small configs.  Freed blocks are therefore kept per (device, size class) and handed out again; the driver is only
// asked when the cache has nothing that fits, and the cache is dropped and the request retried when the driver is out of
// memory.  Every block is only ever recycled after the call that used it has synchronised (scene destruction, end of a
// build), so no stream ordering is needed.  RTW_MEM_CACHE=0 turns the cache off; RTW_MEM_CACHE_GB caps it (default 24).
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "rtw_scene.cuh"

namespace rtw {

namespace {

struct Block {
  size_t bytes;
  int device;  // -1: pinned host memory
};

std::mutex g_mu;
std::map<std::pair<int, size_t>, std::vector<void*>> g_free;  // (device, size class) -> blocks
std::unordered_map<void*, Block> g_live;                      // blocks handed out
size_t g_cached_bytes = 0;

size_t size_class(size_t bytes) {
  if (bytes <= 256) return 256;
  if (bytes >= (1u << 20)) return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
  size_t c = 256;
  while (c < bytes) c <<= 1;
  return c;
}

bool cache_enabled() {
  static const bool on = [] {
    const char* e = getenv("RTW_MEM_CACHE");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

size_t cache_cap() {
  static const size_t cap = [] {
    const char* e = getenv("RTW_MEM_CACHE_GB");
    return (size_t)(e ? atof(e) : 24.0) << 30;
  }();
  return cap;
}

cudaError_t raw_alloc(void** p, size_t bytes, int device) {
  return device < 0 ? cudaMallocHost(p, bytes) : cudaMalloc(p, bytes);
}
void raw_free(void* p, int device) {
  if (device < 0) cudaFreeHost(p);
  else cudaFree(p);
}

void drop_cache_locked() {
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_free) {
    if (kv.first.first >= 0) cudaSetDevice(kv.first.first);
    for (void* p : kv.second) raw_free(p, kv.first.first);
  }
  g_free.clear();
  g_cached_bytes = 0;
  cudaSetDevice(cur);
}

cudaError_t alloc_impl(void** out, size_t bytes, int device) {
  const size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_free.find({device, cls});
    if (it != g_free.end() && !it->second.empty()) {
      *out = it->second.back();
      it->second.pop_back();
      g_cached_bytes -= cls;
      g_live[*out] = Block{cls, device};
      return cudaSuccess;
    }
  }
  cudaError_t e = raw_alloc(out, cls, device);
  if (e != cudaSuccess) {  // out of memory with blocks parked in the cache: give them back and try once more
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_mu);
    drop_cache_locked();
    e = raw_alloc(out, cls, device);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_live[*out] = Block{cls, device};
  }
  return e;
}

}  // namespace

cudaError_t dev_malloc(void** out, size_t bytes) {
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  return alloc_impl(out, bytes ? bytes : 1, device);
}

cudaError_t pinned_malloc(void** out, size_t bytes) { return alloc_impl(out, bytes ? bytes : 1, -1); }

// p may come from dev_malloc or pinned_malloc (on any device); nullptr is ignored
void mem_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_live.find(p);
  if (it == g_live.end()) {  // not ours (cannot happen): hand it to the driver
    cudaFree(p);
    return;
  }
  const Block b = it->second;
  g_live.erase(it);
  if (cache_enabled() && g_cached_bytes + b.bytes <= cache_cap()) {
    g_free[{b.device, b.bytes}].push_back(p);
    g_cached_bytes += b.bytes;
    return;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  if (b.device >= 0 && b.device != cur) cudaSetDevice(b.device);
  raw_free(p, b.device);
  if (b.device >= 0 && b.device != cur) cudaSetDevice(cur);
}

size_t mem_trim() {
  std::lock_guard<std::mutex> lk(g_mu);
  const size_t released = g_cached_bytes;
  drop_cache_locked();
  return released;
}

}  // namespace rtw
