// Host-side scene object behind the opaque `rtw_scene` handle of include/rtw_cuda.h:
// the flattening state (what the emit calls recorded) and the device buffers it owns.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "rtw_device.cuh"

namespace rtw {

// std::vector whose resize() leaves trivially constructible elements UNINITIALISED: the bulk ingest (rtw_add_triangles) grows
// the geometry arrays by hundreds of MB and fills every element right after — from all host threads, which then also take the
// first-touch page faults; a value-initialising resize() zeroes (and faults in) all of it on one thread first.
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_allocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
  template <class U, class... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
template <class T>
using raw_vector = std::vector<T, default_init_allocator<T>>;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// thread-local last-error plumbing (rtw_api.cu)
int set_error(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define RTW_CUDA_TRY(expr)                                        \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::rtw::cuda_fail(_e, #expr);    \
  } while (0)

}  // namespace rtw

struct rtw_scene {
  int device = 0;
  int num_sms = 0;
  bool built = false;

  // ---- flattening state (host) -------------------------------------------------------------
  std::vector<rtw::TextureRec> textures;
  std::vector<rtw::MaterialRec> materials;
  std::vector<rtw::NoiseTable> noise_tables;
  std::vector<uchar4> texels;
  // transform stack: open wrappers, outermost first
  std::vector<rtw::InstOp> open_ops;
  std::vector<int> open_kinds;       // 0 transform, 1 group   (to validate pop / end order)
  std::vector<int> open_emitted;     // primitives emitted inside each open wrapper
  // an open ConstantMedium: -1 none, else the Isotropic material id; its boundary has been emitted or not
  int medium_material = -1;
  float medium_density = 0.f;
  bool medium_has_boundary = false;
  size_t medium_open_depth = 0;      // open_kinds.size() at rtw_begin_medium
  int cur_inst = 0;                  // instance index of the current open chain
  bool cur_inst_valid = true;
  std::vector<uint2> inst_range;     // inst -> (first op, count) ; inst 0 = identity
  std::vector<rtw::InstOp> inst_ops;
  // primitives in canonical order
  rtw::raw_vector<float4> raw_geom;  // 3 per primitive (raw parameters, see rtw_bvh.cu)
  std::vector<uint32_t> prim_meta;   // type | inst << 3
  std::vector<uint32_t> prim_mat;
  std::vector<int32_t> prim_shade;   // TriShade index or -1
  std::vector<rtw::TriShade> tri_shade;

  // ---- device buffers (owned) -----------------------------------------------------------------
  std::vector<void*> allocations;
  std::vector<size_t> allocation_bytes;  // parallel to `allocations` (rtw_scene_clone copies them peer to peer)
  uint64_t device_bytes = 0;
  rtw::SceneDev dev{};
  float root_box[6] = {0, 0, 0, 0, 0, 0};
  uint32_t bvh_height = 0;

  // render scratch kept between calls (rtw_render.cu)
  void* wave = nullptr;
  // rtw_render (host buffers): frame buffer on the device, kept between calls (grow only)
  float* io_frame = nullptr;
  size_t io_bytes = 0;
  // multi-GPU (rtw_render_params::gpus > 1): replicas of this scene on the other devices, created on first use;
  // staging[i] = device memory of THIS scene's device for replica i's frame when peer stores are not possible
  std::vector<rtw_scene*> replicas;
  std::vector<float*> staging;
  size_t staging_bytes = 0;
  bool is_replica = false;
};

namespace rtw {

// rtw_mem.cu: device / pinned memory with a process-wide cache of freed blocks (scene, build and render scratch)
cudaError_t dev_malloc(void** out, size_t bytes);     // on the current device
cudaError_t pinned_malloc(void** out, size_t bytes);  // page-locked host memory
void mem_free(void* p);                               // from either of the two; nullptr ignored
size_t mem_trim();                                    // give every cached block back to the driver; bytes released

// rtw_bvh.cu: upload the flattened scene and build the LBVH on the GPU.
int build_scene_device(rtw_scene* s, float time0, float time1, rtw_build_stats* stats);
void free_scene_device(rtw_scene* s);

// rtw_trace.cu
int trace_closest_device(rtw_scene* s, const rtw_ray* d_rays, uint64_t n, rtw_hit* d_hits, int mode, cudaStream_t st);

// rtw_render.cu
// skip_unowned: pixels outside this call's tile partition are left untouched instead of being written as 0
// (the multi-GPU path: every replica stores its own tiles into ONE frame, rtw_multi.cu)
int render_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* d_accum, cudaStream_t st,
                  rtw_render_stats* stats, bool skip_unowned = false);
void free_wave(rtw_scene* s);
// handles (events, streams, graphs) this library has created and not yet destroyed: the leak check of the tests
int live_handles();
void count_handle(int delta);

// rtw_multi.cu: one frame over `gpus` devices, replicas on first use; the frame lands in d_frame (memory of s->device)
int render_multi_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* d_frame,
                        rtw_render_stats* stats);
int clone_scene(const rtw_scene* src, int device, rtw_scene** out);
void free_replicas(rtw_scene* s);
int resolve_rgb8_device(const float* d_accum, size_t n, uint32_t spp, uint8_t* d_rgb8, cudaStream_t st);

}  // namespace rtw
