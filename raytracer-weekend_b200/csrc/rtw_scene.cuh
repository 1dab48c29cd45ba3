// Host-side scene object behind the opaque `rtw_scene` handle of include/rtw_cuda.h:
// the flattening state (what the emit calls recorded) and the device buffers it owns.
#pragma once
#include <string>
#include <vector>

#include "rtw_device.cuh"

namespace rtw {

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
  std::vector<float4> raw_geom;      // 3 per primitive (raw parameters, see rtw_bvh.cu)
  std::vector<uint32_t> prim_meta;   // type | inst << 3
  std::vector<uint32_t> prim_mat;
  std::vector<int32_t> prim_shade;   // TriShade index or -1
  std::vector<rtw::TriShade> tri_shade;

  // ---- device buffers (owned) -----------------------------------------------------------------
  std::vector<void*> allocations;
  uint64_t device_bytes = 0;
  rtw::SceneDev dev{};
  float root_box[6] = {0, 0, 0, 0, 0, 0};
  uint32_t bvh_height = 0;

  // render scratch kept between calls (rtw_render.cu)
  void* wave = nullptr;
};

namespace rtw {

// rtw_bvh.cu: upload the flattened scene and build the LBVH on the GPU.
int build_scene_device(rtw_scene* s, float time0, float time1, rtw_build_stats* stats);
void free_scene_device(rtw_scene* s);

// rtw_trace.cu
int trace_closest_device(rtw_scene* s, const rtw_ray* d_rays, uint64_t n, rtw_hit* d_hits, int mode, cudaStream_t st);

// rtw_render.cu
int render_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* d_accum, cudaStream_t st,
                  rtw_render_stats* stats);
void free_wave(rtw_scene* s);
int resolve_rgb8_device(const float* d_accum, size_t n, uint32_t spp, uint8_t* d_rgb8, cudaStream_t st);

}  // namespace rtw
