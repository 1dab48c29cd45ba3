// extern "C" surface of include/rtw_cuda.h: the host-side FLATTEN step (scene emit calls ->
// SoA arrays in canonical primitive order + instance chains) and thin wrappers that move host
// buffers to the device and call the kernels.  No CPU fallback exists: every compute entry
// point needs a CUDA device and fails with RTW_ERR_CUDA otherwise.
#include <algorithm>
#include <cmath>
#include <future>
#include <cstring>
#include <string>
#include <thread>

#include "rtw_scene.cuh"

namespace rtw {

static thread_local std::string g_last_error;

int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  cudaGetLastError();  // clear the sticky-less error state
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return RTW_ERR_CUDA;
}

namespace {

bool valid_tex(const rtw_scene* s, int t) { return t >= 0 && t < (int)s->textures.size(); }
bool valid_mat(const rtw_scene* s, int m) { return m >= 0 && m < (int)s->materials.size(); }

// instance index of the currently open wrapper chain (created on first use)
int current_instance(rtw_scene* s) {
  if (s->open_ops.empty()) return 0;
  if (!s->cur_inst_valid) {
    uint2 rg;
    rg.x = (uint32_t)s->inst_ops.size();
    rg.y = (uint32_t)s->open_ops.size();
    s->inst_ops.insert(s->inst_ops.end(), s->open_ops.begin(), s->open_ops.end());
    s->inst_range.push_back(rg);
    s->cur_inst = (int)s->inst_range.size() - 1;
    s->cur_inst_valid = true;
  }
  return s->cur_inst;
}

int emit_prim(rtw_scene* s, uint32_t type, float4 g0, float4 g1, float4 g2, int material, int shade) {
  if (s->medium_material >= 0) {  // inside rtw_begin_medium: only the boundary (sphere / cuboid) may come
    if (type != PT_MEDIUM_SPHERE && type != PT_MEDIUM_BOX)
      return set_error(RTW_ERR_UNSUPPORTED, "medium: the boundary must be one sphere or one cuboid");
    if (s->medium_has_boundary) return set_error(RTW_ERR_INVALID, "medium: a medium needs exactly one boundary object");
    s->medium_has_boundary = true;
  }
  if (s->prim_meta.size() >= (size_t)0x0FFFFFFF) return set_error(RTW_ERR_UNSUPPORTED, "too many primitives");
  int inst = current_instance(s);
  int id = (int)s->prim_meta.size();
  s->raw_geom.push_back(g0);
  s->raw_geom.push_back(g1);
  s->raw_geom.push_back(g2);
  s->prim_meta.push_back(type | ((uint32_t)inst << RTW_META_TYPE_BITS));
  s->prim_mat.push_back((uint32_t)material);
  s->prim_shade.push_back(shade);
  return id;
}

}  // namespace
}  // namespace rtw

using namespace rtw;

#define CHECK_OPEN(s)                                                        \
  if (!(s)) return set_error(RTW_ERR_INVALID, "scene is NULL");              \
  if ((s)->built) return set_error(RTW_ERR_STATE, "scene already built")
#define CHECK_BUILT(s)                                                       \
  if (!(s)) return set_error(RTW_ERR_INVALID, "scene is NULL");              \
  if (!(s)->built) return set_error(RTW_ERR_STATE, "scene not built (call rtw_build)")

extern "C" {

int rtw_abi_version(void) { return RTW_ABI_VERSION; }
const char* rtw_last_error(void) { return g_last_error.c_str(); }

int rtw_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  return n;
}

int rtw_scene_create(int device, rtw_scene** out) {
  if (!out) return set_error(RTW_ERR_INVALID, "out is NULL");
  if (device < 0) return set_error(RTW_ERR_INVALID, "device < 0");
  rtw_scene* s = new rtw_scene();
  s->device = device;
  uint2 ident;
  ident.x = 0; ident.y = 0;
  s->inst_range.push_back(ident);
  *out = s;
  return RTW_OK;
}

int rtw_scene_destroy(rtw_scene* s) {
  if (!s) return RTW_OK;
  if (s->built || s->wave) {
    free_replicas(s);
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();  // (cudaFree used to imply this) the scene's blocks go back to the cache of rtw_mem.cu
    free_wave(s);
    free_scene_device(s);
    if (s->io_frame) mem_free(s->io_frame);
  }
  delete s;
  return RTW_OK;
}

int rtw_scene_clone(const rtw_scene* src, int device, rtw_scene** out) {
  if (!src || !out) return set_error(RTW_ERR_INVALID, "clone: NULL argument");
  if (!src->built) return set_error(RTW_ERR_STATE, "scene not built (call rtw_build)");
  int rc = clone_scene(src, device, out);
  if (rc == RTW_OK) (*out)->is_replica = false;  // a caller-owned scene like any other
  cudaSetDevice(src->device);
  return rc;
}

int rtw_debug_live_handles(void) { return live_handles(); }

int rtw_trim_memory(void) { return (int)(mem_trim() >> 20); }

// ---- textures ----------------------------------------------------------------------------------
int rtw_add_texture_solid(rtw_scene* s, float r, float g, float b) {
  CHECK_OPEN(s);
  TextureRec t{};
  t.type = TT_SOLID;
  t.f0 = r; t.f1 = g; t.f2 = b;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int rtw_add_texture_checker(rtw_scene* s, int odd, int even, float frequency) {
  CHECK_OPEN(s);
  if (!valid_tex(s, odd) || !valid_tex(s, even)) return set_error(RTW_ERR_INVALID, "checker: bad texture id");
  TextureRec t{};
  t.type = TT_CHECKER;
  t.i0 = odd; t.i1 = even; t.f0 = frequency;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int rtw_add_texture_noise(rtw_scene* s, const float* g, const int32_t* px, const int32_t* py, const int32_t* pz,
                          float scale) {
  CHECK_OPEN(s);
  if (!g || !px || !py || !pz) return set_error(RTW_ERR_INVALID, "noise: NULL table");
  NoiseTable nt;
  for (int i = 0; i < 256; ++i) {
    if ((px[i] | py[i] | pz[i]) & ~255) return set_error(RTW_ERR_INVALID, "noise: permutation value out of 0..255");
    nt.grad[i] = make_float4(g[3 * i], g[3 * i + 1], g[3 * i + 2], 0.f);
    nt.perm[0][i] = (uint8_t)px[i];
    nt.perm[1][i] = (uint8_t)py[i];
    nt.perm[2][i] = (uint8_t)pz[i];
  }
  TextureRec t{};
  t.type = TT_NOISE;
  t.i0 = (int)s->noise_tables.size();
  t.f0 = scale;
  s->noise_tables.push_back(nt);
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int rtw_add_texture_uvdebug(rtw_scene* s) {
  CHECK_OPEN(s);
  TextureRec t{};
  t.type = TT_UVDEBUG;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}
int rtw_add_texture_image(rtw_scene* s, const uint8_t* rgb8, uint32_t width, uint32_t height) {
  CHECK_OPEN(s);
  if (!rgb8 || width == 0 || height == 0) return set_error(RTW_ERR_INVALID, "image: empty");
  if ((uint64_t)width * height + s->texels.size() > 0x7FFFFFFFull) return set_error(RTW_ERR_UNSUPPORTED, "image: texel pool full");
  TextureRec t{};
  t.type = TT_IMAGE;
  t.i0 = (int)s->texels.size();
  t.i1 = (int)width;
  t.i2 = (int)height;
  const size_t n = (size_t)width * height, base = s->texels.size();
  s->texels.resize(base + n);
  uchar4* dst = s->texels.data() + base;
  for (size_t i = 0; i < n; ++i) dst[i] = make_uchar4(rgb8[3 * i], rgb8[3 * i + 1], rgb8[3 * i + 2], 255);
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}

// ---- materials ---------------------------------------------------------------------------------
int rtw_add_material_lambertian(rtw_scene* s, int tex) {
  CHECK_OPEN(s);
  if (!valid_tex(s, tex)) return set_error(RTW_ERR_INVALID, "lambertian: bad texture id");
  MaterialRec m{};
  m.type = MT_LAMBERTIAN;
  m.tex = tex;
  s->materials.push_back(m);
  return (int)s->materials.size() - 1;
}
int rtw_add_material_metal(rtw_scene* s, float r, float g, float b, float fuzz) {
  CHECK_OPEN(s);
  if (!(fuzz <= 1.0f)) return set_error(RTW_ERR_INVALID, "metal: fuzz must be <= 1 (material.rs:71)");
  MaterialRec m{};
  m.type = MT_METAL;
  m.tex = -1;
  m.param = fuzz;
  m.r = r; m.g = g; m.b = b;
  s->materials.push_back(m);
  return (int)s->materials.size() - 1;
}
int rtw_add_material_dielectric(rtw_scene* s, float ir) {
  CHECK_OPEN(s);
  MaterialRec m{};
  m.type = MT_DIELECTRIC;
  m.tex = -1;
  m.param = ir;
  s->materials.push_back(m);
  return (int)s->materials.size() - 1;
}
int rtw_add_material_diffuse_light(rtw_scene* s, int tex) {
  CHECK_OPEN(s);
  if (!valid_tex(s, tex)) return set_error(RTW_ERR_INVALID, "diffuse_light: bad texture id");
  MaterialRec m{};
  m.type = MT_DIFFUSE_LIGHT;
  m.tex = tex;
  s->materials.push_back(m);
  return (int)s->materials.size() - 1;
}

// ---- wrappers ------------------------------------------------------------------------------------
static int push_op(rtw_scene* s, InstOp op) {
  if (s->open_ops.size() >= RTW_MAX_CHAIN) return set_error(RTW_ERR_UNSUPPORTED, "transform nesting deeper than 8");
  s->open_ops.push_back(op);
  s->open_kinds.push_back(0);
  s->open_emitted.push_back((int)s->prim_meta.size());
  s->cur_inst_valid = false;
  return RTW_OK;
}
int rtw_push_translation(rtw_scene* s, const float offset[3]) {
  CHECK_OPEN(s);
  if (!offset) return set_error(RTW_ERR_INVALID, "offset is NULL");
  InstOp op;
  op.kind = OP_TRANSLATE;
  op.a = offset[0]; op.b = offset[1]; op.c = offset[2];
  return push_op(s, op);
}
int rtw_push_rotation_y(rtw_scene* s, float angle_degrees) {
  CHECK_OPEN(s);
  // transformations.rs:59-63 ; f32::to_radians = x * (PI / 180) with the constant folded in f32
  const float RADS_PER_DEG = 3.14159274101257324219f / 180.0f;
  float angle_radians = angle_degrees * RADS_PER_DEG;
  InstOp op;
  op.kind = OP_ROTY;
  op.a = std::sin(angle_radians);
  op.b = std::cos(angle_radians);
  op.c = 0.f;
  return push_op(s, op);
}
int rtw_push_rotation_y_sincos(rtw_scene* s, float sin_theta, float cos_theta) {
  CHECK_OPEN(s);
  InstOp op;
  op.kind = OP_ROTY;
  op.a = sin_theta;
  op.b = cos_theta;
  op.c = 0.f;
  return push_op(s, op);
}
int rtw_pop_transform(rtw_scene* s) {
  CHECK_OPEN(s);
  if (s->open_kinds.empty() || s->open_kinds.back() != 0) return set_error(RTW_ERR_STATE, "pop_transform: no open transform");
  if (s->open_emitted.back() == (int)s->prim_meta.size()) return set_error(RTW_ERR_INVALID, "pop_transform: empty instance");
  s->open_ops.pop_back();
  s->open_kinds.pop_back();
  s->open_emitted.pop_back();
  s->cur_inst_valid = false;
  return RTW_OK;
}
int rtw_begin_group(rtw_scene* s) {
  CHECK_OPEN(s);
  s->open_kinds.push_back(1);
  s->open_emitted.push_back((int)s->prim_meta.size());
  return RTW_OK;
}
int rtw_end_group(rtw_scene* s) {
  CHECK_OPEN(s);
  if (s->open_kinds.empty() || s->open_kinds.back() != 1) return set_error(RTW_ERR_STATE, "end_group: no open group");
  s->open_kinds.pop_back();
  s->open_emitted.pop_back();
  return RTW_OK;
}

// ---- primitives ----------------------------------------------------------------------------------
// ---- media ---------------------------------------------------------------------------------------
int rtw_begin_medium(rtw_scene* s, float density, int texture) {
  CHECK_OPEN(s);
  if (!valid_tex(s, texture)) return set_error(RTW_ERR_INVALID, "medium: bad texture id");
  if (s->medium_material >= 0) return set_error(RTW_ERR_UNSUPPORTED, "medium: nested media are not supported");
  if (!s->open_ops.empty())
    return set_error(RTW_ERR_UNSUPPORTED, "medium: a transformed medium is not supported (transform the boundary)");
  MaterialRec m{};
  m.type = MT_ISOTROPIC;  // Isotropic::new(texture) (volumes.rs:26, material.rs:149-152)
  m.tex = texture;
  s->materials.push_back(m);
  s->medium_material = (int)s->materials.size() - 1;
  s->medium_density = density;
  s->medium_has_boundary = false;
  s->medium_open_depth = s->open_kinds.size();
  return (int)s->prim_meta.size();
}
int rtw_end_medium(rtw_scene* s) {
  CHECK_OPEN(s);
  if (s->medium_material < 0) return set_error(RTW_ERR_STATE, "end_medium: no open medium");
  if (s->open_kinds.size() != s->medium_open_depth) return set_error(RTW_ERR_STATE, "end_medium: unbalanced push inside the medium");
  if (!s->medium_has_boundary) return set_error(RTW_ERR_INVALID, "end_medium: a medium needs exactly one boundary object");
  s->medium_material = -1;
  return RTW_OK;
}

int rtw_add_sphere(rtw_scene* s, const float c[3], float radius, int material) {
  CHECK_OPEN(s);
  if (!c || !valid_mat(s, material)) return set_error(RTW_ERR_INVALID, "sphere: bad argument");
  if (s->medium_material >= 0)  // boundary of a ConstantMedium: volumes.rs:24-35 (neg_inv_density = -1/density)
    return emit_prim(s, PT_MEDIUM_SPHERE, make_float4(c[0], c[1], c[2], radius), make_float4(-1.0f / s->medium_density, 0, 0, 0),
                     make_float4(0, 0, 0, 0), s->medium_material, -1);
  return emit_prim(s, PT_SPHERE, make_float4(c[0], c[1], c[2], radius), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0),
                   material, -1);
}
int rtw_add_moving_sphere(rtw_scene* s, const float c0[3], float time0, const float c1[3], float time1, float radius,
                          int material) {
  CHECK_OPEN(s);
  if (!c0 || !c1 || !valid_mat(s, material)) return set_error(RTW_ERR_INVALID, "moving_sphere: bad argument");
  return emit_prim(s, PT_MSPHERE, make_float4(c0[0], c0[1], c0[2], radius), make_float4(c1[0], c1[1], c1[2], time0),
                   make_float4(time1, 0, 0, 0), material, -1);
}
static int add_rect(rtw_scene* s, uint32_t type, float a0, float a1, float b0, float b1, float k, int material) {
  CHECK_OPEN(s);
  if (!valid_mat(s, material)) return set_error(RTW_ERR_INVALID, "rect: bad material id");
  return emit_prim(s, type, make_float4(a0, a1, b0, b1), make_float4(k, 0, 0, 0), make_float4(0, 0, 0, 0), material, -1);
}
int rtw_add_xy_rect(rtw_scene* s, float x0, float x1, float y0, float y1, float k, int m) { return add_rect(s, PT_RECT_XY, x0, x1, y0, y1, k, m); }
int rtw_add_xz_rect(rtw_scene* s, float x0, float x1, float z0, float z1, float k, int m) { return add_rect(s, PT_RECT_XZ, x0, x1, z0, z1, k, m); }
int rtw_add_yz_rect(rtw_scene* s, float y0, float y1, float z0, float z1, float k, int m) { return add_rect(s, PT_RECT_YZ, y0, y1, z0, z1, k, m); }
int rtw_add_cuboid(rtw_scene* s, const float p0[3], const float p1[3], int material) {
  CHECK_OPEN(s);
  if (!p0 || !p1 || !valid_mat(s, material)) return set_error(RTW_ERR_INVALID, "cuboid: bad argument");
  if (s->medium_material >= 0)
    return emit_prim(s, PT_MEDIUM_BOX, make_float4(p0[0], p0[1], p0[2], p1[0]),
                     make_float4(p1[1], p1[2], -1.0f / s->medium_density, 0), make_float4(0, 0, 0, 0), s->medium_material, -1);
  // rectangular.rs:177-234: XY@z1, XY@z0, XZ@y1, XZ@y0, YZ@x1, YZ@x0
  int first = add_rect(s, PT_RECT_XY, p0[0], p1[0], p0[1], p1[1], p1[2], material);
  if (first < 0) return first;
  add_rect(s, PT_RECT_XY, p0[0], p1[0], p0[1], p1[1], p0[2], material);
  add_rect(s, PT_RECT_XZ, p0[0], p1[0], p0[2], p1[2], p1[1], material);
  add_rect(s, PT_RECT_XZ, p0[0], p1[0], p0[2], p1[2], p0[1], material);
  add_rect(s, PT_RECT_YZ, p0[1], p1[1], p0[2], p1[2], p1[0], material);
  add_rect(s, PT_RECT_YZ, p0[1], p1[1], p0[2], p1[2], p0[0], material);
  return first;
}
int rtw_add_triangles(rtw_scene* s, uint32_t n, const float* vertices, const float* normals, const float* uvs,
                      const int32_t* material_ids, int material) {
  CHECK_OPEN(s);
  if (n == 0) return (int)s->prim_meta.size();
  if (!vertices) return set_error(RTW_ERR_INVALID, "triangles: vertices is NULL");
  if (material_ids) {
    for (uint32_t i = 0; i < n; ++i)
      if (!valid_mat(s, material_ids[i])) return set_error(RTW_ERR_INVALID, "triangles: bad material id");
  } else if (!valid_mat(s, material)) {
    return set_error(RTW_ERR_INVALID, "triangles: bad material id");
  }
  if (s->prim_meta.size() + (size_t)n >= (size_t)0x0FFFFFFF) return set_error(RTW_ERR_UNSUPPORTED, "too many primitives");
  if (s->medium_material >= 0) return set_error(RTW_ERR_UNSUPPORTED, "medium: the boundary must be one sphere or one cuboid");
  const int first = (int)s->prim_meta.size();
  const bool shaded = normals || uvs;
  const uint32_t meta = PT_TRI | ((uint32_t)current_instance(s) << RTW_META_TYPE_BITS);
  const size_t shade0 = s->tri_shade.size();
  // bulk ingest: the arrays grow once and are filled by all host threads (a 10 M-triangle mesh is 0.5 GB of copies)
  s->raw_geom.resize(s->raw_geom.size() + 3 * (size_t)n);
  s->prim_meta.resize(s->prim_meta.size() + n, meta);
  s->prim_mat.resize(s->prim_mat.size() + n, (uint32_t)material);
  s->prim_shade.resize(s->prim_shade.size() + n, -1);
  if (shaded) s->tri_shade.resize(shade0 + n);
  auto fill = [&](uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo; i < hi; ++i) {
      const float* v = vertices + 9 * (size_t)i;
      const size_t id = (size_t)first + i;
      if (shaded) {
        TriShade& ts = s->tri_shade[shade0 + i];
        if (normals) {
          memcpy(ts.n, normals + 9 * (size_t)i, 9 * sizeof(float));
        } else {  // triangular.rs:47-55: un-normalised face normal for all three vertices
          v3 a = mk(v[0], v[1], v[2]), b = mk(v[3], v[4], v[5]), c = mk(v[6], v[7], v[8]);
          v3 fn = cross(b - a, c - a);
          for (int k = 0; k < 3; ++k) { ts.n[3 * k] = fn.x; ts.n[3 * k + 1] = fn.y; ts.n[3 * k + 2] = fn.z; }
        }
        if (uvs) {
          memcpy(ts.uv, uvs + 6 * (size_t)i, 6 * sizeof(float));
        } else {  // triangular.rs:57-65
          const float def[6] = {0.f, 0.f, 1.f, 0.f, 0.f, 1.f};
          memcpy(ts.uv, def, sizeof(def));
        }
        ts.pad = 0.f;
        s->prim_shade[id] = (int32_t)(shade0 + i);
      }
      if (material_ids) s->prim_mat[id] = (uint32_t)material_ids[i];
      s->raw_geom[3 * id] = make_float4(v[0], v[1], v[2], v[3]);
      s->raw_geom[3 * id + 1] = make_float4(v[4], v[5], v[6], v[7]);
      s->raw_geom[3 * id + 2] = make_float4(v[8], 0, 0, 0);
    }
  };
  const uint32_t workers = n >= (1u << 16) ? std::max(1u, std::min(std::thread::hardware_concurrency(), 32u)) : 1u;
  if (workers <= 1) {
    fill(0, n);
  } else {
    std::vector<std::thread> th;
    for (uint32_t w = 0; w < workers; ++w)
      th.emplace_back(fill, (uint32_t)((uint64_t)n * w / workers), (uint32_t)((uint64_t)n * (w + 1) / workers));
    for (auto& t : th) t.join();
  }
  return first;
}

// ---- build / introspection ---------------------------------------------------------------------
int rtw_build(rtw_scene* s, float time0, float time1, rtw_build_stats* stats) {
  CHECK_OPEN(s);
  if (!s->open_kinds.empty() || s->medium_material >= 0) return set_error(RTW_ERR_STATE, "build: unbalanced push/begin");
  if (s->prim_meta.empty()) return set_error(RTW_ERR_INVALID, "build: empty scene");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(RTW_ERR_CUDA, std::string("build: no CUDA device (") + cudaGetErrorString(e) +
                                       "); this backend has no CPU fallback");
  }
  if (s->device >= ndev) return set_error(RTW_ERR_INVALID, "build: device index out of range");
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  // (one attribute, not cudaGetDeviceProperties: that call takes 30-120 ms on this pool's boxes — tools/e2e_probe.py)
  RTW_CUDA_TRY(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, s->device));
  int rc = build_scene_device(s, time0, time1, stats);
  if (rc != RTW_OK) {
    free_scene_device(s);
    return rc;
  }
  s->built = true;
  // the host copies of the big arrays are no longer needed
  raw_vector<float4>().swap(s->raw_geom);
  std::vector<TriShade>().swap(s->tri_shade);
  std::vector<uchar4>().swap(s->texels);
  return RTW_OK;
}

int rtw_scene_num_prims(const rtw_scene* s) {
  if (!s) return set_error(RTW_ERR_INVALID, "scene is NULL");
  return (int)s->prim_meta.size();
}
int rtw_scene_num_nodes(const rtw_scene* s) {
  if (!s) return set_error(RTW_ERR_INVALID, "scene is NULL");
  return s->built ? (int)s->dev.num_nodes : 0;
}
int rtw_scene_num_instances(const rtw_scene* s) {
  if (!s) return set_error(RTW_ERR_INVALID, "scene is NULL");
  return (int)s->inst_range.size();
}
// host-side view of the flattening (no GPU needed): type, instance and material of a primitive,
// and the ops of an instance chain (kind, a, b, c per op; outermost first).
int rtw_scene_prim_info(const rtw_scene* s, int prim_id, int32_t* type, int32_t* instance, int32_t* material) {
  if (!s) return set_error(RTW_ERR_INVALID, "scene is NULL");
  if (prim_id < 0 || prim_id >= (int)s->prim_meta.size()) return set_error(RTW_ERR_INVALID, "prim_info: bad id");
  if (type) *type = (int32_t)(s->prim_meta[prim_id] & 7u);
  if (instance) *instance = (int32_t)(s->prim_meta[prim_id] >> RTW_META_TYPE_BITS);
  if (material) *material = (int32_t)s->prim_mat[prim_id];
  return RTW_OK;
}
int rtw_scene_instance_ops(const rtw_scene* s, int instance, int max_ops, int32_t* kinds, float* abc) {
  if (!s) return set_error(RTW_ERR_INVALID, "scene is NULL");
  if (instance < 0 || instance >= (int)s->inst_range.size()) return set_error(RTW_ERR_INVALID, "instance_ops: bad id");
  uint2 rg = s->inst_range[instance];
  int n = (int)rg.y;
  for (int k = 0; k < n && k < max_ops; ++k) {
    const InstOp& op = s->inst_ops[rg.x + k];
    if (kinds) kinds[k] = (int32_t)op.kind;
    if (abc) { abc[3 * k] = op.a; abc[3 * k + 1] = op.b; abc[3 * k + 2] = op.c; }
  }
  return n;
}

int rtw_get_bvh(const rtw_scene* s, rtw_bvh_node* nodes, int32_t* slot_prim_ids, float* root_box) {
  CHECK_BUILT(s);
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  if (nodes)
    RTW_CUDA_TRY(cudaMemcpy(nodes, s->dev.nodes, (size_t)s->dev.num_nodes * 2 * sizeof(rtw_bvh_node), cudaMemcpyDeviceToHost));
  if (slot_prim_ids)
    RTW_CUDA_TRY(cudaMemcpy(slot_prim_ids, s->dev.slot_prim, (size_t)s->dev.num_prims * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (root_box) memcpy(root_box, s->root_box, 6 * sizeof(float));
  return RTW_OK;
}

// ---- trace -----------------------------------------------------------------------------------------
int rtw_trace_closest_device(rtw_scene* s, const rtw_ray* d_rays, uint64_t n, rtw_hit* d_hits, int mode, void* stream) {
  CHECK_BUILT(s);
  if (n && (!d_rays || !d_hits)) return set_error(RTW_ERR_INVALID, "trace: NULL buffer");
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  return trace_closest_device(s, d_rays, n, d_hits, mode, (cudaStream_t)stream);
}

int rtw_trace_closest(rtw_scene* s, const rtw_ray* rays, uint64_t n, rtw_hit* hits, int mode) {
  CHECK_BUILT(s);
  if (n == 0) return RTW_OK;
  if (!rays || !hits) return set_error(RTW_ERR_INVALID, "trace: NULL buffer");
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  rtw_ray* d_rays = nullptr;
  rtw_hit* d_hits = nullptr;
  cudaError_t e = dev_malloc((void**)&d_rays, n * sizeof(rtw_ray));
  if (e == cudaSuccess) e = dev_malloc((void**)&d_hits, n * sizeof(rtw_hit));
  if (e != cudaSuccess) {
    cudaGetLastError();
    mem_free(d_rays);
    return set_error(RTW_ERR_NOMEM, "trace: cudaMalloc failed");
  }
  int rc = RTW_OK;
  e = cudaMemcpy(d_rays, rays, n * sizeof(rtw_ray), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = trace_closest_device(s, d_rays, n, d_hits, mode, 0);
    if (rc == RTW_OK) e = cudaMemcpy(hits, d_hits, n * sizeof(rtw_hit), cudaMemcpyDeviceToHost);
  }
  mem_free(d_rays);
  mem_free(d_hits);
  if (e != cudaSuccess) return cuda_fail(e, "rtw_trace_closest copy");
  return rc;
}

// ---- render ----------------------------------------------------------------------------------------
int rtw_render_device(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* params, float* d_accum_rgb,
                      void* stream, rtw_render_stats* stats) {
  CHECK_BUILT(s);
  if (!cam || !params || !d_accum_rgb) return set_error(RTW_ERR_INVALID, "render: NULL argument");
  if (params->gpus > 1) return set_error(RTW_ERR_INVALID, "render_device: gpus > 1 is a feature of rtw_render / rtw_render_frames");
  int rc = render_device(s, cam, params, d_accum_rgb, (cudaStream_t)stream, stats);
  if (rc == RTW_OK && stats) stats->gpus = 1;
  return rc;
}

// the frame buffer on the scene's device: kept on the scene, grow only
static int ensure_io(rtw_scene* s, size_t bytes) {
  if (s->io_bytes >= bytes) return RTW_OK;
  if (s->io_frame) mem_free(s->io_frame);
  s->io_frame = nullptr;
  s->io_bytes = 0;
  cudaError_t e = dev_malloc((void**)&s->io_frame, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(RTW_ERR_NOMEM, std::string("render: frame buffer allocation failed: ") + cudaGetErrorString(e));
  }
  s->io_bytes = bytes;
  return RTW_OK;
}

int rtw_render(rtw_scene* s, const rtw_camera* cam, const rtw_render_params* params, float* accum_rgb,
               rtw_render_stats* stats) {
  CHECK_BUILT(s);
  if (!cam || !params || !accum_rgb) return set_error(RTW_ERR_INVALID, "render: NULL argument");
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  const size_t bytes = (size_t)params->width * params->height * 3 * sizeof(float);
  int rc = ensure_io(s, std::max<size_t>(bytes, 4));
  if (rc != RTW_OK) return rc;
  rc = params->gpus > 1 ? render_multi_device(s, cam, params, s->io_frame, stats)
                        : render_device(s, cam, params, s->io_frame, 0, stats);
  if (rc != RTW_OK) return rc;
  if (stats && stats->gpus == 0) stats->gpus = 1;
  // Straight into the caller's buffer.  (A pinned staging buffer of the frame's size costs more to allocate — 40 ms for
  // a 4K frame — than the driver's own staged copy of a pageable destination takes; measured r02.)
  cudaError_t e = cudaMemcpy(accum_rgb, s->io_frame, bytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "rtw_render readback");
  return RTW_OK;
}

// scenes.rs:622-667 + main.rs:48-95: every camera over one resident world; D2H of frame i and its callback
// overlap the rendering of frame i+1.
int rtw_render_frames(rtw_scene* s, const rtw_camera* cameras, uint32_t n_frames, const rtw_render_params* params,
                      rtw_frame_callback on_frame, void* user) {
  CHECK_BUILT(s);
  if (!params || (n_frames && !cameras)) return set_error(RTW_ERR_INVALID, "render_frames: NULL argument");
  if (n_frames == 0) return 0;
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  const size_t bytes = std::max<size_t>((size_t)params->width * params->height * 3 * sizeof(float), 4);
  float* d_accum[2] = {nullptr, nullptr};
  float* h_accum[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr};
  rtw_render_stats stats[2];
  std::future<int> pending[2];  // callback of the frame that last used buffer b
  int rc = RTW_OK, delivered = 0;
  bool stop = false;
  auto cleanup = [&]() {
    for (auto& f : pending)
      if (f.valid()) f.wait();
    for (int b = 0; b < 2; ++b) {
      mem_free(d_accum[b]);
      if (h_accum[b]) mem_free(h_accum[b]);
      if (copied[b]) cudaEventDestroy(copied[b]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
  };
  cudaError_t e = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
    e = dev_malloc((void**)&d_accum[b], bytes);
    if (e == cudaSuccess && on_frame) e = pinned_malloc((void**)&h_accum[b], bytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    cleanup();
    return set_error(RTW_ERR_NOMEM, std::string("render_frames: allocation failed: ") + cudaGetErrorString(e));
  }
  std::shared_future<void> previous;  // callbacks run one at a time, in frame order
  for (uint32_t f = 0; f < n_frames && !stop; ++f) {
    const int b = (int)(f & 1u);
    if (pending[b].valid()) {  // buffer b is free once the callback of frame f-2 has returned
      if (pending[b].get() != 0) stop = true;
      delivered++;
      if (stop) break;
    }
    rtw_render_params p = *params;
    p.seed = params->seed + f;
    rc = p.gpus > 1 ? render_multi_device(s, &cameras[f], &p, d_accum[b], &stats[b])
                    : render_device(s, &cameras[f], &p, d_accum[b], 0, &stats[b]);  // returns when the frame is complete on the device
    if (rc == RTW_OK && stats[b].gpus == 0) stats[b].gpus = 1;
    if (rc != RTW_OK) break;
    if (!on_frame) {
      delivered++;
      continue;
    }
    e = cudaMemcpyAsync(h_accum[b], d_accum[b], bytes, cudaMemcpyDeviceToHost, copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(copied[b], copy_stream);
    if (e != cudaSuccess) {
      rc = cuda_fail(e, "render_frames readback");
      break;
    }
    std::shared_future<void> prev = previous;
    std::promise<void> done;
    previous = done.get_future().share();
    const int device = s->device;
    cudaEvent_t ev = copied[b];
    const float* host = h_accum[b];
    const rtw_render_stats* st = &stats[b];
    pending[b] = std::async(std::launch::async, [=, done = std::move(done)]() mutable {
      cudaSetDevice(device);
      cudaEventSynchronize(ev);
      if (prev.valid()) prev.wait();
      int r = on_frame(user, f, host, st);
      done.set_value();
      return r;
    });
  }
  for (auto& f : pending)
    if (f.valid()) {
      f.get();
      delivered++;
    }
  cleanup();
  return rc != RTW_OK ? rc : delivered;
}

int rtw_resolve_rgb8(rtw_scene* s, const float* accum_rgb, uint32_t width, uint32_t height, uint32_t spp, uint8_t* rgb8) {
  if (!s) return set_error(RTW_ERR_INVALID, "scene is NULL");
  if (!accum_rgb || !rgb8) return set_error(RTW_ERR_INVALID, "resolve: NULL buffer");
  RTW_CUDA_TRY(cudaSetDevice(s->device));
  size_t n = (size_t)width * height * 3;
  float* d_in = nullptr;
  uint8_t* d_out = nullptr;
  cudaError_t e = dev_malloc((void**)&d_in, n * sizeof(float) + 4);
  if (e == cudaSuccess) e = dev_malloc((void**)&d_out, n + 4);
  if (e != cudaSuccess) {
    cudaGetLastError();
    mem_free(d_in);
    return set_error(RTW_ERR_NOMEM, "resolve: cudaMalloc failed");
  }
  int rc = RTW_OK;
  e = cudaMemcpy(d_in, accum_rgb, n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = resolve_rgb8_device(d_in, n, spp, d_out, 0);
    if (rc == RTW_OK) e = cudaMemcpy(rgb8, d_out, n, cudaMemcpyDeviceToHost);
  }
  mem_free(d_in);
  mem_free(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "rtw_resolve_rgb8 copy");
  return rc;
}

}  // extern "C"
