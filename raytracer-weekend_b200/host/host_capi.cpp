// C entry points of the host front end (librtw_host.so) for harnesses that are not C++
// (the pytest suite and bench.py drive it through ctypes).  Exceptions become error codes +
// rtwh_last_error().
#include <chrono>
#include <cstring>

#include "obj_loader.hpp"
#include "rtw_host.hpp"

namespace rtwh {
int host_fail_public(int code, const std::string& msg);
}

static thread_local std::string g_capi_error;
extern "C" const char* rtwh_last_error(void);

namespace {
int fail(int code, const std::string& m) {
  g_capi_error = m;
  return code;
}
}  // namespace

extern "C" {

const char* rtwh_capi_error(void) { return g_capi_error.c_str(); }

int rtwh_set_asset_dir(const char* dir) {
  if (!dir) return fail(RTW_ERR_INVALID, "dir is NULL");
  rtwh::set_asset_dir(dir);
  return RTW_OK;
}

int rtwh_register_image(const char* path, const uint8_t* rgb8, uint32_t w, uint32_t h) {
  if (!path || !rgb8) return fail(RTW_ERR_INVALID, "register_image: NULL argument");
  try {
    rtwh::register_image(path, std::vector<uint8_t>(rgb8, rgb8 + (size_t)w * h * 3), w, h);
  } catch (const std::exception& e) {
    return fail(RTW_ERR_INVALID, e.what());
  }
  return RTW_OK;
}

int rtwh_register_mesh(const char* path, uint32_t ntris, const float* verts, const float* normals, const float* uvs) {
  if (!path || !verts) return fail(RTW_ERR_INVALID, "register_mesh: NULL argument");
  std::vector<float> v(verts, verts + (size_t)ntris * 9), n, uv;
  if (normals) n.assign(normals, normals + (size_t)ntris * 9);
  if (uvs) uv.assign(uvs, uvs + (size_t)ntris * 6);
  rtwh::register_mesh(path, std::move(v), std::move(n), std::move(uv));
  return RTW_OK;
}

// names separated by '\n'; returns the number of scenes
int rtwh_scene_names(char* buf, size_t buf_len) {
  std::string all;
  auto names = rtwh::scene_names();
  for (size_t i = 0; i < names.size(); ++i) all += (i ? "\n" : "") + names[i];
  if (buf && buf_len) {
    strncpy(buf, all.c_str(), buf_len - 1);
    buf[buf_len - 1] = 0;
  }
  return (int)names.size();
}

// Scene::generate (scenes.rs:42-60) + flatten into `sink` + build.  cams: up to max_cams cameras
// are written; returns the number of cameras of the scene (>= 1) or an error code.
int rtwh_build_scene(const char* name, float aspect_ratio, uint64_t seed, rtw_sink* sink, rtw_camera* cams, int max_cams,
                     float background[3], rtw_build_stats* stats) {
  if (!name || !sink) return fail(RTW_ERR_INVALID, "build_scene: NULL argument");
  try {
    rtwh::World w = rtwh::generate_scene(name, aspect_ratio, seed);
    rtwh::flatten_world(w.objects, sink, 0.0f, 1.0f, stats);
    for (int i = 0; i < (int)w.cameras.size() && i < max_cams; ++i) cams[i] = w.cameras[i].c;
    if (background) {
      background[0] = w.background.x(); background[1] = w.background.y(); background[2] = w.background.z();
    }
    return (int)w.cameras.size();
  } catch (const rtwh::Error& e) {
    return fail(RTW_ERR_INVALID, e.what());
  } catch (const std::exception& e) {
    return fail(RTW_ERR_INVALID, std::string("build_scene: ") + e.what());
  }
}

// The two halves of rtwh_build_scene, separately: generate a world once (Scene::generate), flatten it into any number
// of sinks.  A world handle is a heap-allocated rtwh::World.
void* rtwh_world_create(const char* name, float aspect_ratio, uint64_t seed) {
  if (!name) { fail(RTW_ERR_INVALID, "world_create: name is NULL"); return nullptr; }
  try {
    return new rtwh::World(rtwh::generate_scene(name, aspect_ratio, seed));
  } catch (const std::exception& e) {
    fail(RTW_ERR_INVALID, e.what());
    return nullptr;
  }
}
void rtwh_world_destroy(void* world) { delete (rtwh::World*)world; }
// cameras: up to max_cams are written; returns the number of cameras of the world
int rtwh_world_info(const void* world, rtw_camera* cams, int max_cams, float background[3]) {
  if (!world) return fail(RTW_ERR_INVALID, "world_info: world is NULL");
  const rtwh::World& w = *(const rtwh::World*)world;
  for (int i = 0; cams && i < (int)w.cameras.size() && i < max_cams; ++i) cams[i] = w.cameras[i].c;
  if (background) {
    background[0] = w.background.x(); background[1] = w.background.y(); background[2] = w.background.z();
  }
  return (int)w.cameras.size();
}
// flatten (emit calls in canonical order) + rtw_build on the sink's backend
int rtwh_world_flatten(const void* world, rtw_sink* sink, rtw_build_stats* stats) {
  if (!world || !sink) return fail(RTW_ERR_INVALID, "world_flatten: NULL argument");
  try {
    rtwh::flatten_world(((const rtwh::World*)world)->objects, sink, 0.0f, 1.0f, stats);
    return RTW_OK;
  } catch (const std::exception& e) {
    return fail(RTW_ERR_INVALID, e.what());
  }
}

// Camera::new (camera.rs:25-64)
int rtwh_camera_new(const float look_from[3], const float look_at[3], const float up[3], float vfov, float aspect,
                    float aperture, float focus_dist, float time0, float time1, rtw_camera* out) {
  if (!look_from || !look_at || !up || !out) return fail(RTW_ERR_INVALID, "camera_new: NULL argument");
  rtwh::Camera c(rtwh::Point3(look_from[0], look_from[1], look_from[2]), rtwh::Point3(look_at[0], look_at[1], look_at[2]),
                 rtwh::Vec3(up[0], up[1], up[2]), vfov, aspect, aperture, focus_dist, time0, time1);
  *out = c.c;
  return RTW_OK;
}

// Perlin::new (perlin.rs:15-29) with the seeded host stream
int rtwh_perlin_new(uint64_t seed, float* gradients_256x3, int32_t* px, int32_t* py, int32_t* pz) {
  rtwh::HostRng rng(seed);
  rtwh::Perlin p(rng);
  memcpy(gradients_256x3, p.gradients, sizeof(p.gradients));
  memcpy(px, p.permutations[0], 256 * 4);
  memcpy(py, p.permutations[1], 256 * 4);
  memcpy(pz, p.permutations[2], 256 * 4);
  return RTW_OK;
}

// load_wavefront_obj's mesh (triangular.rs:170-260): ntris, and optionally the arrays.
// Call once with NULL arrays to size them.  has_normals / has_uvs report what the file carries.
int rtwh_load_obj(const char* path, uint32_t* ntris, float* verts, float* normals, float* uvs, int* has_normals, int* has_uvs) {
  if (!path || !ntris) return fail(RTW_ERR_INVALID, "load_obj: NULL argument");
  try {
    struct Capture {
      std::vector<float> v, n, uv;
    } cap;
    // flatten the mesh into a capturing sink
    static thread_local Capture* tl = nullptr;
    tl = &cap;
    rtw_sink sink;
    memset(&sink, 0, sizeof(sink));
    sink.last_error = []() -> const char* { return ""; };
    sink.add_material_lambertian = [](void*, int) { return 0; };
    sink.add_texture_solid = [](void*, float, float, float) { return 0; };
    sink.add_triangles = [](void*, uint32_t n, const float* v, const float* nr, const float* uv, const int32_t*, int) {
      tl->v.insert(tl->v.end(), v, v + (size_t)n * 9);
      if (nr) tl->n.insert(tl->n.end(), nr, nr + (size_t)n * 9);
      if (uv) tl->uv.insert(tl->uv.end(), uv, uv + (size_t)n * 6);
      return 0;
    };
    auto mesh = rtwh::load_mesh(path, rtwh::Lambertian::new_solid_color(rtwh::Color(0.5f, 0.5f, 0.5f)));
    rtwh::Flattener f(&sink);
    mesh->flatten(f);
    *ntris = (uint32_t)(cap.v.size() / 9);
    if (has_normals) *has_normals = cap.n.empty() ? 0 : 1;
    if (has_uvs) *has_uvs = cap.uv.empty() ? 0 : 1;
    if (verts) memcpy(verts, cap.v.data(), cap.v.size() * 4);
    if (normals && !cap.n.empty()) memcpy(normals, cap.n.data(), cap.n.size() * 4);
    if (uvs && !cap.uv.empty()) memcpy(uvs, cap.uv.data(), cap.uv.size() * 4);
    return RTW_OK;
  } catch (const std::exception& e) {
    return fail(RTW_ERR_INVALID, e.what());
  }
}

// ImageTexture::open (image_texture.rs:23-30) as the scenes use it: registry, PNG decoder, .rtwi / .ppm.  Call with
// rgb8 == NULL to size the buffer (width * height * 3 bytes, row 0 = top).
int rtwh_open_image(const char* path, uint32_t* width, uint32_t* height, uint8_t* rgb8, size_t cap) {
  if (!path || !width || !height) return fail(RTW_ERR_INVALID, "open_image: NULL argument");
  try {
    auto t = rtwh::ImageTexture::open(path);
    *width = t->width(); *height = t->height();
    if (rgb8) {
      if (cap < t->rgb().size()) return fail(RTW_ERR_INVALID, "open_image: buffer too small");
      memcpy(rgb8, t->rgb().data(), t->rgb().size());
    }
    return RTW_OK;
  } catch (const std::exception& e) {
    return fail(RTW_ERR_INVALID, e.what());
  }
}

// OBJ ingest check / benchmark: parse `path` with the parallel parser (mode 0, `threads` <= 0: all) or with the
// single-threaded reference parser (mode 1).  Returns the triangle count; checksum = FNV-1a over every output array
// (positions, normals, uvs, presence flags, per-face material names), so two parses can be compared without moving
// the arrays across the boundary; seconds = wall time of the parse.
long long rtwh_parse_obj(const char* path, int mode, int threads, uint64_t* checksum, double* seconds) {
  if (!path) return fail(RTW_ERR_INVALID, "parse_obj: path is NULL");
  try {
    auto t0 = std::chrono::steady_clock::now();
    rtwh::ObjData d = mode == 0 ? rtwh::parse_obj_fast(path, threads) : rtwh::parse_obj_simple(path);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (checksum) {
      uint64_t h = 1469598103934665603ull;
      auto mix = [&](const void* p, size_t n) {
        const uint8_t* b = (const uint8_t*)p;
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
      };
      mix(d.v.data(), d.v.size() * 4); mix(d.n.data(), d.n.size() * 4); mix(d.uv.data(), d.uv.size() * 4);
      mix(d.has_n.data(), d.has_n.size()); mix(d.has_uv.data(), d.has_uv.size());
      for (int32_t mi : d.face_mtl) {
        const std::string name = mi < 0 ? std::string() : d.mtl_names[(size_t)mi];
        mix(name.data(), name.size());
        mix("|", 1);
      }
      mix(d.mtllib.data(), d.mtllib.size());
      const uint8_t flags[2] = {(uint8_t)d.all_normals, (uint8_t)d.all_uvs};
      mix(flags, 2);
      *checksum = h;
    }
    return (long long)d.triangles();
  } catch (const std::exception& e) {
    return fail(RTW_ERR_INVALID, e.what());
  }
}

// ---- ProgressMessage wire format (include/rtw_sink.h) ----------------------------------------------------
static int emit_message(const rtwh::ProgressMessage& m, uint8_t* out, size_t cap) {
  std::vector<uint8_t> b = rtwh::to_vec_cobs(m);
  if (!out || cap < b.size()) return fail(RTW_ERR_INVALID, "progress: output buffer too small");
  memcpy(out, b.data(), b.size());
  return (int)b.size();
}
int rtwh_progress_image_start(uint32_t width, uint32_t height, uint32_t spp, uint8_t* out, size_t cap) {
  rtwh::ProgressMessage m;
  m.kind = rtwh::ProgressMessage::ImageStart;
  m.width = width; m.height = height; m.samples_per_pixel = spp;
  return emit_message(m, out, cap);
}
int rtwh_progress_pixel(uint32_t row, uint32_t column, const float color[3], uint8_t* out, size_t cap) {
  if (!color) return fail(RTW_ERR_INVALID, "progress: color is NULL");
  rtwh::ProgressMessage m;
  m.kind = rtwh::ProgressMessage::PixelMsg;
  m.pixel.row = row; m.pixel.column = column;
  m.pixel.color = rtwh::Color(color[0], color[1], color[2]);
  return emit_message(m, out, cap);
}
int rtwh_progress_image_end(uint8_t* out, size_t cap) {
  rtwh::ProgressMessage m;
  m.kind = rtwh::ProgressMessage::ImageEnd;
  return emit_message(m, out, cap);
}
size_t rtwh_progress_frame_bound(uint32_t width, uint32_t height) {
  return 32 + (size_t)width * height * 24 + 8;  // COBS adds 1 byte per message below 254 bytes, postcard 1 terminator
}
long long rtwh_progress_frame(const float* accum_rgb, uint32_t width, uint32_t height, uint32_t spp, uint8_t* out, size_t cap) {
  if (!accum_rgb || !out) return fail(RTW_ERR_INVALID, "progress: NULL buffer");
  size_t n = 0;
  auto put = [&](const rtwh::ProgressMessage& m) {
    std::vector<uint8_t> b = rtwh::to_vec_cobs(m);
    if (n + b.size() > cap) return false;
    memcpy(out + n, b.data(), b.size());
    n += b.size();
    return true;
  };
  rtwh::ProgressMessage m;
  m.kind = rtwh::ProgressMessage::ImageStart;
  m.width = width; m.height = height; m.samples_per_pixel = spp;
  if (!put(m)) return fail(RTW_ERR_INVALID, "progress: output buffer too small");
  m.kind = rtwh::ProgressMessage::PixelMsg;
  for (const rtwh::Pixel& px : rtwh::pixels_from_accum(accum_rgb, width, height)) {
    m.pixel = px;
    if (!put(m)) return fail(RTW_ERR_INVALID, "progress: output buffer too small");
  }
  m.kind = rtwh::ProgressMessage::ImageEnd;
  if (!put(m)) return fail(RTW_ERR_INVALID, "progress: output buffer too small");
  return (long long)n;
}

}  // extern "C"
