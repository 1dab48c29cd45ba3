// Host front end implementation: sink loading, flatten(), Camera::new, Perlin::new, the OBJ/MTL
// loader and the Raytracer facade.  See rtw_host.hpp.  Compiled with -ffp-contract=off.
#include "rtw_host.hpp"
#include "obj_loader.hpp"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rtwh {

// ---------------------------------------------------------------------------------------------
// sink
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_host_error;

static int host_fail(int code, const std::string& msg) {
  g_host_error = msg;
  return code;
}

int Flattener::check(int rc, const char* what) {
  if (rc < 0) {
    const char* m = sink_->last_error ? sink_->last_error() : "";
    throw Error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (m ? m : ""));
  }
  return rc;
}

}  // namespace rtwh

extern "C" {

const char* rtwh_last_error(void) { return rtwh::g_host_error.c_str(); }

int rtwh_sink_open(const char* path, const char* prefix, int device, rtw_sink* out) {
  if (!path || !prefix || !out) return rtwh::host_fail(RTW_ERR_INVALID, "sink_open: NULL argument");
  memset(out, 0, sizeof(*out));
  void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return rtwh::host_fail(RTW_ERR_INVALID, std::string("sink_open: cannot load backend library: ") + dlerror());
  std::string missing;
  auto sym = [&](const char* name) -> void* {
    std::string full = std::string(prefix) + name;
    void* p = dlsym(lib, full.c_str());
    if (!p) missing += " " + full;
    return p;
  };
#define RTW_BIND(field, name) out->field = reinterpret_cast<decltype(out->field)>(sym(name))
  RTW_BIND(last_error, "last_error");
  RTW_BIND(scene_create, "scene_create");
  RTW_BIND(scene_destroy, "scene_destroy");
  RTW_BIND(add_texture_solid, "add_texture_solid");
  RTW_BIND(add_texture_checker, "add_texture_checker");
  RTW_BIND(add_texture_noise, "add_texture_noise");
  RTW_BIND(add_texture_uvdebug, "add_texture_uvdebug");
  RTW_BIND(add_texture_image, "add_texture_image");
  RTW_BIND(add_material_lambertian, "add_material_lambertian");
  RTW_BIND(add_material_metal, "add_material_metal");
  RTW_BIND(add_material_dielectric, "add_material_dielectric");
  RTW_BIND(add_material_diffuse_light, "add_material_diffuse_light");
  RTW_BIND(push_translation, "push_translation");
  RTW_BIND(push_rotation_y, "push_rotation_y");
  RTW_BIND(pop_transform, "pop_transform");
  RTW_BIND(begin_group, "begin_group");
  RTW_BIND(end_group, "end_group");
  RTW_BIND(begin_medium, "begin_medium");
  RTW_BIND(end_medium, "end_medium");
  RTW_BIND(add_sphere, "add_sphere");
  RTW_BIND(add_moving_sphere, "add_moving_sphere");
  RTW_BIND(add_xy_rect, "add_xy_rect");
  RTW_BIND(add_xz_rect, "add_xz_rect");
  RTW_BIND(add_yz_rect, "add_yz_rect");
  RTW_BIND(add_cuboid, "add_cuboid");
  RTW_BIND(add_triangles, "add_triangles");
  RTW_BIND(build, "build");
  RTW_BIND(render, "render");
  RTW_BIND(render_frames, "render_frames");
#undef RTW_BIND
  if (!missing.empty()) {
    dlclose(lib);
    memset(out, 0, sizeof(*out));
    return rtwh::host_fail(RTW_ERR_INVALID, "sink_open: backend library lacks symbols:" + missing);
  }
  out->lib = lib;
  int rc = out->scene_create(device, &out->scene);
  if (rc < 0) {
    std::string m = out->last_error();
    dlclose(lib);
    memset(out, 0, sizeof(*out));
    return rtwh::host_fail(rc, "sink_open: scene_create failed: " + m);
  }
  return RTW_OK;
}

int rtwh_sink_close(rtw_sink* sink) {
  if (!sink) return RTW_OK;
  if (sink->scene && sink->scene_destroy) sink->scene_destroy(sink->scene);
  if (sink->lib) dlclose(sink->lib);
  memset(sink, 0, sizeof(*sink));
  return RTW_OK;
}

}  // extern "C"

namespace rtwh {

// ---------------------------------------------------------------------------------------------
// HostRng
// ---------------------------------------------------------------------------------------------
static void philox(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3], k0 = k_in[0], k1 = k_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

HostRng::HostRng(uint64_t seed) {
  key_[0] = (uint32_t)seed; key_[1] = (uint32_t)(seed >> 32);
  ctr_[0] = 0; ctr_[1] = 0; ctr_[2] = 0x5CE9Eu; ctr_[3] = 0;
  idx_ = 4;
}
uint32_t HostRng::next_u32() {
  if (idx_ == 4) {
    philox(ctr_, key_, buf_);
    if (++ctr_[0] == 0) ++ctr_[1];
    idx_ = 0;
  }
  return buf_[idx_++];
}
float HostRng::gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
double HostRng::gen_f64() {
  uint64_t lo = next_u32(), hi = next_u32();
  return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
float HostRng::gen_range(float lo, float hi) {
  uint32_t bits = 0x3F800000u | (next_u32() >> 9);
  float v12;
  memcpy(&v12, &bits, 4);
  float scale = hi - lo, offset = lo - scale;
  float res = v12 * scale + offset;
  if (!(res < hi)) res = std::nextafter(hi, lo);
  return res;
}
uint32_t HostRng::gen_below(uint32_t n) { return (uint32_t)(((uint64_t)next_u32() * n) >> 32); }

// ---------------------------------------------------------------------------------------------
// textures
// ---------------------------------------------------------------------------------------------
int Texture::flatten(Flattener& f) const {
  auto it = f.texture_ids.find(this);
  if (it != f.texture_ids.end()) return it->second;
  int id = emit(f);
  f.texture_ids[this] = id;
  return id;
}
int SolidColor::emit(Flattener& f) const {
  return f.check(f.sink()->add_texture_solid(f.scene(), color_.x(), color_.y(), color_.z()), "add_texture_solid");
}
int Checker::emit(Flattener& f) const {
  int o = odd_->flatten(f), e = even_->flatten(f);
  return f.check(f.sink()->add_texture_checker(f.scene(), o, e, frequency_), "add_texture_checker");
}
// perlin.rs:15-48
Perlin::Perlin(HostRng& rng) {
  for (int i = 0; i < 256; ++i) {
    Vec3 g = rng.random_min_max(-1.0f, 1.0f).unit_vector();
    gradients[i][0] = g.x(); gradients[i][1] = g.y(); gradients[i][2] = g.z();
  }
  for (int a = 0; a < 3; ++a) {
    for (int i = 0; i < 256; ++i) permutations[a][i] = i;
    for (int i = 255; i >= 1; --i) {  // [QUIRK] gen_range(0..i) excludes i (perlin.rs:43-48)
      uint32_t target = rng.gen_below((uint32_t)i);
      std::swap(permutations[a][i], permutations[a][target]);
    }
  }
}
int Noise::emit(Flattener& f) const {
  return f.check(f.sink()->add_texture_noise(f.scene(), &perlin_.gradients[0][0], perlin_.permutations[0],
                                             perlin_.permutations[1], perlin_.permutations[2], scale_),
                 "add_texture_noise");
}
int UVDebug::emit(Flattener& f) const { return f.check(f.sink()->add_texture_uvdebug(f.scene()), "add_texture_uvdebug"); }
int ImageTexture::emit(Flattener& f) const {
  return f.check(f.sink()->add_texture_image(f.scene(), rgb_.data(), w_, h_), "add_texture_image");
}

// ---------------------------------------------------------------------------------------------
// materials
// ---------------------------------------------------------------------------------------------
int Material::flatten(Flattener& f) const {
  auto it = f.material_ids.find(this);
  if (it != f.material_ids.end()) return it->second;
  int id = emit(f);
  f.material_ids[this] = id;
  return id;
}
int Lambertian::emit(Flattener& f) const {
  return f.check(f.sink()->add_material_lambertian(f.scene(), albedo_->flatten(f)), "add_material_lambertian");
}
Metal::Metal(Color albedo, float fuzz) : albedo_(albedo), fuzz_(fuzz) {
  if (!(fuzz <= 1.0f)) throw Error("assertion failed: fuzz <= 1.0 (material.rs:71)");
}
int Metal::emit(Flattener& f) const {
  return f.check(f.sink()->add_material_metal(f.scene(), albedo_.x(), albedo_.y(), albedo_.z(), fuzz_), "add_material_metal");
}
int Dielectric::emit(Flattener& f) const { return f.check(f.sink()->add_material_dielectric(f.scene(), ir_), "add_material_dielectric"); }
int DiffuseLight::emit(Flattener& f) const {
  return f.check(f.sink()->add_material_diffuse_light(f.scene(), emit_->flatten(f)), "add_material_diffuse_light");
}

// ---------------------------------------------------------------------------------------------
// hittables
// ---------------------------------------------------------------------------------------------
void Sphere::flatten(Flattener& f) const {
  int m = m_->flatten(f);
  f.check(f.sink()->add_sphere(f.scene(), c_.e, r_, m), "add_sphere");
}
void MovingSphere::flatten(Flattener& f) const {
  int m = m_->flatten(f);
  f.check(f.sink()->add_moving_sphere(f.scene(), c0_.e, t0_, c1_.e, t1_, r_, m), "add_moving_sphere");
}
void XYRectangle::flatten(Flattener& f) const {
  int m = m_->flatten(f);
  f.check(f.sink()->add_xy_rect(f.scene(), a0_, a1_, b0_, b1_, k_, m), "add_xy_rect");
}
void XZRectangle::flatten(Flattener& f) const {
  int m = m_->flatten(f);
  f.check(f.sink()->add_xz_rect(f.scene(), a0_, a1_, b0_, b1_, k_, m), "add_xz_rect");
}
void YZRectangle::flatten(Flattener& f) const {
  int m = m_->flatten(f);
  f.check(f.sink()->add_yz_rect(f.scene(), a0_, a1_, b0_, b1_, k_, m), "add_yz_rect");
}
void Cuboid::flatten(Flattener& f) const {
  int m = m_->flatten(f);
  f.check(f.sink()->add_cuboid(f.scene(), p0_.e, p1_.e, m), "add_cuboid");
}
HittablePtr TriangleMesh::new_flat_shaded(const Point3 v[3], MaterialPtr m) {
  std::vector<float> verts;
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a) verts.push_back(v[i].e[a]);
  return std::make_shared<TriangleMesh>(verts, std::vector<float>(), std::vector<float>(), m);
}
void TriangleMesh::flatten(Flattener& f) const {
  size_t n = len();
  if (n == 0) return;
  if (!per_face_.empty()) {
    if (per_face_.size() != n) throw Error("TriangleMesh: per-face material count mismatch");
    // (consecutive faces mostly share a material: look the id up once per run, not once per face — 10 M map lookups were
    //  half of the flatten time of config C5)
    std::vector<int32_t> ids(n);
    const Material* last = nullptr;
    int32_t last_id = -1;
    for (size_t i = 0; i < n; ++i) {
      const Material* m = per_face_[i].get();
      if (m != last) {
        last = m;
        last_id = m->flatten(f);
      }
      ids[i] = last_id;
    }
    f.check(f.sink()->add_triangles(f.scene(), (uint32_t)n, v_.data(), n_.empty() ? nullptr : n_.data(),
                                    uv_.empty() ? nullptr : uv_.data(), ids.data(), -1),
            "add_triangles");
  } else {
    int m = m_->flatten(f);
    f.check(f.sink()->add_triangles(f.scene(), (uint32_t)n, v_.data(), n_.empty() ? nullptr : n_.data(),
                                    uv_.empty() ? nullptr : uv_.data(), nullptr, m),
            "add_triangles");
  }
}
void BvhNode::flatten(Flattener& f) const {
  f.check(f.sink()->begin_group(f.scene()), "begin_group");
  for (const auto& o : objs_) o->flatten(f);
  f.check(f.sink()->end_group(f.scene()), "end_group");
}
void Translation::flatten(Flattener& f) const {
  f.check(f.sink()->push_translation(f.scene(), off_.e), "push_translation");
  inner_->flatten(f);
  f.check(f.sink()->pop_transform(f.scene()), "pop_transform");
}
void YRotation::flatten(Flattener& f) const {
  f.check(f.sink()->push_rotation_y(f.scene(), deg_), "push_rotation_y");
  inner_->flatten(f);
  f.check(f.sink()->pop_transform(f.scene()), "pop_transform");
}

void ConstantMedium::flatten(Flattener& f) const {
  int t = tex_->flatten(f);
  f.check(f.sink()->begin_medium(f.scene(), density_, t), "begin_medium");
  boundary_->flatten(f);
  f.check(f.sink()->end_medium(f.scene()), "end_medium");
}

void flatten_world(const HittableList& world, rtw_sink* sink, float time0, float time1, rtw_build_stats* stats) {
  Flattener f(sink);
  for (const auto& o : world) o->flatten(f);
  f.check(sink->build(sink->scene, time0, time1, stats), "build");
}

// ---------------------------------------------------------------------------------------------
// camera.rs:25-64
// ---------------------------------------------------------------------------------------------
Camera::Camera(Point3 look_from, Point3 look_at, Vec3 up, float vfov, float aspect_ratio, float aperture, float focus_dist,
               float time0, float time1) {
  const float RADS_PER_DEG = 3.14159274101257324219f / 180.0f;  // f32::to_radians
  float theta = vfov * RADS_PER_DEG;
  float h = std::tan(theta / 2.0f);
  float viewport_height = 2.0f * h;
  float viewport_width = aspect_ratio * viewport_height;
  Vec3 w = (look_from - look_at).unit_vector();
  Vec3 u = up.cross(w).unit_vector();
  Vec3 v = w.cross(u);
  Point3 origin = look_from;
  Vec3 horizontal = focus_dist * viewport_width * u;
  Vec3 vertical = focus_dist * viewport_height * v;
  Point3 llc = origin - horizontal / 2.0f - vertical / 2.0f - focus_dist * w;
  for (int a = 0; a < 3; ++a) {
    c.origin[a] = origin.e[a];
    c.lower_left_corner[a] = llc.e[a];
    c.horizontal[a] = horizontal.e[a];
    c.vertical[a] = vertical.e[a];
    c.u[a] = u.e[a];
    c.v[a] = v.e[a];
    c.w[a] = w.e[a];
  }
  c.lens_radius = aperture / 2.0f;
  c.time0 = time0;
  c.time1 = time1;
}

// ---------------------------------------------------------------------------------------------
// Raytracer (lib.rs:40-76)
// ---------------------------------------------------------------------------------------------
std::vector<Pixel> Raytracer::render(rtw_sink* sink, uint64_t seed, rtw_render_stats* stats) const {
  flatten_world(world_, sink);
  rtw_render_params p;
  memset(&p, 0, sizeof(p));
  p.width = w_;
  p.height = h_;
  p.spp = spp_;
  p.max_depth = 50;  // MAX_DEPTH (lib.rs:32)
  p.background[0] = bg_.x(); p.background[1] = bg_.y(); p.background[2] = bg_.z();
  p.seed = seed;
  std::vector<float> accum((size_t)w_ * h_ * 3);
  int rc = sink->render(sink->scene, &cam_.c, &p, accum.data(), stats);
  if (rc < 0) throw Error(std::string("render failed: ") + sink->last_error());
  return pixels_from_accum(accum.data(), w_, h_);
}

std::vector<Pixel> pixels_from_accum(const float* accum, uint32_t w, uint32_t h) {
  std::vector<Pixel> out((size_t)w * h);
  size_t i = 0;
  for (uint32_t j = h; j-- > 0;)  // (0..h).rev()
    for (uint32_t col = 0; col < w; ++col, ++i) {
      out[i].row = j;
      out[i].column = col;
      out[i].color = Color(accum[3 * i], accum[3 * i + 1], accum[3 * i + 2]);
    }
  return out;
}

// ---------------------------------------------------------------------------------------------
// animation: main.rs:48-95 over a resident scene
// ---------------------------------------------------------------------------------------------
namespace {
struct AnimCtx {
  const FrameFn* fn;
  uint32_t w, h;
  std::string error;
};
int anim_trampoline(void* user, uint32_t frame, const float* accum, const rtw_render_stats* st) {
  AnimCtx* c = (AnimCtx*)user;
  try {
    return (*c->fn)(frame, pixels_from_accum(accum, c->w, c->h), *st) ? 0 : 1;
  } catch (const std::exception& e) {  // never unwind through the C ABI
    c->error = e.what();
    return 1;
  }
}
}  // namespace

uint32_t render_animation(const World& world, uint32_t w, uint32_t h, uint32_t spp, rtw_sink* sink, uint64_t seed,
                          const FrameFn& on_frame, uint32_t gpus) {
  flatten_world(world.objects, sink);
  rtw_render_params p;
  memset(&p, 0, sizeof(p));
  p.width = w;
  p.height = h;
  p.spp = spp;
  p.max_depth = 50;  // MAX_DEPTH (lib.rs:32)
  p.background[0] = world.background.x(); p.background[1] = world.background.y(); p.background[2] = world.background.z();
  p.seed = seed;
  p.gpus = gpus;
  std::vector<rtw_camera> cams;
  for (const auto& c : world.cameras) cams.push_back(c.c);
  AnimCtx ctx{&on_frame, w, h, {}};
  int rc = sink->render_frames(sink->scene, cams.data(), (uint32_t)cams.size(), &p, anim_trampoline, &ctx);
  if (rc < 0) throw Error(std::string("render_frames failed: ") + sink->last_error());
  if (!ctx.error.empty()) throw Error("frame callback failed: " + ctx.error);
  return (uint32_t)rc;
}

// ---------------------------------------------------------------------------------------------
// ProgressMessage wire format: postcard 0.7.3 + COBS (include/rtw_sink.h)
// ---------------------------------------------------------------------------------------------
namespace {
void put_u32(std::vector<uint8_t>& v, uint32_t x) {
  for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i)));
}
void put_f32(std::vector<uint8_t>& v, float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  put_u32(v, x);
}
// Consistent Overhead Byte Stuffing (Cheshire & Baker 1999) + the terminating zero postcard appends
std::vector<uint8_t> cobs_frame(const std::vector<uint8_t>& in) {
  std::vector<uint8_t> out;
  out.reserve(in.size() + in.size() / 254 + 2);
  size_t code_at = 0;
  out.push_back(0);
  uint8_t code = 1;
  for (uint8_t b : in) {
    if (b == 0) {
      out[code_at] = code;
      code_at = out.size();
      out.push_back(0);
      code = 1;
    } else {
      out.push_back(b);
      if (++code == 0xFF) {
        out[code_at] = code;
        code_at = out.size();
        out.push_back(0);
        code = 1;
      }
    }
  }
  out[code_at] = code;
  out.push_back(0);
  return out;
}
}  // namespace

std::vector<uint8_t> to_vec_cobs(const ProgressMessage& m) {
  std::vector<uint8_t> raw;
  raw.push_back((uint8_t)m.kind);  // variant index as a varint: 0, 1, 2 are one byte
  if (m.kind == ProgressMessage::ImageStart) {
    put_u32(raw, m.width); put_u32(raw, m.height); put_u32(raw, m.samples_per_pixel);
  } else if (m.kind == ProgressMessage::PixelMsg) {
    put_u32(raw, m.pixel.row); put_u32(raw, m.pixel.column);
    put_f32(raw, m.pixel.color.x()); put_f32(raw, m.pixel.color.y()); put_f32(raw, m.pixel.color.z());
  }
  return cobs_frame(raw);
}

// ---------------------------------------------------------------------------------------------
// assets
// ---------------------------------------------------------------------------------------------
namespace {
struct ImageAsset { std::vector<uint8_t> rgb; uint32_t w, h; };
struct MeshAsset { std::vector<float> v, n, uv; bool uses_mtl = false; };
std::map<std::string, ImageAsset>& image_registry() { static std::map<std::string, ImageAsset> r; return r; }
std::map<std::string, MeshAsset>& mesh_registry() { static std::map<std::string, MeshAsset> r; return r; }
std::string& asset_dir() { static std::string d = "assets"; return d; }

std::string basename_of(const std::string& p) {
  size_t s = p.find_last_of('/');
  return s == std::string::npos ? p : p.substr(s + 1);
}
std::string stem_of(const std::string& p) {
  std::string b = basename_of(p);
  size_t d = b.find_last_of('.');
  return d == std::string::npos ? b : b.substr(0, d);
}
std::string dir_of(const std::string& p) {
  size_t s = p.find_last_of('/');
  return s == std::string::npos ? std::string(".") : p.substr(0, s);
}
bool file_exists(const std::string& p) { std::ifstream f(p, std::ios::binary); return (bool)f; }

bool read_rtwi(const std::string& path, ImageAsset& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  char magic[4];
  uint32_t hdr[3];
  f.read(magic, 4);
  f.read((char*)hdr, 12);
  if (!f || memcmp(magic, "RTWI", 4) != 0 || hdr[0] != 1) throw Error("bad .rtwi file: " + path);
  out.w = hdr[1]; out.h = hdr[2];
  out.rgb.resize((size_t)out.w * out.h * 3);
  f.read((char*)out.rgb.data(), (std::streamsize)out.rgb.size());
  if (!f) throw Error("truncated .rtwi file: " + path);
  return true;
}
bool read_ppm(const std::string& path, ImageAsset& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::string magic;
  f >> magic;
  if (magic != "P6") return false;
  auto next_int = [&]() {
    for (;;) {
      int c = f.peek();
      if (c == '#') { std::string l; std::getline(f, l); }
      else if (isspace(c)) f.get();
      else break;
    }
    int v; f >> v; return v;
  };
  int w = next_int(), h = next_int(), maxv = next_int();
  f.get();
  if (!f || maxv != 255 || w <= 0 || h <= 0) throw Error("unsupported .ppm file: " + path);
  out.w = (uint32_t)w; out.h = (uint32_t)h;
  out.rgb.resize((size_t)w * h * 3);
  f.read((char*)out.rgb.data(), (std::streamsize)out.rgb.size());
  if (!f) throw Error("truncated .ppm file: " + path);
  return true;
}
bool read_rtwm(const std::string& path, MeshAsset& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  char magic[4];
  uint32_t hdr[3];
  f.read(magic, 4);
  f.read((char*)hdr, 12);
  if (!f || memcmp(magic, "RTWM", 4) != 0 || hdr[0] != 1) throw Error("bad .rtwm file: " + path);
  size_t n = hdr[1];
  out.uses_mtl = (hdr[2] & 4) != 0;
  out.v.resize(n * 9);
  f.read((char*)out.v.data(), (std::streamsize)(n * 9 * 4));
  if (hdr[2] & 1) { out.n.resize(n * 9); f.read((char*)out.n.data(), (std::streamsize)(n * 9 * 4)); }
  if (hdr[2] & 2) { out.uv.resize(n * 6); f.read((char*)out.uv.data(), (std::streamsize)(n * 6 * 4)); }
  if (!f) throw Error("truncated .rtwm file: " + path);
  return true;
}
}  // namespace

void register_image(const std::string& path, std::vector<uint8_t> rgb, uint32_t w, uint32_t h) {
  if (rgb.size() != (size_t)w * h * 3) throw Error("register_image: size mismatch");
  image_registry()[path] = ImageAsset{std::move(rgb), w, h};
}
void register_mesh(const std::string& path, std::vector<float> v, std::vector<float> n, std::vector<float> uv) {
  MeshAsset a;
  a.v = std::move(v); a.n = std::move(n); a.uv = std::move(uv);
  mesh_registry()[path] = std::move(a);
}
void set_asset_dir(const std::string& dir) { asset_dir() = dir; }

bool read_png(const std::string& path, std::vector<uint8_t>& rgb, uint32_t& width, uint32_t& height);   // png_reader.cpp
bool read_jpeg(const std::string& path, std::vector<uint8_t>& rgb, uint32_t& width, uint32_t& height);  // jpeg_reader.cpp

std::shared_ptr<ImageTexture> ImageTexture::open(const std::string& path) {
  auto it = image_registry().find(path);
  if (it != image_registry().end()) return std::make_shared<ImageTexture>(it->second.rgb, it->second.w, it->second.h);
  ImageAsset a;
  if (read_png(path, a.rgb, a.w, a.h)) return std::make_shared<ImageTexture>(std::move(a.rgb), a.w, a.h);  // lossless: exact texels
  // the JPEG file itself (relative to the working directory like the reference, image_texture.rs:24, or under the
  // asset directory): decoded here with libjpeg's default arithmetic (jpeg_reader.cpp)
  if (read_jpeg(path, a.rgb, a.w, a.h) || read_jpeg(asset_dir() + "/" + path, a.rgb, a.w, a.h))
    return std::make_shared<ImageTexture>(std::move(a.rgb), a.w, a.h);
  if (read_rtwi(asset_dir() + "/" + stem_of(path) + ".rtwi", a) || read_rtwi(path, a) ||
      read_ppm(dir_of(path) + "/" + stem_of(path) + ".ppm", a) || read_ppm(path, a))
    return std::make_shared<ImageTexture>(std::move(a.rgb), a.w, a.h);
  throw Error("ImageTexture::open(\"" + path + "\"): file not found. PNG and JPEG files are decoded here (the path itself, then "
              "<assets>/<path>); alternatively register a decoded RGB8 buffer (rtwh_register_image) or provide <assets>/" + stem_of(path) + ".rtwi / a .ppm");
}

// ---------------------------------------------------------------------------------------------
// OBJ / MTL loader (triangular.rs:170-312; parser semantics of wavefront_obj 10.0.0)
// ---------------------------------------------------------------------------------------------
namespace {

int resolve_index(long idx, size_t count, const std::string& path) {
  long r = idx > 0 ? idx - 1 : (long)count + idx;
  if (idx == 0 || r < 0 || r >= (long)count) throw Error("OBJ index out of range in " + path);
  return (int)r;
}

}  // namespace

ObjData parse_obj_simple(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw Error("cannot open OBJ file: " + path);
  std::vector<double> pos, tex, nrm;
  ObjData out;
  std::string line, cur_mtl;
  std::map<std::string, int32_t> mtl_index;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag) || tag[0] == '#') continue;
    if (tag == "v") {
      double x = 0, y = 0, z = 0;
      ss >> x >> y >> z;
      pos.push_back(x); pos.push_back(y); pos.push_back(z);
    } else if (tag == "vt") {
      double u = 0, v = 0;
      ss >> u >> v;
      tex.push_back(u); tex.push_back(v);
    } else if (tag == "vn") {
      double x = 0, y = 0, z = 0;
      ss >> x >> y >> z;
      nrm.push_back(x); nrm.push_back(y); nrm.push_back(z);
    } else if (tag == "mtllib") {
      ss >> out.mtllib;
    } else if (tag == "usemtl") {
      ss >> cur_mtl;
    } else if (tag == "f") {
      struct Corner { int v, t, n; };
      std::vector<Corner> cs;
      std::string tok;
      while (ss >> tok) {
        Corner c{-1, -1, -1};
        size_t s1 = tok.find('/');
        std::string a = tok.substr(0, s1), b, d;
        if (s1 != std::string::npos) {
          size_t s2 = tok.find('/', s1 + 1);
          b = tok.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
          if (s2 != std::string::npos) d = tok.substr(s2 + 1);
        }
        c.v = resolve_index(std::stol(a), pos.size() / 3, path);
        if (!b.empty()) c.t = resolve_index(std::stol(b), tex.size() / 2, path);
        if (!d.empty()) c.n = resolve_index(std::stol(d), nrm.size() / 3, path);
        cs.push_back(c);
      }
      if (cs.size() < 3) throw Error("OBJ points / lines are not supported (triangular.rs:186-191): " + path);
      for (size_t k = 2; k < cs.size(); ++k) {  // triangle fan
        const Corner tri[3] = {cs[0], cs[k - 1], cs[k]};
        bool hn = true, ht = true;
        for (const Corner& c : tri) { hn = hn && c.n >= 0; ht = ht && c.t >= 0; }
        for (const Corner& c : tri) {
          for (int a2 = 0; a2 < 3; ++a2) out.v.push_back((float)pos[3 * c.v + a2]);  // f64 -> `as f32`
          for (int a2 = 0; a2 < 3; ++a2) out.n.push_back(c.n >= 0 ? (float)nrm[3 * c.n + a2] : 0.f);
          for (int a2 = 0; a2 < 2; ++a2) out.uv.push_back(c.t >= 0 ? (float)tex[2 * c.t + a2] : 0.f);
        }
        out.has_n.push_back(hn); out.has_uv.push_back(ht);
        out.all_normals = out.all_normals && hn;
        out.all_uvs = out.all_uvs && ht;
        int32_t mi = -1;
        if (!cur_mtl.empty()) {
          auto it = mtl_index.find(cur_mtl);
          if (it == mtl_index.end()) {
            it = mtl_index.emplace(cur_mtl, (int32_t)out.mtl_names.size()).first;
            out.mtl_names.push_back(cur_mtl);
          }
          mi = it->second;
        }
        out.face_mtl.push_back(mi);
      }
    }
  }
  return out;
}

namespace {

// triangular.rs:278-312
std::map<std::string, MaterialPtr> parse_mtl(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw Error("cannot open MTL file: " + path);
  struct M { int illum = -1; std::string map_kd; };
  std::vector<std::pair<std::string, M>> mats;
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag) || tag[0] == '#') continue;
    if (tag == "newmtl") { std::string n; ss >> n; mats.push_back({n, M()}); }
    else if (mats.empty()) continue;
    else if (tag == "illum") ss >> mats.back().second.illum;
    else if (tag == "map_Kd") ss >> mats.back().second.map_kd;
  }
  std::map<std::string, MaterialPtr> out;
  for (auto& m : mats) {
    if (m.second.illum != 1) throw Error("MTL material '" + m.first + "': only illum 1 is supported (triangular.rs:300-302)");
    if (m.second.map_kd.empty()) throw Error("MTL material '" + m.first + "': map_Kd is required (triangular.rs:304-309)");
    out[m.first] = std::make_shared<Lambertian>(ImageTexture::open(dir_of(path) + "/" + m.second.map_kd));
  }
  return out;
}

// For faces that lack some vertex normals / uvs the reference fills them per vertex
// (triangular.rs:47-65); the C ABI takes whole arrays, so fill here with the same values.
void fill_defaults(ObjData& d) {
  size_t n = d.has_n.size();
  for (size_t i = 0; i < n; ++i) {
    if (!d.has_n[i] && !d.all_normals) {
      const float* v = &d.v[9 * i];
      Vec3 a(v[0], v[1], v[2]), b(v[3], v[4], v[5]), c(v[6], v[7], v[8]);
      Vec3 fn = (b - a).cross(c - a);
      for (int k = 0; k < 3; ++k)
        for (int a2 = 0; a2 < 3; ++a2) d.n[9 * i + 3 * k + a2] = fn.e[a2];
    }
    if (!d.has_uv[i] && !d.all_uvs) {
      const float def[6] = {0.f, 0.f, 1.f, 0.f, 0.f, 1.f};
      memcpy(&d.uv[6 * i], def, sizeof(def));
    }
  }
}

}  // namespace

std::shared_ptr<TriangleMesh> load_mesh(const std::string& path, MaterialPtr material) {
  auto it = mesh_registry().find(path);
  if (it != mesh_registry().end()) return std::make_shared<TriangleMesh>(it->second.v, it->second.n, it->second.uv, material);
  if (file_exists(path) && path.size() > 4 && path.substr(path.size() - 4) == ".obj") {
    ObjData d = parse_obj_fast(path);
    bool any_n = false, any_uv = false;
    for (auto h : d.has_n) any_n = any_n || h;
    for (auto h : d.has_uv) any_uv = any_uv || h;
    fill_defaults(d);
    return std::make_shared<TriangleMesh>(std::move(d.v), any_n ? std::move(d.n) : std::vector<float>(),
                                          any_uv ? std::move(d.uv) : std::vector<float>(), material);
  }
  MeshAsset a;
  if (read_rtwm(asset_dir() + "/" + stem_of(path) + ".rtwm", a) || read_rtwm(path, a))
    return std::make_shared<TriangleMesh>(std::move(a.v), std::move(a.n), std::move(a.uv), material);
  throw Error("load_mesh(\"" + path + "\"): neither the OBJ file nor <assets>/" + stem_of(path) + ".rtwm exists");
}

HittablePtr load_wavefront_obj(const std::string& path, MaterialPtr override_material) {
  HittableList tris;
  if (override_material) {
    tris.push_back(load_mesh(path, override_material));
  } else if (file_exists(path)) {
    ObjData d = parse_obj_fast(path);
    std::map<std::string, MaterialPtr> lib;
    bool have_lib = false;
    if (!d.mtllib.empty()) { lib = parse_mtl(dir_of(path) + "/" + d.mtllib); have_lib = true; }
    MaterialPtr magenta = std::make_shared<DiffuseLight>(SolidColor::new_rgb(1.0f, 0.0f, 1.0f));  // triangular.rs:181
    std::vector<MaterialPtr> by_index;
    for (const auto& name : d.mtl_names) {
      if (!have_lib) throw Error("OBJ uses usemtl without mtllib (triangular.rs:177-179 unwraps None): " + path);
      auto m = lib.find(name);
      if (m == lib.end()) throw Error("OBJ material not found in MTL: " + name);
      by_index.push_back(m->second);
    }
    std::vector<MaterialPtr> per_face;
    per_face.reserve(d.face_mtl.size());
    for (int32_t mi : d.face_mtl) per_face.push_back(mi < 0 ? magenta : by_index[(size_t)mi]);
    bool any_n = false, any_uv = false;
    for (auto h : d.has_n) any_n = any_n || h;
    for (auto h : d.has_uv) any_uv = any_uv || h;
    fill_defaults(d);
    tris.push_back(std::make_shared<TriangleMesh>(d.v, any_n ? d.n : std::vector<float>(), any_uv ? d.uv : std::vector<float>(), per_face));
  } else {
    // binary fixture: carries geometry only.  Without usemtl the reference gives every face the magenta
    // emitter (triangular.rs:177-182); with usemtl it needs the MTL + its map_Kd image, which a fixture
    // cannot provide (for the monument the PNG is missing from the reference tree itself).
    MeshAsset probe;
    if ((read_rtwm(asset_dir() + "/" + stem_of(path) + ".rtwm", probe) || read_rtwm(path, probe)) && probe.uses_mtl)
      throw Error("load_wavefront_obj(\"" + path + "\"): the OBJ names materials (usemtl) but neither the OBJ/MTL files nor "
                  "their map_Kd image are available: no decoded image for the material library");
    tris.push_back(load_mesh(path, std::make_shared<DiffuseLight>(SolidColor::new_rgb(1.0f, 0.0f, 1.0f))));
  }
  return std::make_shared<BvhNode>(tris, 0.0f, 1.0f);  // triangular.rs:259
}

}  // namespace rtwh
