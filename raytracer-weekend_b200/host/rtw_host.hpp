// Host front end: a C++ mirror of raytracer_weekend_lib's public scene API
// (Hittable / Material / Texture / Camera / Raytracer) whose objects do not intersect anything
// themselves — each one knows how to FLATTEN itself into a rtw_sink (include/rtw_sink.h), i.e.
// into the C ABI of the CUDA backend.  This is the C++ stand-in for the
// `fn flatten(&self, &mut dyn SceneSink)` hook SURVEY.md §8b proposes for the Rust traits
// (no Rust toolchain exists in this image; INTEGRATION.md shows the Rust side).
//
// Names, constructor argument order and error behaviour follow the reference:
//   texture.rs, material.rs, light_source.rs, hittable/{spherical,rectangular,triangular,
//   transformations}.rs, bvh.rs, camera.rs, lib.rs (Raytracer, Pixel).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <unordered_map>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtw_sink.h"

namespace rtwh {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- vec3.rs (host-side subset; compiled with -ffp-contract=off) ---------------------------------
struct Vec3 {
  float e[3];
  Vec3() : e{0.f, 0.f, 0.f} {}
  Vec3(float a, float b, float c) : e{a, b, c} {}
  float x() const { return e[0]; }
  float y() const { return e[1]; }
  float z() const { return e[2]; }
  float length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
  float length() const { return std::sqrt(length_squared()); }
  Vec3 cross(const Vec3& r) const {
    return Vec3(e[1] * r.e[2] - e[2] * r.e[1], e[2] * r.e[0] - e[0] * r.e[2], e[0] * r.e[1] - e[1] * r.e[0]);
  }
  Vec3 unit_vector() const { float l = length(); return Vec3(e[0] / l, e[1] / l, e[2] / l); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
inline Vec3 operator*(float s, Vec3 a) { return Vec3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }
inline Vec3 operator/(Vec3 a, float s) { return Vec3(a.e[0] / s, a.e[1] / s, a.e[2] / s); }
typedef Vec3 Point3;
typedef Vec3 Color;

// ---- host random stream (scene generation only; the reference uses thread_rng() there) ----------
// Philox4x32-10 counter stream keyed by a 64-bit seed; float conversions as rand 0.9 does them.
class HostRng {
 public:
  explicit HostRng(uint64_t seed);
  uint32_t next_u32();
  float gen_f32();                         // gen::<f32>()
  double gen_f64();                        // gen::<f64>()
  float gen_range(float lo, float hi);     // gen_range(lo..hi)
  uint32_t gen_below(uint32_t n);          // gen_range(0..n)
  Vec3 random_vec() { float a = gen_range(0.f, 1.f), b = gen_range(0.f, 1.f), c = gen_range(0.f, 1.f); return Vec3(a, b, c); }
  Vec3 random_min_max(float lo, float hi) { float a = gen_range(lo, hi), b = gen_range(lo, hi), c = gen_range(lo, hi); return Vec3(a, b, c); }

 private:
  uint32_t key_[2], ctr_[4], buf_[4];
  int idx_;
};

// ---- flatten context ---------------------------------------------------------------------------
class Flattener {
 public:
  explicit Flattener(rtw_sink* sink) : sink_(sink) {}
  rtw_sink* sink() { return sink_; }
  void* scene() { return sink_->scene; }
  int check(int rc, const char* what);  // throws Error with the sink's message on rc < 0
  std::unordered_map<const void*, int> texture_ids, material_ids;  // object -> emitted id (every object is emitted once)

 private:
  rtw_sink* sink_;
};

// ---- texture.rs / perlin.rs / image_texture.rs ------------------------------------------------
class Texture {
 public:
  virtual ~Texture() {}
  int flatten(Flattener& f) const;  // memoised per Flattener

 protected:
  virtual int emit(Flattener& f) const = 0;
};
typedef std::shared_ptr<const Texture> TexturePtr;

class SolidColor : public Texture {  // texture.rs:45-60
 public:
  explicit SolidColor(Color c) : color_(c) {}
  static TexturePtr new_rgb(float r, float g, float b) { return std::make_shared<SolidColor>(Color(r, g, b)); }

 protected:
  int emit(Flattener& f) const override;
  Color color_;
};

class Checker : public Texture {  // texture.rs:62-81 — Checker::new(odd, even, frequency)
 public:
  Checker(TexturePtr odd, TexturePtr even, float frequency) : odd_(odd), even_(even), frequency_(frequency) {}

 protected:
  int emit(Flattener& f) const override;
  TexturePtr odd_, even_;
  float frequency_;
};

struct Perlin {  // perlin.rs:9-48
  float gradients[256][3];
  int32_t permutations[3][256];
  explicit Perlin(HostRng& rng);
};

class Noise : public Texture {  // texture.rs:83-95
 public:
  Noise(const Perlin& p, float scale) : perlin_(p), scale_(scale) {}

 protected:
  int emit(Flattener& f) const override;
  Perlin perlin_;
  float scale_;
};

class UVDebug : public Texture {  // texture.rs:97-104
 protected:
  int emit(Flattener& f) const override;
};

class ImageTexture : public Texture {  // image_texture.rs:17-51
 public:
  // decoded RGB8, row 0 = top
  ImageTexture(std::vector<uint8_t> rgb, uint32_t w, uint32_t h) : rgb_(std::move(rgb)), w_(w), h_(h) {}
  // ImageTexture::open: resolves `path` through the asset registry, then decodes a PNG file (lossless: the texels
  // are exactly the reference decoder's), then .rtwi / .ppm files with pre-decoded pixels
  static std::shared_ptr<ImageTexture> open(const std::string& path);
  uint32_t width() const { return w_; }
  uint32_t height() const { return h_; }
  const std::vector<uint8_t>& rgb() const { return rgb_; }

 protected:
  int emit(Flattener& f) const override;
  std::vector<uint8_t> rgb_;
  uint32_t w_, h_;
};

// ---- material.rs / light_source.rs ---------------------------------------------------------------
class Material {
 public:
  virtual ~Material() {}
  int flatten(Flattener& f) const;

 protected:
  virtual int emit(Flattener& f) const = 0;
};
typedef std::shared_ptr<const Material> MaterialPtr;

class Lambertian : public Material {  // material.rs:30-61
 public:
  explicit Lambertian(TexturePtr albedo) : albedo_(albedo) {}
  static MaterialPtr new_solid_color(Color c) { return std::make_shared<Lambertian>(std::make_shared<SolidColor>(c)); }

 protected:
  int emit(Flattener& f) const override;
  TexturePtr albedo_;
};
class Metal : public Material {  // material.rs:63-100 ; assert!(fuzz <= 1.0) -> throws
 public:
  Metal(Color albedo, float fuzz);

 protected:
  int emit(Flattener& f) const override;
  Color albedo_;
  float fuzz_;
};
class Dielectric : public Material {  // material.rs:102-147
 public:
  explicit Dielectric(float ir) : ir_(ir) {}

 protected:
  int emit(Flattener& f) const override;
  float ir_;
};
class DiffuseLight : public Material {  // light_source.rs:13-24
 public:
  explicit DiffuseLight(TexturePtr emit_tex) : emit_(emit_tex) {}

 protected:
  int emit(Flattener& f) const override;
  TexturePtr emit_;
};

// ---- hittable/*.rs ---------------------------------------------------------------------------------
class Hittable {
 public:
  virtual ~Hittable() {}
  // emits the object's primitives in the reference's traversal order -> canonical primitive ids
  virtual void flatten(Flattener& f) const = 0;
};
typedef std::shared_ptr<const Hittable> HittablePtr;
typedef std::vector<HittablePtr> HittableList;  // `Vec<Box<dyn Hittable>>`

class Sphere : public Hittable {  // spherical.rs:80-105
 public:
  Sphere(Point3 center, float radius, MaterialPtr m) : c_(center), r_(radius), m_(m) {}
  void flatten(Flattener& f) const override;

 private:
  Point3 c_;
  float r_;
  MaterialPtr m_;
};
class MovingSphere : public Hittable {  // spherical.rs:107-151 — new(center0, time0, center1, time1, radius, material)
 public:
  MovingSphere(Point3 c0, float t0, Point3 c1, float t1, float radius, MaterialPtr m)
      : c0_(c0), c1_(c1), t0_(t0), t1_(t1), r_(radius), m_(m) {}
  void flatten(Flattener& f) const override;

 private:
  Point3 c0_, c1_;
  float t0_, t1_, r_;
  MaterialPtr m_;
};
class XYRectangle : public Hittable {  // rectangular.rs:16-65 — new(x0, x1, y0, y1, k, material)
 public:
  XYRectangle(float x0, float x1, float y0, float y1, float k, MaterialPtr m) : a0_(x0), a1_(x1), b0_(y0), b1_(y1), k_(k), m_(m) {}
  void flatten(Flattener& f) const override;

 private:
  float a0_, a1_, b0_, b1_, k_;
  MaterialPtr m_;
};
class XZRectangle : public Hittable {  // rectangular.rs:67-116
 public:
  XZRectangle(float x0, float x1, float z0, float z1, float k, MaterialPtr m) : a0_(x0), a1_(x1), b0_(z0), b1_(z1), k_(k), m_(m) {}
  void flatten(Flattener& f) const override;

 private:
  float a0_, a1_, b0_, b1_, k_;
  MaterialPtr m_;
};
class YZRectangle : public Hittable {  // rectangular.rs:118-167
 public:
  YZRectangle(float y0, float y1, float z0, float z1, float k, MaterialPtr m) : a0_(y0), a1_(y1), b0_(z0), b1_(z1), k_(k), m_(m) {}
  void flatten(Flattener& f) const override;

 private:
  float a0_, a1_, b0_, b1_, k_;
  MaterialPtr m_;
};
class Cuboid : public Hittable {  // rectangular.rs:170-245
 public:
  Cuboid(Point3 p0, Point3 p1, MaterialPtr m) : p0_(p0), p1_(p1), m_(m) {}
  void flatten(Flattener& f) const override;

 private:
  Point3 p0_, p1_;
  MaterialPtr m_;
};

// A batch of triangles sharing the flatten call (one reference `Triangle` each, triangular.rs:34-73).
class TriangleMesh : public Hittable {
 public:
  // verts: n*9; normals: n*9 or empty; uvs: n*6 or empty; one material for all faces or one per face
  TriangleMesh(std::vector<float> verts, std::vector<float> normals, std::vector<float> uvs, MaterialPtr m)
      : v_(std::move(verts)), n_(std::move(normals)), uv_(std::move(uvs)), m_(m) {}
  TriangleMesh(std::vector<float> verts, std::vector<float> normals, std::vector<float> uvs, std::vector<MaterialPtr> per_face)
      : v_(std::move(verts)), n_(std::move(normals)), uv_(std::move(uvs)), per_face_(std::move(per_face)) {}
  // Triangle::new_flat_shaded (triangular.rs:75-77)
  static HittablePtr new_flat_shaded(const Point3 v[3], MaterialPtr m);
  size_t len() const { return v_.size() / 9; }
  // same geometry, one material per face
  std::shared_ptr<TriangleMesh> with_materials(std::vector<MaterialPtr> per_face) const {
    return std::make_shared<TriangleMesh>(v_, n_, uv_, std::move(per_face));
  }
  void flatten(Flattener& f) const override;

 private:
  std::vector<float> v_, n_, uv_;
  MaterialPtr m_;
  std::vector<MaterialPtr> per_face_;
};

class BvhNode : public Hittable {  // bvh.rs:12-74: an acceleration hint; flattens to begin/end_group
 public:
  BvhNode(HittableList objects, float time0, float time1) : objs_(std::move(objects)), t0_(time0), t1_(time1) {
    if (objs_.empty()) throw Error("BvhNode::new: empty object list");
  }
  void flatten(Flattener& f) const override;

 private:
  HittableList objs_;
  float t0_, t1_;
};

class Translation : public Hittable {  // transformations.rs:16-48
 public:
  Translation(HittablePtr inner, Vec3 offset) : inner_(inner), off_(offset) {}
  void flatten(Flattener& f) const override;

 private:
  HittablePtr inner_;
  Vec3 off_;
};
class YRotation : public Hittable {  // transformations.rs:50-153
 public:
  YRotation(HittablePtr inner, float angle_degrees) : inner_(inner), deg_(angle_degrees) {}
  void flatten(Flattener& f) const override;

 private:
  HittablePtr inner_;
  float deg_;
};
class ConstantMedium : public Hittable {  // volumes.rs:18-35 — ConstantMedium::new(boundary, density, texture)
 public:
  ConstantMedium(HittablePtr boundary, float density, TexturePtr texture) : boundary_(boundary), density_(density), tex_(texture) {}
  void flatten(Flattener& f) const override;

 private:
  HittablePtr boundary_;
  float density_;
  TexturePtr tex_;
};
// Transformable sugar (transformations.rs:155-172)
inline HittablePtr rotate_y(HittablePtr h, float deg) { return std::make_shared<YRotation>(h, deg); }
inline HittablePtr translate(HittablePtr h, Vec3 off) { return std::make_shared<Translation>(h, off); }

// load_wavefront_obj (triangular.rs:241-260): OBJ (+MTL) -> BvhNode of triangles in file order.
// Geometry without `usemtl` gets DiffuseLight(SolidColor(1,0,1)) (triangular.rs:177-182);
// MTL materials must be illum 1 with a map_Kd -> Lambertian<ImageTexture> (triangular.rs:299-312).
// `override_material`, if set, replaces every material (how the bench variants are built).
HittablePtr load_wavefront_obj(const std::string& path, MaterialPtr override_material = nullptr);
// the mesh only (no BvhNode wrapper): triangles in file order
std::shared_ptr<TriangleMesh> load_mesh(const std::string& path, MaterialPtr material);

// ---- camera.rs ---------------------------------------------------------------------------------------
struct Camera {
  rtw_camera c;
  // camera.rs:25-64
  Camera(Point3 look_from, Point3 look_at, Vec3 up, float vfov, float aspect_ratio, float aperture, float focus_dist,
         float time0, float time1);
};

// ---- lib.rs ------------------------------------------------------------------------------------------
struct Pixel {  // lib.rs:120-126
  uint32_t row, column;
  Color color;
};

struct World {  // scenes.rs:860
  HittableList objects;
  std::vector<Camera> cameras;
  Color background;
};

class Raytracer {  // lib.rs:40-76 — Raytracer::new(world, cam, background, image_width, image_height, samples_per_pixel)
 public:
  Raytracer(const HittableList& world, const Camera& cam, Color background, uint32_t image_width, uint32_t image_height,
            uint32_t samples_per_pixel)
      : world_(world), cam_(cam), bg_(background), w_(image_width), h_(image_height), spp_(samples_per_pixel) {}
  // flatten + build + render on the sink's backend; Pixels in the reference's order
  // ((0..h).rev() x (0..w), lib.rs:58), color = un-normalised sum over spp.
  std::vector<Pixel> render(rtw_sink* sink, uint64_t seed = 0, rtw_render_stats* stats = nullptr) const;

 private:
  const HittableList& world_;
  const Camera& cam_;
  Color bg_;
  uint32_t w_, h_, spp_;
};

// lib.rs:128-138
struct ProgressMessage {
  enum Kind : uint32_t { ImageStart = 0, PixelMsg = 1, ImageEnd = 2 } kind;
  uint32_t width = 0, height = 0, samples_per_pixel = 0;  // ImageStart
  Pixel pixel{};                                          // Pixel
};
// postcard 0.7.3 `to_vec_cobs` bytes of one message (see include/rtw_sink.h)
std::vector<uint8_t> to_vec_cobs(const ProgressMessage& m);
// the accumulation buffer of a frame as the reference's Pixel sequence ((0..h).rev() x (0..w), lib.rs:58)
std::vector<Pixel> pixels_from_accum(const float* accum_rgb, uint32_t w, uint32_t h);

// main.rs:48-95 for every camera of a world, with the scene kept resident on the backend (flattened and built
// once): frame i = cameras[i], stream seed + i.  on_frame(frame_no, pixels, stats) runs on a helper thread while
// the next frame renders; returning false stops the animation.  Returns the number of frames delivered.
using FrameFn = std::function<bool(uint32_t, const std::vector<Pixel>&, const rtw_render_stats&)>;
// gpus > 1: every frame is spread over that many devices of the box (rtw_render_params::gpus) — same bits.
uint32_t render_animation(const World& world, uint32_t image_width, uint32_t image_height, uint32_t samples_per_pixel,
                          rtw_sink* sink, uint64_t seed, const FrameFn& on_frame, uint32_t gpus = 1);

// flatten a whole world (top-level list order = canonical order) and build it
void flatten_world(const HittableList& world, rtw_sink* sink, float time0 = 0.f, float time1 = 1.f,
                   rtw_build_stats* stats = nullptr);

// ---- asset registry: lets a harness hand decoded images / meshes to the scenes by path -------------
void register_image(const std::string& path, std::vector<uint8_t> rgb, uint32_t w, uint32_t h);
void register_mesh(const std::string& path, std::vector<float> verts, std::vector<float> normals, std::vector<float> uvs);
void set_asset_dir(const std::string& dir);  // where *.rtwm / *.rtwi fixtures live (default "assets")

// ---- scenes.rs ---------------------------------------------------------------------------------------
std::vector<std::string> scene_names();
// aspect_ratio like Scene::generate (scenes.rs:42-60); `seed` replaces thread_rng()
World generate_scene(const std::string& name, float aspect_ratio, uint64_t seed);

}  // namespace rtwh
