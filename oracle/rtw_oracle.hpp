// TEST INFRASTRUCTURE — CPU oracle. Not part of the product; see oracle/README.md.
//
// Op-for-op C++ restatement of the render hot path of AndreasKarg/raytracer-weekend
// (raytracer_weekend_lib/src/*.rs).  Every function cites the Rust lines it follows.
// Must be compiled with  -ffp-contract=off -fno-fast-math  (rustc never fuses mul+add).
// PARITY UNPINNED BY THE REFERENCE: the reference has no tests / golden vectors for this path
//  and cannot be built here (no Rust toolchain), so this file is itself the
// pin; its own known-answer tests live in tests/test_oracle_*.py.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

#include "philox.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------
// vec3.rs
// ---------------------------------------------------------------------------------------------
struct Vec3 {
  float e[3];
  Vec3() : e{0.f, 0.f, 0.f} {}
  Vec3(float a, float b, float c) : e{a, b, c} {}
  float x() const { return e[0]; }
  float y() const { return e[1]; }
  float z() const { return e[2]; }
  float operator[](int i) const { return e[i]; }
  float& operator[](int i) { return e[i]; }
  // vec3.rs:41-44
  float length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
  // vec3.rs:46-48
  float dot(const Vec3& r) const { return e[0] * r.e[0] + e[1] * r.e[1] + e[2] * r.e[2]; }
  // vec3.rs:50-56
  Vec3 cross(const Vec3& r) const {
    return Vec3(e[1] * r.e[2] - e[2] * r.e[1], e[2] * r.e[0] - e[0] * r.e[2], e[0] * r.e[1] - e[1] * r.e[0]);
  }
  // vec3.rs:58-62
  float internal_product() const { return e[0] * e[1] * e[2]; }
  // vec3.rs:81-83
  float length() const { return std::sqrt(length_squared()); }
  Vec3 unit_vector() const;  // vec3.rs:85-87
  // vec3.rs:133-138
  bool is_near_zero() const {
    const float S = 1e-8f;
    return (std::fabs(e[0]) < S) && (std::fabs(e[1]) < S) && (std::fabs(e[2]) < S);
  }
  Vec3 reflect(const Vec3& n) const;                       // vec3.rs:140-142
  Vec3 refract(const Vec3& n, float eta_i_over_eta_t) const;  // vec3.rs:144-151
};
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }  // :205-215
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }  // :217-227
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.e[0] * s, a.e[1] * s, a.e[2] * s); }                // :237-248
inline Vec3 operator*(float s, Vec3 a) { return a * s; }                                                    // :250-256
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }  // :258-268
inline Vec3 operator/(Vec3 a, float s) { return Vec3(a.e[0] / s, a.e[1] / s, a.e[2] / s); }                // :278-284
inline Vec3 operator-(Vec3 a) { return Vec3(-a.e[0], -a.e[1], -a.e[2]); }                                   // :306-312
inline Vec3 Vec3::unit_vector() const { return *this / length(); }
inline Vec3 Vec3::reflect(const Vec3& n) const { return *this - 2.0f * this->dot(n) * n; }
inline Vec3 Vec3::refract(const Vec3& n, float eta) const {
  Vec3 uv = *this;
  float cos_theta = std::fmin((-uv).dot(n), 1.0f);
  Vec3 r_out_perp = eta * (uv + cos_theta * n);
  Vec3 r_out_parallel = -std::sqrt(std::fabs(1.0f - r_out_perp.length_squared())) * n;
  return r_out_perp + r_out_parallel;
}
typedef Vec3 Point3;
typedef Vec3 Color;

// vec3.rs:93-99
inline Vec3 random_min_max(Rng& rng, float lo, float hi) {
  float a = rng.gen_range(lo, hi);
  float b = rng.gen_range(lo, hi);
  float c = rng.gen_range(lo, hi);
  return Vec3(a, b, c);
}
// vec3.rs:89-91
inline Vec3 random_vec(Rng& rng) { return random_min_max(rng, 0.0f, 1.0f); }
// vec3.rs:101-108
inline Vec3 random_in_unit_sphere(Rng& rng) {
  for (;;) {
    Vec3 p = random_min_max(rng, -1.0f, 1.0f);
    if (p.length_squared() < 1.0f) return p;
  }
}
// vec3.rs:110-112
inline Vec3 random_unit_vector(Rng& rng) { return random_in_unit_sphere(rng).unit_vector(); }
// vec3.rs:124-131
inline Vec3 random_in_unit_disk(Rng& rng) {
  for (;;) {
    float a = rng.gen_range(-1.0f, 1.0f);
    float b = rng.gen_range(-1.0f, 1.0f);
    Vec3 p(a, b, 0.0f);
    if (p.length_squared() < 1.0f) return p;
  }
}

// texture.rs:13-39
struct Point2d {
  float u, v;
};
inline Point2d operator*(float s, Point2d p) { return Point2d{s * p.u, s * p.v}; }
inline Point2d operator+(Point2d a, Point2d b) { return Point2d{a.u + b.u, a.v + b.v}; }

// ---------------------------------------------------------------------------------------------
// ray.rs
// ---------------------------------------------------------------------------------------------
struct Ray {
  Point3 origin;
  Vec3 direction;
  float time;
  Ray() : time(0.f) {}
  Ray(Point3 o, Vec3 d, float t) : origin(o), direction(d), time(t) {}
  Point3 at(float t) const { return origin + t * direction; }  // ray.rs:25-27
};

// Rust f32::max / f32::min: if one operand is NaN the other is returned (== C fmaxf/fminf)
inline float rmax(float a, float b) { return std::fmax(a, b); }
inline float rmin(float a, float b) { return std::fmin(a, b); }

// ---------------------------------------------------------------------------------------------
// aabb.rs
// ---------------------------------------------------------------------------------------------
struct Counters {
  uint64_t box_tests = 0;
  uint64_t prim_tests = 0;
};
extern thread_local Counters g_counters;

struct Aabb {
  Point3 minimum, maximum;
  Aabb() {}
  Aabb(Point3 a, Point3 b) : minimum(a), maximum(b) {}
  // aabb.rs:23-48
  bool hit(const Ray& ray, float t_min, float t_max) const {
    ++g_counters.box_tests;
    for (int a = 0; a < 3; ++a) {
      float inverted_denominator = 1.0f / ray.direction[a];
      float t0 = (minimum[a] - ray.origin[a]) * inverted_denominator;
      float t1 = (maximum[a] - ray.origin[a]) * inverted_denominator;
      if (inverted_denominator < 0.0f) std::swap(t0, t1);
      t_min = rmax(t0, t_min);
      t_max = rmin(t1, t_max);
      if (t_max <= t_min) return false;
    }
    return true;
  }
  // aabb.rs:74-88
  static Aabb surrounding_box(const Aabb& b1, const Aabb& b2) {
    Point3 small(rmin(b1.minimum.x(), b2.minimum.x()), rmin(b1.minimum.y(), b2.minimum.y()),
                 rmin(b1.minimum.z(), b2.minimum.z()));
    Point3 big(rmax(b1.maximum.x(), b2.maximum.x()), rmax(b1.maximum.y(), b2.maximum.y()),
               rmax(b1.maximum.z(), b2.maximum.z()));
    return Aabb(small, big);
  }
};

// ---------------------------------------------------------------------------------------------
// texture.rs, perlin.rs, image_texture.rs
// ---------------------------------------------------------------------------------------------
struct Texture {
  virtual ~Texture() {}
  virtual Color value(Point2d uv, const Vec3& p) const = 0;  // texture.rs:41-43
};
typedef std::shared_ptr<const Texture> TexturePtr;

struct SolidColor : Texture {  // texture.rs:45-60
  Color color_value;
  explicit SolidColor(Color c) : color_value(c) {}
  Color value(Point2d, const Vec3&) const override { return color_value; }
};

struct Checker : Texture {  // texture.rs:62-81 (constructor order: odd, even, frequency)
  TexturePtr odd, even;
  float frequency;
  Checker(TexturePtr o, TexturePtr e, float f) : odd(o), even(e), frequency(f) {}
  Color value(Point2d uv, const Vec3& p) const override {
    float sines = std::sin(frequency * p.x()) * std::sin(frequency * p.y()) * std::sin(frequency * p.z());
    if (sines < 0.0f) return odd->value(uv, p);
    return even->value(uv, p);
  }
};

struct Perlin {  // perlin.rs:9-122
  Vec3 gradients[256];
  uint64_t permutations[3][256];  // usize

  // perlin.rs:119-122
  static Point3 filter_hermit(Point3 p) {
    Vec3 offset(3.0f, 3.0f, 3.0f);
    return p * p * (offset - 2.0f * p);
  }
  // perlin.rs:91-117
  static float perlin_interp(const Vec3 gradient_cube[2][2][2], Vec3 point_within_lattice_cell) {
    point_within_lattice_cell = filter_hermit(point_within_lattice_cell);
    float accum = 0.0f;
    for (int aisle = 0; aisle < 2; ++aisle)
      for (int row = 0; row < 2; ++row)
        for (int column = 0; column < 2; ++column) {
          Vec3 cur((float)aisle, (float)row, (float)column);
          Vec3 weight_v = point_within_lattice_cell - cur;  // [QUIRK] filtered offset, perlin.rs:103
          Vec3 unit_vector(1.0f, 1.0f, 1.0f);
          Vec3 blend = cur * point_within_lattice_cell + (unit_vector - cur) * (unit_vector - point_within_lattice_cell);
          float blend_factor = blend.internal_product();
          accum += blend_factor * gradient_cube[aisle][row][column].dot(weight_v);
        }
    return accum;
  }
  // perlin.rs:50-75
  float noise(const Point3& p_in) const {
    Point3 p = p_in;
    Vec3 fl(std::floor(p.e[0]), std::floor(p.e[1]), std::floor(p.e[2]));
    uint64_t base[3];
    for (int a = 0; a < 3; ++a) {
      // `as i64` saturates (NaN -> 0), `as usize` reinterprets (vec3.rs:158-174)
      float f = fl.e[a];
      int64_t i;
      if (f != f) i = 0;
      else if (f >= 9223372036854775808.0f) i = INT64_MAX;
      else if (f <= -9223372036854775808.0f) i = INT64_MIN;
      else i = (int64_t)f;
      base[a] = (uint64_t)i;
    }
    Vec3 within = p - fl;
    Vec3 cube[2][2][2];
    for (int xo = 0; xo < 2; ++xo)
      for (int yo = 0; yo < 2; ++yo)
        for (int zo = 0; zo < 2; ++zo) {
          uint64_t lx = (base[0] + (uint64_t)xo) & 255;  // overflowing_add(..).0 & 255
          uint64_t ly = (base[1] + (uint64_t)yo) & 255;
          uint64_t lz = (base[2] + (uint64_t)zo) & 255;
          uint64_t hash = permutations[0][lx] ^ permutations[1][ly] ^ permutations[2][lz];
          cube[xo][yo][zo] = gradients[hash];
        }
    return perlin_interp(cube, within);
  }
  // perlin.rs:77-89
  float turbulence(const Point3& p, int depth) const {
    float accum = 0.0f;
    Point3 temp_p = p;
    float weight = 1.0f;
    for (int i = 0; i < depth; ++i) {
      accum += weight * noise(temp_p);
      weight *= 0.5f;
      temp_p = temp_p * 2.0f;
    }
    return std::fabs(accum);
  }
};

struct Noise : Texture {  // texture.rs:83-95
  Perlin noise;
  float scale;
  Color value(Point2d, const Vec3& p) const override {
    return Color(1.0f, 1.0f, 1.0f) * 0.5f * (1.0f + std::sin(scale * p.z() + 10.0f * noise.turbulence(p, 7)));
  }
};

struct UVDebug : Texture {  // texture.rs:97-104
  Color value(Point2d uv, const Vec3&) const override { return Color(uv.u, uv.v, 0.0f); }
};

// Rust `as u32` from f32: saturating, NaN -> 0
inline uint32_t f32_as_u32(float f) {
  if (!(f > 0.0f)) return 0;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}

struct ImageTexture : Texture {  // image_texture.rs:17-51
  std::vector<uint8_t> rgb;
  uint32_t width = 0, height = 0;
  Color value(Point2d uv, const Vec3&) const override {
    float u = std::fmin(std::fmax(uv.u, 0.0f), 1.0f);          // clamp(0,1); NaN stays NaN in Rust,
    if (uv.u != uv.u) u = uv.u;                                  //   then `as u32` gives 0
    float vc = std::fmin(std::fmax(uv.v, 0.0f), 1.0f);
    if (uv.v != uv.v) vc = uv.v;
    float v = 1.0f - vc;
    uint32_t i = std::min(f32_as_u32(u * (float)width), width - 1);
    uint32_t j = std::min(f32_as_u32(v * (float)height), height - 1);
    const float color_scale = 1.0f / 255.0f;
    const uint8_t* px = &rgb[((size_t)j * width + i) * 3];
    return Color((float)px[0] * color_scale, (float)px[1] * color_scale, (float)px[2] * color_scale);
  }
};

// ---------------------------------------------------------------------------------------------
// hittable/mod.rs, material.rs, light_source.rs
// ---------------------------------------------------------------------------------------------
struct Material;

struct HitRecord {  // hittable/mod.rs:22-29 (+ prim_id, which the reference does not carry)
  Point3 p;
  Vec3 normal;
  const Material* material = nullptr;
  float t = 0.f;
  Point2d texture_uv{0.f, 0.f};
  bool is_front_face = false;
  int32_t prim_id = -1;

  // hittable/mod.rs:32-48
  static HitRecord new_with_face_normal(Point3 p, float t, Point2d uv, const Material* m, const Ray& ray,
                                        Vec3 outward_normal, int32_t prim_id) {
    HitRecord h;
    h.is_front_face = ray.direction.dot(outward_normal) < 0.0f;
    h.normal = h.is_front_face ? outward_normal : -outward_normal;
    h.p = p; h.t = t; h.texture_uv = uv; h.material = m; h.prim_id = prim_id;
    return h;
  }
};

struct Scatter {  // material.rs:18-21
  Color attenuation;
  Ray scattered_ray;
};

struct Material {  // material.rs:23-26
  int32_t id = -1;
  virtual ~Material() {}
  virtual bool scatter(const Ray& r_in, const HitRecord& rec, Rng& rng, Scatter& out) const = 0;
  virtual Color emitted(Point2d, const Point3&) const { return Color(0.f, 0.f, 0.f); }  // material.rs:170-172
};
typedef std::shared_ptr<const Material> MaterialPtr;

struct Lambertian : Material {  // material.rs:30-61
  TexturePtr albedo;
  explicit Lambertian(TexturePtr a) : albedo(a) {}
  bool scatter(const Ray& r_in, const HitRecord& rec, Rng& rng, Scatter& out) const override {
    Vec3 scatter_direction = rec.normal + random_unit_vector(rng);
    if (scatter_direction.is_near_zero()) scatter_direction = rec.normal;
    out.scattered_ray = Ray(rec.p, scatter_direction, r_in.time);
    out.attenuation = albedo->value(rec.texture_uv, rec.p);
    return true;
  }
};

struct Metal : Material {  // material.rs:63-100
  Color albedo;
  float fuzz;
  Metal(Color a, float f) : albedo(a), fuzz(f) {}
  bool scatter(const Ray& r_in, const HitRecord& rec, Rng& rng, Scatter& out) const override {
    Vec3 reflected = r_in.direction.unit_vector().reflect(rec.normal);
    out.scattered_ray = Ray(rec.p, reflected + fuzz * random_in_unit_sphere(rng), r_in.time);
    out.attenuation = albedo;
    return out.scattered_ray.direction.dot(rec.normal) > 0.0f;
  }
};

struct Dielectric : Material {  // material.rs:102-147
  float ir;
  explicit Dielectric(float i) : ir(i) {}
  // material.rs:108-112 ; powi(5) = x * ((x*x)*(x*x)) (LLVM powi expansion / compiler-rt __powisf2)
  static float reflectance(float cosine, float ref_idx) {
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    float x = 1.0f - cosine;
    float x2 = x * x;
    float x4 = x2 * x2;
    return r0 + (1.0f - r0) * (x * x4);
  }
  bool scatter(const Ray& r_in, const HitRecord& rec, Rng& rng, Scatter& out) const override {
    out.attenuation = Color(1.0f, 1.0f, 1.0f);
    float refraction_ratio = rec.is_front_face ? 1.0f / ir : ir;
    Vec3 unit_direction = r_in.direction.unit_vector();
    float cos_theta = rmin((-unit_direction).dot(rec.normal), 1.0f);
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    bool cannot_refract = (refraction_ratio * sin_theta) > 1.0f;
    Vec3 direction;
    // `||` short-circuits: no draw when cannot_refract (material.rs:128-129)
    if (cannot_refract || reflectance(cos_theta, refraction_ratio) > rng.gen_f32())
      direction = unit_direction.reflect(rec.normal);
    else
      direction = unit_direction.refract(rec.normal, refraction_ratio);
    out.scattered_ray = Ray(rec.p, direction, r_in.time);
    return true;
  }
};

struct Isotropic : Material {  // material.rs:149-168
  TexturePtr albedo;
  explicit Isotropic(TexturePtr a) : albedo(a) {}
  bool scatter(const Ray& r_in, const HitRecord& rec, Rng& rng, Scatter& out) const override {
    out.attenuation = albedo->value(rec.texture_uv, rec.p);
    out.scattered_ray = Ray(rec.p, random_in_unit_sphere(rng), r_in.time);
    return true;
  }
};

struct DiffuseLight : Material {  // light_source.rs:13-24
  TexturePtr emit;
  explicit DiffuseLight(TexturePtr e) : emit(e) {}
  bool scatter(const Ray&, const HitRecord&, Rng&, Scatter&) const override { return false; }
  Color emitted(Point2d uv, const Point3& p) const override { return emit->value(uv, p); }
};

struct Hittable {  // hittable/mod.rs:51-54
  virtual ~Hittable() {}
  virtual bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& rec) const = 0;
  virtual bool bounding_box(float time0, float time1, Aabb& out) const = 0;
};
typedef std::unique_ptr<Hittable> HittablePtr;

// hittable/mod.rs:56-88  (`impl Hittable for [Box<dyn Hittable>]`)
struct HittableList : Hittable {
  std::vector<HittablePtr> objects;
  bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& rec) const override {
    float closest_so_far = t_max;
    bool any = false;
    HitRecord temp;
    for (const auto& object : objects) {
      if (object->hit(r, t_min, closest_so_far, rng, temp)) {
        closest_so_far = temp.t;
        rec = temp;
        any = true;
      }
    }
    return any;
  }
  bool bounding_box(float t0, float t1, Aabb& out) const override {
    if (objects.empty()) return false;
    bool have = false;
    Aabb acc;
    for (const auto& object : objects) {
      Aabb tmp;
      if (!object->bounding_box(t0, t1, tmp)) return false;
      acc = have ? Aabb::surrounding_box(acc, tmp) : tmp;
      have = true;
    }
    out = acc;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// hittable/spherical.rs
// ---------------------------------------------------------------------------------------------
// spherical.rs:62-77
inline Point2d get_sphere_uv(const Point3& p) {
  const float PI = 3.14159274101257324219f;  // core::f32::consts::PI
  float theta = std::acos(-p.y());
  float phi = std::atan2(-p.z(), p.x()) + PI;
  float u = phi / (2.0f * PI);
  float v = theta / PI;
  return Point2d{u, v};
}
// spherical.rs:18-60
inline bool hit_sphere(const Ray& ray, float t_min, float t_max, Vec3 center, float radius, const Material* material,
                       int32_t prim_id, HitRecord& rec) {
  ++g_counters.prim_tests;
  Vec3 origin_to_center = ray.origin - center;
  float a = ray.direction.length_squared();
  float half_b = origin_to_center.dot(ray.direction);
  float c = origin_to_center.length_squared() - radius * radius;
  float discriminant = half_b * half_b - a * c;  // half_b.powi(2)
  if (discriminant < 0.0f) return false;
  float sqrtd = std::sqrt(discriminant);
  float root = (-half_b - sqrtd) / a;
  if (root < t_min || t_max < root) {
    root = (-half_b + sqrtd) / a;
    if (root < t_min || t_max < root) return false;
  }
  float t = root;
  Point3 hit_point = ray.at(root);
  Vec3 outward_normal = (hit_point - center) / radius;
  Point2d uv = get_sphere_uv(outward_normal);
  rec = HitRecord::new_with_face_normal(hit_point, t, uv, material, ray, outward_normal, prim_id);
  return true;
}

struct Sphere : Hittable {  // spherical.rs:80-105
  Point3 center;
  float radius;
  MaterialPtr material;
  int32_t prim_id;
  bool hit(const Ray& ray, float t_min, float t_max, Rng&, HitRecord& rec) const override {
    return hit_sphere(ray, t_min, t_max, center, radius, material.get(), prim_id, rec);
  }
  bool bounding_box(float, float, Aabb& out) const override {
    Vec3 rv(radius, radius, radius);
    out = Aabb(center - rv, center + rv);  // [QUIRK] min > max for negative radius
    return true;
  }
};

struct MovingSphere : Hittable {  // spherical.rs:107-151
  Point3 center0;
  float time0;
  Point3 center1;
  float time1;
  float radius;
  MaterialPtr material;
  int32_t prim_id;
  // spherical.rs:117-123
  Point3 center_at_time(float time) const { return center0 + ((time - time0) / (time1 - time0)) * (center1 - center0); }
  bool hit(const Ray& ray, float t_min, float t_max, Rng&, HitRecord& rec) const override {
    return hit_sphere(ray, t_min, t_max, center_at_time(ray.time), radius, material.get(), prim_id, rec);
  }
  bool bounding_box(float t0, float t1, Aabb& out) const override {
    Point3 sc = center_at_time(t0), ec = center_at_time(t1);
    Vec3 rv(radius, radius, radius);
    out = Aabb::surrounding_box(Aabb(sc - rv, sc + rv), Aabb(ec - rv, ec + rv));
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// hittable/rectangular.rs
// ---------------------------------------------------------------------------------------------
// AXIS = the constant axis: 2 -> XYRectangle (:16-65), 1 -> XZRectangle (:67-116), 0 -> YZRectangle (:118-167).
// (A, B) are the in-plane axes in the order the reference names them: XY->(0,1), XZ->(0,2), YZ->(1,2).
struct AxisRect : Hittable {
  int axis;  // constant axis
  float a0, a1, b0, b1, k;
  MaterialPtr material;
  int32_t prim_id;
  bool hit(const Ray& r, float t_min, float t_max, Rng&, HitRecord& rec) const override {
    ++g_counters.prim_tests;
    const int A = (axis == 0) ? 1 : 0;
    const int B = (axis == 2) ? 1 : 2;
    float t = (k - r.origin[axis]) / r.direction[axis];
    if (t < t_min || t > t_max) return false;
    float a = r.origin[A] + t * r.direction[A];
    float b = r.origin[B] + t * r.direction[B];
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    float u = (a - a0) / (a1 - a0);
    float v = (b - b0) / (b1 - b0);
    Vec3 outward_normal(0.f, 0.f, 0.f);
    outward_normal[axis] = 1.0f;
    Point3 p = r.at(t);
    rec = HitRecord::new_with_face_normal(p, t, Point2d{u, v}, material.get(), r, outward_normal, prim_id);
    return true;
  }
  bool bounding_box(float, float, Aabb& out) const override {
    const int A = (axis == 0) ? 1 : 0;
    const int B = (axis == 2) ? 1 : 2;
    Point3 lo, hi;
    lo[A] = a0; hi[A] = a1; lo[B] = b0; hi[B] = b1;
    lo[axis] = k - 0.0001f; hi[axis] = k + 0.0001f;
    out = Aabb(lo, hi);
    return true;
  }
};

struct Cuboid : Hittable {  // rectangular.rs:170-245
  Point3 box_min, box_max;
  HittableList sides;
  bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& rec) const override {
    return sides.hit(r, t_min, t_max, rng, rec);
  }
  bool bounding_box(float, float, Aabb& out) const override {
    out = Aabb(box_min, box_max);
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// hittable/triangular.rs
// ---------------------------------------------------------------------------------------------
struct Triangle : Hittable {  // triangular.rs:34-149
  Point3 vertices[3];
  Vec3 normals[3];
  Point2d texture_uv[3];
  MaterialPtr material;
  int32_t prim_id;

  // triangular.rs:42-73: has_normals / has_uvs = false -> defaults
  void init(const float* v9, const float* n9, const float* uv6) {
    for (int i = 0; i < 3; ++i) vertices[i] = Point3(v9[3 * i], v9[3 * i + 1], v9[3 * i + 2]);
    Vec3 a_to_b = vertices[1] - vertices[0];
    Vec3 a_to_c = vertices[2] - vertices[0];
    Vec3 triangle_normal = a_to_b.cross(a_to_c);  // [QUIRK] un-normalised
    for (int i = 0; i < 3; ++i) normals[i] = n9 ? Vec3(n9[3 * i], n9[3 * i + 1], n9[3 * i + 2]) : triangle_normal;
    const Point2d def[3] = {{0.f, 0.f}, {1.f, 0.f}, {0.f, 1.f}};
    for (int i = 0; i < 3; ++i) texture_uv[i] = uv6 ? Point2d{uv6[2 * i], uv6[2 * i + 1]} : def[i];
  }
  // triangular.rs:315-323
  template <class T>
  static T interpolate_barycentric(float u, float v, const T x[3]) {
    return (1.0f - u - v) * x[0] + u * x[1] + v * x[2];
  }
  // triangular.rs:97-138
  bool hit(const Ray& ray, float t_min, float t_max, Rng&, HitRecord& rec) const override {
    ++g_counters.prim_tests;
    Vec3 vertex_a = vertices[0], vertex_b = vertices[1], vertex_c = vertices[2];
    Vec3 a_to_b = vertex_b - vertex_a;
    Vec3 a_to_c = vertex_c - vertex_a;
    Vec3 normal = a_to_b.cross(a_to_c);
    float determinant = -ray.direction.dot(normal);
    float inv_determinant = 1.0f / determinant;
    Vec3 a_to_ray_origin = ray.origin - vertex_a;
    Vec3 dao = a_to_ray_origin.cross(ray.direction);
    float u = a_to_c.dot(dao) * inv_determinant;
    float v = -a_to_b.dot(dao) * inv_determinant;
    float t = a_to_ray_origin.dot(normal) * inv_determinant;
    if (t < t_min || t > t_max) return false;
    bool triangle_was_hit = t >= 0.0f && u >= 0.0f && v >= 0.0f && (u + v) <= 1.0f;
    if (!triangle_was_hit) return false;
    Point3 p = ray.at(t);
    Vec3 hit_normal = interpolate_barycentric(u, v, normals);
    Point2d hit_uv = interpolate_barycentric(u, v, texture_uv);
    rec = HitRecord::new_with_face_normal(p, t, hit_uv, material.get(), ray, hit_normal, prim_id);
    return true;
  }
  // triangular.rs:79-93 (itertools minmax over the three coordinates, then the thin-axis pad)
  static void min_max(float a, float b, float c, float& lo, float& hi) {
    // itertools::minmax: pairwise; equivalent to min/max for non-NaN input
    lo = a; hi = a;
    if (b < lo) lo = b; else if (b >= hi) hi = b;
    if (c < lo) lo = c; else if (c >= hi) hi = c;
    if (std::fabs(lo - hi) < 0.0002f) { lo = lo - 0.0001f; hi = hi + 0.0001f; }
  }
  // triangular.rs:140-149
  bool bounding_box(float, float, Aabb& out) const override {
    Point3 lo, hi;
    for (int a = 0; a < 3; ++a) min_max(vertices[0][a], vertices[1][a], vertices[2][a], lo[a], hi[a]);
    out = Aabb(lo, hi);
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// hittable/transformations.rs
// ---------------------------------------------------------------------------------------------
struct Translation : Hittable {  // transformations.rs:16-48
  HittablePtr inner;
  Vec3 offset;
  bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& rec) const override {
    Ray translated_ray(r.origin - offset, r.direction, r.time);
    HitRecord h;
    if (!inner->hit(translated_ray, t_min, t_max, rng, h)) return false;
    Point3 translated_hitpoint = h.p + offset;
    // [QUIRK] the face normal is re-evaluated on an already ray-facing normal (:30-37)
    rec = HitRecord::new_with_face_normal(translated_hitpoint, h.t, h.texture_uv, h.material, translated_ray, h.normal,
                                          h.prim_id);
    return true;
  }
  bool bounding_box(float t0, float t1, Aabb& out) const override {
    Aabb b;
    if (!inner->bounding_box(t0, t1, b)) return false;
    out = Aabb(b.minimum + offset, b.maximum + offset);
    return true;
  }
};

struct YRotation : Hittable {  // transformations.rs:50-153
  HittablePtr inner;
  float sin_theta, cos_theta;
  bool has_box = false;
  Aabb bbox;
  // transformations.rs:77-111
  static Aabb rotate_bounding_box(const Aabb& b, float sin_theta, float cos_theta) {
    const float INF = std::numeric_limits<float>::infinity();
    Point3 mn(INF, INF, INF), mx(-INF, -INF, -INF);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) {
          float fi = (float)i, fj = (float)j, fk = (float)k;
          float x = fi * b.maximum.x() + (1.0f - fi) * b.minimum.x();
          float y = fj * b.maximum.y() + (1.0f - fj) * b.minimum.y();
          float z = fk * b.maximum.z() + (1.0f - fk) * b.minimum.z();
          float new_x = cos_theta * x + sin_theta * z;
          float new_z = -sin_theta * x + cos_theta * z;
          Vec3 tester(new_x, y, new_z);
          for (int axis = 0; axis < 3; ++axis) {
            mn[axis] = rmin(mn[axis], tester[axis]);
            mx[axis] = rmax(mx[axis], tester[axis]);
          }
        }
    return Aabb(mn, mx);
  }
  // transformations.rs:59-75 ; f32::to_radians = x * (PI/180) with the constant folded in f32
  void init(HittablePtr in, float angle_degrees) {
    inner = std::move(in);
    const float RADS_PER_DEG = 3.14159274101257324219f / 180.0f;
    float angle_radians = angle_degrees * RADS_PER_DEG;
    sin_theta = std::sin(angle_radians);
    cos_theta = std::cos(angle_radians);
    Aabb b;
    has_box = inner->bounding_box(0.0f, 1.0f, b);
    if (has_box) bbox = rotate_bounding_box(b, sin_theta, cos_theta);
  }
  // the struct as it is stored (transformations.rs:51-56): sin / cos given, box from them (:64-75)
  void init_sincos(HittablePtr in, float s, float c) {
    inner = std::move(in);
    sin_theta = s;
    cos_theta = c;
    Aabb b;
    has_box = inner->bounding_box(0.0f, 1.0f, b);
    if (has_box) bbox = rotate_bounding_box(b, sin_theta, cos_theta);
  }
  // transformations.rs:115-148
  bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& out) const override {
    Point3 origin = r.origin;
    Vec3 direction = r.direction;
    origin[0] = cos_theta * r.origin[0] - sin_theta * r.origin[2];
    origin[2] = sin_theta * r.origin[0] + cos_theta * r.origin[2];
    direction[0] = cos_theta * r.direction[0] - sin_theta * r.direction[2];
    direction[2] = sin_theta * r.direction[0] + cos_theta * r.direction[2];
    Ray rotated_r(origin, direction, r.time);
    HitRecord rec;
    if (!inner->hit(rotated_r, t_min, t_max, rng, rec)) return false;
    Point3 p = rec.p;
    Vec3 normal = rec.normal;
    p[0] = cos_theta * rec.p[0] + sin_theta * rec.p[2];
    p[2] = -sin_theta * rec.p[0] + cos_theta * rec.p[2];
    normal[0] = cos_theta * rec.normal[0] + sin_theta * rec.normal[2];
    normal[2] = -sin_theta * rec.normal[0] + cos_theta * rec.normal[2];
    // [QUIRK] face normal evaluated with the ROTATED ray against the world-space normal (:140-147)
    out = HitRecord::new_with_face_normal(p, rec.t, rec.texture_uv, rec.material, rotated_r, normal, rec.prim_id);
    return true;
  }
  bool bounding_box(float, float, Aabb& out) const override {
    if (!has_box) return false;
    out = bbox;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// hittable/volumes.rs
// ---------------------------------------------------------------------------------------------
struct ConstantMedium : Hittable {  // volumes.rs:18-83
  HittablePtr boundary;
  std::shared_ptr<Isotropic> phase_function;
  float neg_inv_density;
  int32_t prim_id;
  // volumes.rs:38-78.  The one draw (volumes.rs:58) is keyed by prim_id (see philox.hpp).
  bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& rec) const override {
    const float INF = std::numeric_limits<float>::infinity();
    HitRecord rec1, rec2;
    if (!boundary->hit(r, -INF, INF, rng, rec1)) return false;
    if (!boundary->hit(r, rec1.t + 0.0001f, INF, rng, rec2)) return false;
    float rec1_t = rmax(rec1.t, t_min);
    float rec2_t = rmin(rec2.t, t_max);
    if (rec1_t >= rec2_t) return false;
    rec1_t = rmax(rec1_t, 0.0f);
    float ray_length = r.direction.length();
    float distance_inside_boundary = (rec2_t - rec1_t) * ray_length;
    float hit_distance = neg_inv_density * std::log10(rng.gen_f32_keyed((uint32_t)prim_id));  // [QUIRK] log10, not ln
    if (hit_distance > distance_inside_boundary) return false;
    float t = rec1_t + hit_distance / ray_length;
    rec.p = r.at(t);
    rec.normal = Vec3(1.0f, 0.0f, 0.0f);  // arbitrary
    rec.material = phase_function.get();
    rec.t = t;
    rec.texture_uv = Point2d{0.0f, 0.0f};
    rec.is_front_face = true;
    rec.prim_id = prim_id;
    return true;
  }
  bool bounding_box(float t0, float t1, Aabb& out) const override { return boundary->bounding_box(t0, t1, out); }
};

// ---------------------------------------------------------------------------------------------
// bvh.rs
// ---------------------------------------------------------------------------------------------
struct BvhNode : Hittable {  // bvh.rs:12-120
  HittablePtr left;
  HittablePtr right;  // may be null (bvh.rs:37-39)
  Aabb bbox;

  // bvh.rs:88-97
  static bool box_less(const Hittable* a, const Hittable* b, int axis) {
    Aabb ba, bb;
    a->bounding_box(0.0f, 0.0f, ba);
    b->bounding_box(0.0f, 0.0f, bb);
    return ba.minimum[axis] < bb.minimum[axis];
  }
  // bvh.rs:19-74. `rng` is a host-side generator (topology is RNG dependent in the reference too).
  static HittablePtr build(std::vector<HittablePtr>& src, size_t lo, size_t hi, float time0, float time1, Rng& rng) {
    std::unique_ptr<BvhNode> node(new BvhNode());
    int axis = (int)rng.gen_below(3);  // gen_range(0..=2), drawn for leaves too (bvh.rs:25)
    size_t n = hi - lo;
    if (n == 1) {
      node->left = std::move(src[lo]);
    } else if (n == 2) {
      node->left = std::move(src[lo + 1]);  // pop() = last
      node->right = std::move(src[lo]);     // pop() = first
    } else {
      std::stable_sort(src.begin() + lo, src.begin() + hi,
                       [axis](const HittablePtr& l, const HittablePtr& r) { return box_less(l.get(), r.get(), axis); });
      size_t mid = lo + n / 2;
      node->left = build(src, lo, mid, time0, time1, rng);
      node->right = build(src, mid, hi, time0, time1, rng);
    }
    Aabb bl;
    node->left->bounding_box(time0, time1, bl);
    if (node->right) {
      Aabb br;
      node->right->bounding_box(time0, time1, br);
      node->bbox = Aabb::surrounding_box(bl, br);
    } else {
      node->bbox = bl;
    }
    return HittablePtr(node.release());
  }
  // bvh.rs:101-120
  bool hit(const Ray& r, float t_min, float t_max, Rng& rng, HitRecord& rec) const override {
    if (!bbox.hit(r, t_min, t_max)) return false;
    HitRecord hl;
    bool hit_left = left->hit(r, t_min, t_max, rng, hl);
    float tm = hit_left ? hl.t : t_max;
    HitRecord hr;
    bool hit_right = right ? right->hit(r, t_min, tm, rng, hr) : false;
    if (hit_right) { rec = hr; return true; }
    if (hit_left) { rec = hl; return true; }
    return false;
  }
  bool bounding_box(float, float, Aabb& out) const override {
    out = bbox;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// camera.rs
// ---------------------------------------------------------------------------------------------
struct Camera {  // camera.rs:8-19
  Point3 origin, lower_left_corner;
  Vec3 horizontal, vertical, u, v, w;
  float lens_radius, time0, time1;

  // camera.rs:25-64
  static Camera make(Point3 look_from, Point3 look_at, Vec3 up, float vfov, float aspect_ratio, float aperture,
                     float focus_dist, float time0, float time1) {
    Camera c;
    const float RADS_PER_DEG = 3.14159274101257324219f / 180.0f;
    float theta = vfov * RADS_PER_DEG;
    float h = std::tan(theta / 2.0f);
    float viewport_height = 2.0f * h;
    float viewport_width = aspect_ratio * viewport_height;
    c.w = (look_from - look_at).unit_vector();
    c.u = up.cross(c.w).unit_vector();
    c.v = c.w.cross(c.u);
    c.origin = look_from;
    c.horizontal = focus_dist * viewport_width * c.u;
    c.vertical = focus_dist * viewport_height * c.v;
    c.lower_left_corner = c.origin - c.horizontal / 2.0f - c.vertical / 2.0f - focus_dist * c.w;
    c.lens_radius = aperture / 2.0f;
    c.time0 = time0;
    c.time1 = time1;
    return c;
  }
  // camera.rs:66-74 (argument evaluation order: disk draws, then the time draw)
  Ray get_ray(float s, float t, Rng& rng) const {
    Vec3 rd = lens_radius * random_in_unit_disk(rng);
    Vec3 offset = u * rd.x() + v * rd.y();
    Point3 o = origin + offset;
    Vec3 d = lower_left_corner + s * horizontal + t * vertical - origin - offset;
    float time = rng.gen_range(time0, time1);
    return Ray(o, d, time);
  }
};

}  // namespace orc
