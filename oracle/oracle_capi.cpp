// TEST INFRASTRUCTURE — CPU oracle. Not part of the product; see oracle/README.md.
//
// C API of the oracle.  It mirrors include/rtw_cuda.h function for function (prefix orc_ instead
// of rtw_) so that one scene-emitting front end can feed the oracle and the CUDA backend with the
// very same calls.  On top of that it exposes the reference's render loop (lib.rs:57-117) and the
// helpers the parity tests need (ray capture, traversal of a GPU-built LBVH, Philox KATs).
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "../include/rtw_cuda.h"
#include "rtw_oracle.hpp"

namespace orc {
thread_local Counters g_counters;
}
using namespace orc;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// ---- scene description: recorded by the emit calls, instantiated twice by orc_build ----------
enum DescKind { D_ROOT, D_GROUP, D_TRANSLATE, D_ROTY, D_SPHERE, D_MSPHERE, D_RECT, D_CUBOID, D_TRIS, D_MEDIUM };
struct Desc {
  DescKind kind;
  float f[12] = {0};
  int axis = 0;
  int material = -1;
  int first_prim = -1;
  uint32_t ntris = 0;
  std::vector<float> verts, normals, uvs;
  std::vector<int32_t> mats;
  std::vector<std::unique_ptr<Desc>> children;
};

}  // namespace

struct orc_scene {
  std::vector<TexturePtr> textures;
  std::vector<std::shared_ptr<Material>> materials;
  Desc root;
  std::vector<Desc*> stack;
  int num_prims = 0;
  bool built = false;
  HittableList world_flat;  // ground truth: every BvhNode replaced by the list it was built from
  HittableList world_ref;   // the reference's structure: groups are BvhNode::new(..)
  float time0 = 0.f, time1 = 1.f;
  orc_scene() {
    root.kind = D_ROOT;
    stack.push_back(&root);
  }
};

namespace {

bool valid_tex(orc_scene* s, int t) { return t >= 0 && t < (int)s->textures.size(); }
bool valid_mat(orc_scene* s, int m) { return m >= 0 && m < (int)s->materials.size(); }

Desc* push_child(orc_scene* s, DescKind k) {
  std::unique_ptr<Desc> d(new Desc());
  d->kind = k;
  Desc* raw = d.get();
  s->stack.back()->children.push_back(std::move(d));
  return raw;
}

HittablePtr make_rect(orc_scene* s, int axis, float a0, float a1, float b0, float b1, float k, int material, int id) {
  std::unique_ptr<AxisRect> r(new AxisRect());
  r->axis = axis; r->a0 = a0; r->a1 = a1; r->b0 = b0; r->b1 = b1; r->k = k;
  r->material = s->materials[material];
  r->prim_id = id;
  return HittablePtr(r.release());
}

void instantiate(orc_scene* s, const Desc& d, bool use_bvh, Rng& build_rng, std::vector<HittablePtr>& out);

HittablePtr wrap_children(orc_scene* s, const Desc& d, bool use_bvh, Rng& build_rng) {
  std::vector<HittablePtr> kids;
  for (const auto& c : d.children) instantiate(s, *c, use_bvh, build_rng, kids);
  if (kids.size() == 1) return std::move(kids[0]);
  std::unique_ptr<HittableList> l(new HittableList());
  l->objects = std::move(kids);
  return HittablePtr(l.release());
}

void instantiate(orc_scene* s, const Desc& d, bool use_bvh, Rng& build_rng, std::vector<HittablePtr>& out) {
  switch (d.kind) {
    case D_ROOT: break;
    case D_GROUP: {
      std::vector<HittablePtr> kids;
      for (const auto& c : d.children) instantiate(s, *c, use_bvh, build_rng, kids);
      if (kids.empty()) break;
      if (use_bvh) {
        out.push_back(BvhNode::build(kids, 0, kids.size(), s->time0, s->time1, build_rng));
      } else {
        std::unique_ptr<HittableList> l(new HittableList());
        l->objects = std::move(kids);
        out.push_back(HittablePtr(l.release()));
      }
      break;
    }
    case D_TRANSLATE: {
      std::unique_ptr<Translation> t(new Translation());
      t->inner = wrap_children(s, d, use_bvh, build_rng);
      t->offset = Vec3(d.f[0], d.f[1], d.f[2]);
      out.push_back(HittablePtr(t.release()));
      break;
    }
    case D_ROTY: {
      std::unique_ptr<YRotation> r(new YRotation());
      if (d.f[3] != 0.f) r->init_sincos(wrap_children(s, d, use_bvh, build_rng), d.f[1], d.f[2]);
      else r->init(wrap_children(s, d, use_bvh, build_rng), d.f[0]);
      out.push_back(HittablePtr(r.release()));
      break;
    }
    case D_SPHERE: {
      std::unique_ptr<Sphere> sp(new Sphere());
      sp->center = Point3(d.f[0], d.f[1], d.f[2]);
      sp->radius = d.f[3];
      sp->material = s->materials[d.material];
      sp->prim_id = d.first_prim;
      out.push_back(HittablePtr(sp.release()));
      break;
    }
    case D_MSPHERE: {
      std::unique_ptr<MovingSphere> sp(new MovingSphere());
      sp->center0 = Point3(d.f[0], d.f[1], d.f[2]);
      sp->time0 = d.f[3];
      sp->center1 = Point3(d.f[4], d.f[5], d.f[6]);
      sp->time1 = d.f[7];
      sp->radius = d.f[8];
      sp->material = s->materials[d.material];
      sp->prim_id = d.first_prim;
      out.push_back(HittablePtr(sp.release()));
      break;
    }
    case D_RECT:
      out.push_back(make_rect(s, d.axis, d.f[0], d.f[1], d.f[2], d.f[3], d.f[4], d.material, d.first_prim));
      break;
    case D_CUBOID: {
      // rectangular.rs:177-234: XY@z1, XY@z0, XZ@y1, XZ@y0, YZ@x1, YZ@x0
      std::unique_ptr<Cuboid> c(new Cuboid());
      const float* p0 = &d.f[0];
      const float* p1 = &d.f[3];
      c->box_min = Point3(p0[0], p0[1], p0[2]);
      c->box_max = Point3(p1[0], p1[1], p1[2]);
      int id = d.first_prim;
      c->sides.objects.push_back(make_rect(s, 2, p0[0], p1[0], p0[1], p1[1], p1[2], d.material, id + 0));
      c->sides.objects.push_back(make_rect(s, 2, p0[0], p1[0], p0[1], p1[1], p0[2], d.material, id + 1));
      c->sides.objects.push_back(make_rect(s, 1, p0[0], p1[0], p0[2], p1[2], p1[1], d.material, id + 2));
      c->sides.objects.push_back(make_rect(s, 1, p0[0], p1[0], p0[2], p1[2], p0[1], d.material, id + 3));
      c->sides.objects.push_back(make_rect(s, 0, p0[1], p1[1], p0[2], p1[2], p1[0], d.material, id + 4));
      c->sides.objects.push_back(make_rect(s, 0, p0[1], p1[1], p0[2], p1[2], p0[0], d.material, id + 5));
      out.push_back(HittablePtr(c.release()));
      break;
    }
    case D_MEDIUM: {  // ConstantMedium::new(boundary, density, texture) (volumes.rs:24-35)
      std::unique_ptr<ConstantMedium> m(new ConstantMedium());
      m->boundary = wrap_children(s, d, use_bvh, build_rng);
      m->neg_inv_density = -1.0f / d.f[0];
      m->phase_function = std::static_pointer_cast<Isotropic>(s->materials[d.material]);
      m->prim_id = d.first_prim;
      out.push_back(HittablePtr(m.release()));
      break;
    }
    case D_TRIS: {
      for (uint32_t i = 0; i < d.ntris; ++i) {
        std::unique_ptr<Triangle> t(new Triangle());
        t->init(&d.verts[9 * (size_t)i], d.normals.empty() ? nullptr : &d.normals[9 * (size_t)i],
                d.uvs.empty() ? nullptr : &d.uvs[6 * (size_t)i]);
        int m = d.mats.empty() ? d.material : d.mats[i];
        t->material = s->materials[m];
        t->prim_id = d.first_prim + (int)i;
        out.push_back(HittablePtr(t.release()));
      }
      break;
    }
  }
}

const HittableList& world_of(const orc_scene* s, int mode) { return mode == 0 ? s->world_flat : s->world_ref; }

void fill_hit(const HitRecord& h, bool hit, rtw_hit* out) {
  std::memset(out, 0, sizeof(*out));
  if (!hit) {
    out->prim_id = -1;
    out->material_id = -1;
    return;
  }
  out->prim_id = h.prim_id;
  out->material_id = h.material ? h.material->id : -1;
  out->t = h.t;
  for (int a = 0; a < 3; ++a) {
    out->p[a] = h.p[a];
    out->normal[a] = h.normal[a];
  }
  out->u = h.texture_uv.u;
  out->v = h.texture_uv.v;
  out->front_face = h.is_front_face ? 1 : 0;
}

Camera to_camera(const rtw_camera* c) {
  Camera cam;
  cam.origin = Point3(c->origin[0], c->origin[1], c->origin[2]);
  cam.lower_left_corner = Point3(c->lower_left_corner[0], c->lower_left_corner[1], c->lower_left_corner[2]);
  cam.horizontal = Vec3(c->horizontal[0], c->horizontal[1], c->horizontal[2]);
  cam.vertical = Vec3(c->vertical[0], c->vertical[1], c->vertical[2]);
  cam.u = Vec3(c->u[0], c->u[1], c->u[2]);
  cam.v = Vec3(c->v[0], c->v[1], c->v[2]);
  cam.w = Vec3(c->w[0], c->w[1], c->w[2]);
  cam.lens_radius = c->lens_radius;
  cam.time0 = c->time0;
  cam.time1 = c->time1;
  return cam;
}

// lib.rs:84-86: the camera ray of (pixel_row, pixel_column, sample); stage 0 of the stream
Ray camera_ray(const Camera& cam, uint32_t w, uint32_t h, uint32_t row, uint32_t col, Rng& rng) {
  float u = ((float)col + rng.gen_f32()) / (float)(w - 1);
  float v = ((float)row + rng.gen_f32()) / (float)(h - 1);
  return cam.get_ray(u, v, rng);
}

// lib.rs:97-117, recursive exactly as written.  `bounce` = MAX_DEPTH - depth selects the RNG stage.
Color sample_ray_recursive(const HittableList& world, const Color& background, const Ray& r, Rng& rng, uint32_t depth,
                           uint32_t bounce, uint64_t& segments) {
  if (depth == 0) return Color(0.f, 0.f, 0.f);
  rng.set_stage(bounce + 1);
  ++segments;
  HitRecord rec;
  if (!world.hit(r, 0.001f, std::numeric_limits<float>::infinity(), rng, rec)) return background;
  Color emitted = rec.material->emitted(rec.texture_uv, rec.p);
  Scatter sc;
  if (!rec.material->scatter(r, rec, rng, sc)) return emitted;
  return emitted + sc.attenuation * sample_ray_recursive(world, background, sc.scattered_ray, rng, depth - 1, bounce + 1, segments);
}

// The iterative form of the same integrator: L = sum_b T_b * e_b with
// T_b = a_0 * ... * a_{b-1}.  Same terms as the recursion, associated left to right — this is the
// association a wavefront path tracer has to use, so it is the form that is compared bit for bit
// against the GPU; test_oracle_integrator.py bounds its distance to the recursive form.
Color sample_ray_iterative(const HittableList& world, const Color& background, Ray r, Rng& rng, uint32_t max_depth,
                           uint64_t& segments) {
  Color L(0.f, 0.f, 0.f);
  Color T(1.f, 1.f, 1.f);
  for (uint32_t bounce = 0; bounce < max_depth; ++bounce) {
    rng.set_stage(bounce + 1);
    ++segments;
    HitRecord rec;
    if (!world.hit(r, 0.001f, std::numeric_limits<float>::infinity(), rng, rec)) {
      L = L + T * background;
      break;
    }
    Color emitted = rec.material->emitted(rec.texture_uv, rec.p);
    L = L + T * emitted;
    Scatter sc;
    if (!rec.material->scatter(r, rec, rng, sc)) break;
    T = T * sc.attenuation;
    r = sc.scattered_ray;
  }
  return L;
}

void tile_grid(const rtw_render_params* p, uint32_t& ts, uint32_t& tiles_x) {
  ts = p->tile_size ? p->tile_size : 32;
  tiles_x = (p->width + ts - 1) / ts;
}
bool pixel_in_part(const rtw_render_params* p, uint32_t x, uint32_t y_top) {
  if (p->part_count <= 1) return true;
  uint32_t ts, tiles_x;
  tile_grid(p, ts, tiles_x);
  uint32_t tile = (y_top / ts) * tiles_x + (x / ts);
  return tile % p->part_count == p->part_rank;
}

}  // namespace

extern "C" {

int orc_abi_version(void) { return RTW_ABI_VERSION; }
const char* orc_last_error(void) { return g_err.c_str(); }
int orc_device_count(void) { return 0; }

int orc_scene_create(int, orc_scene** out) {
  if (!out) return fail(RTW_ERR_INVALID, "out is NULL");
  *out = new orc_scene();
  return RTW_OK;
}
int orc_scene_destroy(orc_scene* s) {
  delete s;
  return RTW_OK;
}

#define CHECK_SCENE(s)                                            \
  if (!(s)) return fail(RTW_ERR_INVALID, "scene is NULL");        \
  if ((s)->built) return fail(RTW_ERR_STATE, "scene already built")

int orc_add_texture_solid(orc_scene* s, float r, float g, float b) {
  CHECK_SCENE(s);
  s->textures.push_back(std::make_shared<SolidColor>(Color(r, g, b)));
  return (int)s->textures.size() - 1;
}
int orc_add_texture_checker(orc_scene* s, int odd, int even, float frequency) {
  CHECK_SCENE(s);
  if (!valid_tex(s, odd) || !valid_tex(s, even)) return fail(RTW_ERR_INVALID, "checker: bad texture id");
  s->textures.push_back(std::make_shared<Checker>(s->textures[odd], s->textures[even], frequency));
  return (int)s->textures.size() - 1;
}
int orc_add_texture_noise(orc_scene* s, const float* g, const int32_t* px, const int32_t* py, const int32_t* pz,
                          float scale) {
  CHECK_SCENE(s);
  if (!g || !px || !py || !pz) return fail(RTW_ERR_INVALID, "noise: NULL table");
  auto n = std::make_shared<Noise>();
  for (int i = 0; i < 256; ++i) {
    n->noise.gradients[i] = Vec3(g[3 * i], g[3 * i + 1], g[3 * i + 2]);
    if ((px[i] | py[i] | pz[i]) & ~255) return fail(RTW_ERR_INVALID, "noise: permutation value out of 0..255");
    n->noise.permutations[0][i] = (uint64_t)px[i];
    n->noise.permutations[1][i] = (uint64_t)py[i];
    n->noise.permutations[2][i] = (uint64_t)pz[i];
  }
  n->scale = scale;
  s->textures.push_back(n);
  return (int)s->textures.size() - 1;
}
int orc_add_texture_uvdebug(orc_scene* s) {
  CHECK_SCENE(s);
  s->textures.push_back(std::make_shared<UVDebug>());
  return (int)s->textures.size() - 1;
}
int orc_add_texture_image(orc_scene* s, const uint8_t* rgb8, uint32_t width, uint32_t height) {
  CHECK_SCENE(s);
  if (!rgb8 || width == 0 || height == 0) return fail(RTW_ERR_INVALID, "image: empty");
  auto t = std::make_shared<ImageTexture>();
  t->rgb.assign(rgb8, rgb8 + (size_t)width * height * 3);
  t->width = width;
  t->height = height;
  s->textures.push_back(t);
  return (int)s->textures.size() - 1;
}

static int push_material(orc_scene* s, std::shared_ptr<Material> m) {
  m->id = (int)s->materials.size();
  s->materials.push_back(m);
  return m->id;
}
int orc_add_material_lambertian(orc_scene* s, int tex) {
  CHECK_SCENE(s);
  if (!valid_tex(s, tex)) return fail(RTW_ERR_INVALID, "lambertian: bad texture id");
  return push_material(s, std::make_shared<Lambertian>(s->textures[tex]));
}
int orc_add_material_metal(orc_scene* s, float r, float g, float b, float fuzz) {
  CHECK_SCENE(s);
  if (!(fuzz <= 1.0f)) return fail(RTW_ERR_INVALID, "metal: fuzz must be <= 1 (material.rs:71)");
  return push_material(s, std::make_shared<Metal>(Color(r, g, b), fuzz));
}
int orc_add_material_dielectric(orc_scene* s, float ir) {
  CHECK_SCENE(s);
  return push_material(s, std::make_shared<Dielectric>(ir));
}
int orc_add_material_diffuse_light(orc_scene* s, int tex) {
  CHECK_SCENE(s);
  if (!valid_tex(s, tex)) return fail(RTW_ERR_INVALID, "diffuse_light: bad texture id");
  return push_material(s, std::make_shared<DiffuseLight>(s->textures[tex]));
}

int orc_push_translation(orc_scene* s, const float offset[3]) {
  CHECK_SCENE(s);
  if (!offset) return fail(RTW_ERR_INVALID, "offset is NULL");
  Desc* d = push_child(s, D_TRANSLATE);
  d->f[0] = offset[0]; d->f[1] = offset[1]; d->f[2] = offset[2];
  s->stack.push_back(d);
  return RTW_OK;
}
int orc_push_rotation_y(orc_scene* s, float angle_degrees) {
  CHECK_SCENE(s);
  Desc* d = push_child(s, D_ROTY);
  d->f[0] = angle_degrees;
  s->stack.push_back(d);
  return RTW_OK;
}
int orc_push_rotation_y_sincos(orc_scene* s, float sin_theta, float cos_theta) {
  CHECK_SCENE(s);
  Desc* d = push_child(s, D_ROTY);
  d->f[0] = 0.f;
  d->f[1] = sin_theta; d->f[2] = cos_theta; d->f[3] = 1.f;  // f[3] != 0: take (sin, cos) as given
  s->stack.push_back(d);
  return RTW_OK;
}
int orc_pop_transform(orc_scene* s) {
  CHECK_SCENE(s);
  Desc* top = s->stack.back();
  if (top->kind != D_TRANSLATE && top->kind != D_ROTY) return fail(RTW_ERR_STATE, "pop_transform: no open transform");
  if (top->children.empty()) return fail(RTW_ERR_INVALID, "pop_transform: empty instance");
  s->stack.pop_back();
  return RTW_OK;
}
int orc_begin_group(orc_scene* s) {
  CHECK_SCENE(s);
  s->stack.push_back(push_child(s, D_GROUP));
  return RTW_OK;
}
int orc_end_group(orc_scene* s) {
  CHECK_SCENE(s);
  if (s->stack.back()->kind != D_GROUP) return fail(RTW_ERR_STATE, "end_group: no open group");
  s->stack.pop_back();
  return RTW_OK;
}

// true while the boundary of a medium is being emitted: its primitives get no canonical id
static bool in_medium(orc_scene* s) {
  for (Desc* d : s->stack)
    if (d->kind == D_MEDIUM) return true;
  return false;
}

int orc_begin_medium(orc_scene* s, float density, int texture) {
  CHECK_SCENE(s);
  if (!valid_tex(s, texture)) return fail(RTW_ERR_INVALID, "medium: bad texture id");
  if (in_medium(s)) return fail(RTW_ERR_UNSUPPORTED, "medium: nested media are not supported");
  for (Desc* d : s->stack)
    if (d->kind == D_TRANSLATE || d->kind == D_ROTY) return fail(RTW_ERR_UNSUPPORTED, "medium: a transformed medium is not supported (transform the boundary)");
  auto iso = std::make_shared<Isotropic>(s->textures[texture]);
  iso->id = (int)s->materials.size();
  s->materials.push_back(iso);
  Desc* d = push_child(s, D_MEDIUM);
  d->f[0] = density;
  d->material = iso->id;
  d->first_prim = s->num_prims;
  s->num_prims += 1;
  s->stack.push_back(d);
  return d->first_prim;
}
int orc_end_medium(orc_scene* s) {
  CHECK_SCENE(s);
  Desc* top = s->stack.back();
  if (top->kind != D_MEDIUM) return fail(RTW_ERR_STATE, "end_medium: no open medium");
  if (top->children.size() != 1) return fail(RTW_ERR_INVALID, "end_medium: a medium needs exactly one boundary object");
  s->stack.pop_back();
  return RTW_OK;
}

int orc_add_sphere(orc_scene* s, const float c[3], float radius, int material) {
  CHECK_SCENE(s);
  if (!c || !valid_mat(s, material)) return fail(RTW_ERR_INVALID, "sphere: bad argument");
  Desc* d = push_child(s, D_SPHERE);
  d->f[0] = c[0]; d->f[1] = c[1]; d->f[2] = c[2]; d->f[3] = radius;
  d->material = material;
  if (in_medium(s)) { d->first_prim = -1; return s->num_prims - 1; }
  d->first_prim = s->num_prims;
  s->num_prims += 1;
  return d->first_prim;
}
int orc_add_moving_sphere(orc_scene* s, const float c0[3], float time0, const float c1[3], float time1, float radius,
                          int material) {
  CHECK_SCENE(s);
  if (!c0 || !c1 || !valid_mat(s, material)) return fail(RTW_ERR_INVALID, "moving_sphere: bad argument");
  Desc* d = push_child(s, D_MSPHERE);
  d->f[0] = c0[0]; d->f[1] = c0[1]; d->f[2] = c0[2]; d->f[3] = time0;
  d->f[4] = c1[0]; d->f[5] = c1[1]; d->f[6] = c1[2]; d->f[7] = time1;
  d->f[8] = radius;
  d->material = material;
  d->first_prim = s->num_prims;
  s->num_prims += 1;
  return d->first_prim;
}
static int add_rect(orc_scene* s, int axis, float a0, float a1, float b0, float b1, float k, int material) {
  CHECK_SCENE(s);
  if (!valid_mat(s, material)) return fail(RTW_ERR_INVALID, "rect: bad material id");
  Desc* d = push_child(s, D_RECT);
  d->axis = axis;
  d->f[0] = a0; d->f[1] = a1; d->f[2] = b0; d->f[3] = b1; d->f[4] = k;
  d->material = material;
  d->first_prim = s->num_prims;
  s->num_prims += 1;
  return d->first_prim;
}
int orc_add_xy_rect(orc_scene* s, float x0, float x1, float y0, float y1, float k, int m) { return add_rect(s, 2, x0, x1, y0, y1, k, m); }
int orc_add_xz_rect(orc_scene* s, float x0, float x1, float z0, float z1, float k, int m) { return add_rect(s, 1, x0, x1, z0, z1, k, m); }
int orc_add_yz_rect(orc_scene* s, float y0, float y1, float z0, float z1, float k, int m) { return add_rect(s, 0, y0, y1, z0, z1, k, m); }
int orc_add_cuboid(orc_scene* s, const float p0[3], const float p1[3], int material) {
  CHECK_SCENE(s);
  if (!p0 || !p1 || !valid_mat(s, material)) return fail(RTW_ERR_INVALID, "cuboid: bad argument");
  Desc* d = push_child(s, D_CUBOID);
  for (int a = 0; a < 3; ++a) { d->f[a] = p0[a]; d->f[3 + a] = p1[a]; }
  d->material = material;
  if (in_medium(s)) { d->first_prim = -7; return s->num_prims - 1; }
  d->first_prim = s->num_prims;
  s->num_prims += 6;
  return d->first_prim;
}
int orc_add_triangles(orc_scene* s, uint32_t n, const float* vertices, const float* normals, const float* uvs,
                      const int32_t* material_ids, int material) {
  CHECK_SCENE(s);
  if (n == 0) return s->num_prims;
  if (!vertices) return fail(RTW_ERR_INVALID, "triangles: vertices is NULL");
  if (material_ids) {
    for (uint32_t i = 0; i < n; ++i)
      if (!valid_mat(s, material_ids[i])) return fail(RTW_ERR_INVALID, "triangles: bad material id");
  } else if (!valid_mat(s, material)) {
    return fail(RTW_ERR_INVALID, "triangles: bad material id");
  }
  Desc* d = push_child(s, D_TRIS);
  d->ntris = n;
  d->verts.assign(vertices, vertices + (size_t)n * 9);
  if (normals) d->normals.assign(normals, normals + (size_t)n * 9);
  if (uvs) d->uvs.assign(uvs, uvs + (size_t)n * 6);
  if (material_ids) d->mats.assign(material_ids, material_ids + n);
  d->material = material;
  d->first_prim = s->num_prims;
  s->num_prims += (int)n;
  return d->first_prim;
}

int orc_build(orc_scene* s, float time0, float time1, rtw_build_stats* stats) {
  CHECK_SCENE(s);
  if (s->stack.size() != 1) return fail(RTW_ERR_STATE, "build: unbalanced push/begin");
  if (s->num_prims == 0) return fail(RTW_ERR_INVALID, "build: empty scene");
  s->time0 = time0;
  s->time1 = time1;
  auto t0 = std::chrono::steady_clock::now();
  Rng build_rng(0x0b5eed5ull, 0, 0, 0xB0);  // host-side stream for BvhNode::new's axis draws
  for (const auto& c : s->root.children) instantiate(s, *c, false, build_rng, s->world_flat.objects);
  for (const auto& c : s->root.children) instantiate(s, *c, true, build_rng, s->world_ref.objects);
  auto t1 = std::chrono::steady_clock::now();
  s->built = true;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->num_prims = (uint32_t)s->num_prims;
    stats->ms_build = std::chrono::duration<float, std::milli>(t1 - t0).count();
  }
  return RTW_OK;
}
int orc_scene_num_prims(const orc_scene* s) { return s ? s->num_prims : RTW_ERR_INVALID; }
int orc_scene_num_nodes(const orc_scene*) { return 0; }

// mode 0 (RTW_TRACE_BVH slot): the flat canonical-order list = ground truth.
// mode 1: same.  mode 2: the reference's own structure (BvhNode for groups).
int orc_trace_closest(orc_scene* s, const rtw_ray* rays, uint64_t n, rtw_hit* hits, int mode) {
  if (!s || !s->built) return fail(RTW_ERR_STATE, "trace: scene not built");
  if (n && (!rays || !hits)) return fail(RTW_ERR_INVALID, "trace: NULL buffer");
  const HittableList& world = world_of(s, mode == 2 ? 1 : 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const rtw_ray& rr = rays[i];
    Ray r(Point3(rr.origin[0], rr.origin[1], rr.origin[2]), Vec3(rr.direction[0], rr.direction[1], rr.direction[2]),
          rr.time);
    Rng rng;
    HitRecord rec;
    bool h = world.hit(r, rr.t_min, rr.t_max, rng, rec);
    fill_hit(rec, h, &hits[i]);
  }
  return RTW_OK;
}

// Closest hit through a GPU-built LBVH (rtw_get_bvh) with the oracle's primitive tests: walks
// every pair whose box the ray's slab test (aabb.rs:23-48, non-strict reject) does not cull, and
// counts pair fetches and primitive tests — the n̄_node / n̄_prim of SURVEY.md §8(d).  Leaves are
// resolved through `leaf_objects[slot]` = the flat-world primitive with that slot's canonical id
// wrapped in its instance chain, which the caller obtains via orc_trace_single below.
// counts[0] += pair fetches, counts[1] += primitive tests.

int orc_render_ex(orc_scene* s, const rtw_camera* cam_in, const rtw_render_params* p, float* accum_rgb,
                  rtw_render_stats* stats, int mode /*0 flat, 2 reference structure*/,
                  int integrator /*0 iterative, 1 recursive*/, int threads /*0 = all*/) {
  if (!s || !s->built) return fail(RTW_ERR_STATE, "render: scene not built");
  if (!cam_in || !p || !accum_rgb) return fail(RTW_ERR_INVALID, "render: NULL argument");
  if (p->width < 2 || p->height < 2) return fail(RTW_ERR_INVALID, "render: width and height must be >= 2");
  const HittableList& world = world_of(s, mode == 2 ? 1 : 0);
  Camera cam = to_camera(cam_in);
  const uint32_t w = p->width, h = p->height;
  const uint32_t max_depth = p->max_depth ? p->max_depth : 50;
  uint32_t s0 = p->sample_begin, s1 = p->sample_end;
  if (s0 == 0 && s1 == 0) s1 = p->spp;
  if (s1 < s0) return fail(RTW_ERR_INVALID, "render: sample_end < sample_begin");
  const uint32_t nsamp = s1 - s0;
  uint32_t slices = p->slices ? p->slices : 1;
  if (slices > nsamp && nsamp > 0) slices = nsamp;
  const Color background(p->background[0], p->background[1], p->background[2]);
  uint64_t segments = 0, paths = 0;
  if (threads > 0) omp_set_num_threads(threads);
  auto t0 = std::chrono::steady_clock::now();
  // lib.rs:58: (0..h).rev() x (0..w); output index = yielded order
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : segments, paths)
  for (int64_t idx = 0; idx < (int64_t)w * h; ++idx) {
    uint32_t y_top = (uint32_t)(idx / w);
    uint32_t col = (uint32_t)(idx % w);
    uint32_t row = h - 1 - y_top;  // Pixel.row, bottom-up
    Color pixel_color(0.f, 0.f, 0.f);
    if (pixel_in_part(p, col, y_top)) {
      uint32_t pixel_index = row * w + col;
      for (uint32_t k = 0; k < slices; ++k) {
        uint32_t b = s0 + (uint32_t)(((uint64_t)nsamp * k) / slices);
        uint32_t e = s0 + (uint32_t)(((uint64_t)nsamp * (k + 1)) / slices);
        Color slice_color(0.f, 0.f, 0.f);
        for (uint32_t smp = b; smp < e; ++smp) {  // lib.rs:83-88
          Rng rng(p->seed, pixel_index, smp, 0);
          Ray r = camera_ray(cam, w, h, row, col, rng);
          ++paths;
          Color c = integrator == 1
                        ? sample_ray_recursive(world, background, r, rng, max_depth, 0, segments)
                        : sample_ray_iterative(world, background, r, rng, max_depth, segments);
          slice_color = slice_color + c;
        }
        pixel_color = (slices == 1) ? slice_color : pixel_color + slice_color;
      }
    }
    accum_rgb[3 * idx + 0] = pixel_color[0];
    accum_rgb[3 * idx + 1] = pixel_color[1];
    accum_rgb[3 * idx + 2] = pixel_color[2];
  }
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->segments = segments;
    stats->paths = paths;
    stats->slices = slices;
    stats->ms_render = std::chrono::duration<float, std::milli>(t1 - t0).count();
  }
  return RTW_OK;
}

// same signature as rtw_render: flat world, iterative integrator, all host threads
int orc_render(orc_scene* s, const rtw_camera* cam, const rtw_render_params* p, float* accum_rgb, rtw_render_stats* stats) {
  return orc_render_ex(s, cam, p, accum_rgb, stats, 0, 0, 0);
}

// same signature as rtw_render_frames: frame i = cameras[i], seed + i; frames and callbacks strictly in order
// (the checker has nothing to overlap)
int orc_render_frames(orc_scene* s, const rtw_camera* cameras, uint32_t n_frames, const rtw_render_params* params,
                      rtw_frame_callback on_frame, void* user) {
  if (!s || !s->built) return fail(RTW_ERR_STATE, "render_frames: scene not built");
  if (!params || (n_frames && !cameras)) return fail(RTW_ERR_INVALID, "render_frames: NULL argument");
  std::vector<float> accum((size_t)params->width * params->height * 3);
  int delivered = 0;
  for (uint32_t f = 0; f < n_frames; ++f) {
    rtw_render_params p = *params;
    p.seed = params->seed + f;
    rtw_render_stats st;
    int rc = orc_render(s, &cameras[f], &p, accum.data(), &st);
    if (rc != RTW_OK) return rc;
    delivered++;
    if (on_frame && on_frame(user, f, accum.data(), &st) != 0) break;
  }
  return delivered;
}

// Capture the ray batch that enters bounce `bounce` of sample `sample` of every pixel (bounce 0 =
// camera rays).  Paths that ended earlier yield a null ray (t_max < t_min, never hits).
// rays: width*height entries in the reference's pixel order.
int orc_capture_rays(orc_scene* s, const rtw_camera* cam_in, uint32_t w, uint32_t h, uint64_t seed, uint32_t sample,
                     uint32_t bounce, rtw_ray* rays) {
  if (!s || !s->built) return fail(RTW_ERR_STATE, "capture: scene not built");
  if (!cam_in || !rays || w < 2 || h < 2) return fail(RTW_ERR_INVALID, "capture: bad argument");
  const HittableList& world = s->world_flat;
  Camera cam = to_camera(cam_in);
  const float INF = std::numeric_limits<float>::infinity();
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t idx = 0; idx < (int64_t)w * h; ++idx) {
    uint32_t y_top = (uint32_t)(idx / w), col = (uint32_t)(idx % w), row = h - 1 - y_top;
    Rng rng(seed, row * w + col, sample, 0);
    Ray r = camera_ray(cam, w, h, row, col, rng);
    bool alive = true;
    for (uint32_t b = 0; b < bounce && alive; ++b) {
      rng.set_stage(b + 1);
      HitRecord rec;
      if (!world.hit(r, 0.001f, INF, rng, rec)) { alive = false; break; }
      Scatter sc;
      if (!rec.material->scatter(r, rec, rng, sc)) { alive = false; break; }
      r = sc.scattered_ray;
    }
    rtw_ray& o = rays[idx];
    for (int a = 0; a < 3; ++a) { o.origin[a] = r.origin[a]; o.direction[a] = r.direction[a]; }
    o.time = r.time;
    if (alive) { o.t_min = 0.001f; o.t_max = INF; }
    else { o.t_min = 1.0f; o.t_max = -1.0f; }
  }
  return RTW_OK;
}

// main.rs:73-86
int orc_resolve_rgb8(orc_scene*, const float* accum_rgb, uint32_t width, uint32_t height, uint32_t spp, uint8_t* rgb8) {
  if (!accum_rgb || !rgb8) return fail(RTW_ERR_INVALID, "resolve: NULL buffer");
  float scale = 1.0f / (float)spp;
  for (size_t i = 0; i < (size_t)width * height * 3; ++i) {
    float c = std::sqrt(scale * accum_rgb[i]);
    float cl = c;  // f32::clamp(0.0, 0.999); NaN stays NaN -> `as u8` = 0
    if (cl < 0.0f) cl = 0.0f;
    if (cl > 0.999f) cl = 0.999f;
    float v = 255.999f * cl;
    uint8_t b;
    if (!(v > 0.0f)) b = 0;
    else if (v >= 255.0f) b = 255;
    else b = (uint8_t)v;
    rgb8[i] = b;
  }
  return RTW_OK;
}

// ---- helpers for the tests -------------------------------------------------------------------
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

// n draws of the stream (seed, pixel, sample, stage): kind 0 = u32 bits (as float bit pattern is
// NOT used; written to out_u32), 1 = gen_f32, 2 = gen_range(lo, hi)
int orc_rng_draws(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage, int kind, float lo, float hi,
                  uint32_t n, uint32_t* out_u32, float* out_f32) {
  Rng rng(seed, pixel, sample, stage);
  for (uint32_t i = 0; i < n; ++i) {
    if (kind == 0) out_u32[i] = rng.next_u32();
    else if (kind == 1) out_f32[i] = rng.gen_f32();
    else out_f32[i] = rng.gen_range(lo, hi);
  }
  return RTW_OK;
}

// texture / material probes: evaluate Texture::value and one Material::scatter on the host
int orc_texture_value(orc_scene* s, int tex, float u, float v, const float p[3], float out_rgb[3]) {
  if (!s || !valid_tex(s, tex)) return fail(RTW_ERR_INVALID, "texture_value: bad id");
  Color c = s->textures[tex]->value(Point2d{u, v}, Vec3(p[0], p[1], p[2]));
  out_rgb[0] = c[0]; out_rgb[1] = c[1]; out_rgb[2] = c[2];
  return RTW_OK;
}

// Aabb::hit probe
int orc_aabb_hit(const float bmin[3], const float bmax[3], const rtw_ray* r) {
  Aabb b(Point3(bmin[0], bmin[1], bmin[2]), Point3(bmax[0], bmax[1], bmax[2]));
  Ray ray(Point3(r->origin[0], r->origin[1], r->origin[2]), Vec3(r->direction[0], r->direction[1], r->direction[2]), r->time);
  return b.hit(ray, r->t_min, r->t_max) ? 1 : 0;
}

// Camera::new (camera.rs:25-64)
int orc_camera_new(const float look_from[3], const float look_at[3], const float up[3], float vfov, float aspect,
                   float aperture, float focus_dist, float time0, float time1, rtw_camera* out) {
  Camera c = Camera::make(Point3(look_from[0], look_from[1], look_from[2]), Point3(look_at[0], look_at[1], look_at[2]),
                          Vec3(up[0], up[1], up[2]), vfov, aspect, aperture, focus_dist, time0, time1);
  for (int a = 0; a < 3; ++a) {
    out->origin[a] = c.origin[a];
    out->lower_left_corner[a] = c.lower_left_corner[a];
    out->horizontal[a] = c.horizontal[a];
    out->vertical[a] = c.vertical[a];
    out->u[a] = c.u[a];
    out->v[a] = c.v[a];
    out->w[a] = c.w[a];
  }
  out->lens_radius = c.lens_radius;
  out->time0 = c.time0;
  out->time1 = c.time1;
  return RTW_OK;
}

// world bounding box of primitive `prim_id` in the flat world (through its wrappers), for the
// LBVH containment test.  Linear search; test sizes only.
int orc_counters_reset(void) {
  g_counters = Counters();
  return 0;
}

int orc_num_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
