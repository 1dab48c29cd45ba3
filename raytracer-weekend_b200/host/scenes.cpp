// Scene generators: a C++ mirror of console_app/src/scenes.rs (same objects, same order, same
// constants) plus the bench variants that BASELINE.json's configs describe.
// `thread_rng()` of the reference is replaced by a seeded HostRng with the same draw order.
#include <cstring>
#include <functional>

#include "rtw_host.hpp"

namespace rtwh {

namespace {

const Color DEFAULT_BACKGROUND(0.7f, 0.8f, 1.00f);  // scenes.rs:862

TexturePtr ground_checker() {  // scenes.rs:64-68
  return std::make_shared<Checker>(SolidColor::new_rgb(0.2f, 0.3f, 0.1f), SolidColor::new_rgb(0.9f, 0.9f, 0.9f), 10.0f);
}

// the camera block every scene repeats
Camera make_cam(Point3 from, Point3 at, Vec3 up, float vfov, float aspect, float aperture, float focus) {
  return Camera(from, at, up, vfov, aspect, aperture, focus, 0.0f, 1.0f);
}

// scenes.rs:63-162
World jumpy_balls(float aspect, HostRng& rng) {
  World w;
  auto material_ground = std::make_shared<Lambertian>(ground_checker());
  auto lambertian = Lambertian::new_solid_color(Color(0.4f, 0.2f, 0.1f));
  auto glass = std::make_shared<Dielectric>(1.5f);
  auto metal = std::make_shared<Metal>(Color(0.7f, 0.6f, 0.5f), 0.0f);
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -1000.0f, 0.0f), 1000.0f, material_ground));
  w.objects.push_back(std::make_shared<Sphere>(Point3(-4.0f, 0.2f, 0.1f), 1.0f, lambertian));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 1.0f, 0.0f), 1.0f, glass));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 1.0f, 0.0f), -0.95f, glass));
  w.objects.push_back(std::make_shared<Sphere>(Point3(4.0f, 1.0f, 0.0f), 1.0f, metal));
  for (int ai = -11; ai < 11; ++ai)
    for (int bi = -11; bi < 11; ++bi) {
      float a = (float)ai, b = (float)bi;
      float cx = a + 0.9f * rng.gen_f32();
      float cz = b + 0.9f * rng.gen_f32();
      Point3 center(cx, 0.2f, cz);
      if ((center - Point3(4.0f, 0.2f, 0.0f)).length() <= 0.9f) continue;  // scenes.rs:109-111 (as written)
      MaterialPtr sphere_material;
      double choose_mat = rng.gen_f64();
      if (choose_mat < 0.8) {
        Color c1 = rng.random_vec();
        Color c2 = rng.random_vec();
        sphere_material = Lambertian::new_solid_color(c1 * c2);
      } else if (choose_mat < 0.95) {
        Color albedo = rng.random_min_max(0.5f, 1.0f);
        float fuzz = rng.gen_range(0.0f, 0.5f);
        sphere_material = std::make_shared<Metal>(albedo, fuzz);
      } else {
        sphere_material = std::make_shared<Dielectric>(1.5f);
      }
      Point3 center2 = center + Vec3(0.0f, rng.gen_range(0.0f, 0.5f), 0.0f);
      w.objects.push_back(std::make_shared<MovingSphere>(center, 0.0f, center2, 1.0f, 0.2f, sphere_material));
    }
  w.cameras.push_back(make_cam(Point3(13.0f, 2.0f, 3.0f), Point3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 20.0f, aspect, 0.1f, 10.0f));
  w.background = DEFAULT_BACKGROUND;
  return w;
}

// scenes.rs:164-209
World two_spheres(float aspect, HostRng&) {
  World w;
  auto material_ground = std::make_shared<Lambertian>(ground_checker());
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -10.0f, 0.0f), 10.0f, material_ground));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 10.0f, 0.0f), 10.0f, material_ground));
  w.cameras.push_back(make_cam(Point3(13.0f, 2.0f, 3.0f), Point3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = DEFAULT_BACKGROUND;
  return w;
}

// scenes.rs:211-252
World two_perlin_spheres(float aspect, HostRng& rng) {
  World w;
  auto material_ground = std::make_shared<Lambertian>(std::make_shared<Noise>(Perlin(rng), 4.0f));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -1000.0f, 0.0f), 1000.0f, material_ground));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 2.0f, 0.0f), 2.0f, material_ground));
  w.cameras.push_back(make_cam(Point3(13.0f, 2.0f, 3.0f), Point3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = DEFAULT_BACKGROUND;
  return w;
}

// scenes.rs:254-288
World earth(float aspect, HostRng&) {
  World w;
  auto earth_surface = std::make_shared<Lambertian>(ImageTexture::open("models/earthmap.jpg"));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 0.0f, 0.0f), 2.0f, earth_surface));
  w.cameras.push_back(make_cam(Point3(13.0f, 2.0f, 3.0f), Point3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 20.0f, aspect, 0.0f, 10.0f));
  w.background = DEFAULT_BACKGROUND;
  return w;
}

// scenes.rs:290-348
World simple_light(float aspect, HostRng& rng) {
  World w;
  auto earth_surface = std::make_shared<DiffuseLight>(ImageTexture::open("models/earthmap.jpg"));
  auto material_ground = std::make_shared<Lambertian>(std::make_shared<Noise>(Perlin(rng), 4.0f));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -1000.0f, 0.0f), 1000.0f, material_ground));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 2.0f, 0.0f), 2.0f, material_ground));
  w.objects.push_back(std::make_shared<XYRectangle>(3.0f, 5.0f, 1.0f, 3.0f, -2.0f, earth_surface));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 6.0f, 0.0f), 2.0f, earth_surface));
  w.cameras.push_back(make_cam(Point3(26.0f, 3.0f, 6.0f), Point3(0.0f, 2.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 20.0f, aspect, 0.0f, 10.0f));
  w.background = Color(0.0f, 0.0f, 0.0f);
  return w;
}

// scenes.rs:350-414
World cornell_box(float aspect, HostRng&) {
  World w;
  auto red = Lambertian::new_solid_color(Color(0.65f, 0.05f, 0.05f));
  auto white = Lambertian::new_solid_color(Color(0.73f, 0.73f, 0.73f));
  auto green = Lambertian::new_solid_color(Color(0.12f, 0.45f, 0.15f));
  auto light = std::make_shared<DiffuseLight>(SolidColor::new_rgb(15.0f, 15.0f, 15.0f));
  HittablePtr box1 = translate(rotate_y(std::make_shared<Cuboid>(Point3(0.0f, 0.0f, 0.0f), Point3(165.0f, 330.0f, 165.0f), white), 15.0f),
                               Vec3(265.0f, 0.0f, 295.0f));
  HittablePtr box2 = translate(rotate_y(std::make_shared<Cuboid>(Point3(0.0f, 0.0f, 0.0f), Point3(165.0f, 165.0f, 165.0f), white), -18.0f),
                               Vec3(130.0f, 0.0f, 65.0f));
  w.objects.push_back(std::make_shared<YZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 555.0f, green));
  w.objects.push_back(std::make_shared<YZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 0.0f, red));
  w.objects.push_back(std::make_shared<XZRectangle>(213.0f, 343.0f, 227.0f, 332.0f, 554.0f, light));
  w.objects.push_back(std::make_shared<XZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 0.0f, white));
  w.objects.push_back(std::make_shared<XZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 555.0f, white));
  w.objects.push_back(std::make_shared<XYRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 555.0f, white));
  w.objects.push_back(box1);
  w.objects.push_back(box2);
  w.cameras.push_back(make_cam(Point3(278.0f, 278.0f, -800.0f), Point3(278.0f, 278.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = Color(0.0f, 0.0f, 0.0f);
  return w;
}

// scenes.rs:416-483
World smokey_cornell_box(float aspect, HostRng&) {
  World w;
  auto red = Lambertian::new_solid_color(Color(0.65f, 0.05f, 0.05f));
  auto white = Lambertian::new_solid_color(Color(0.73f, 0.73f, 0.73f));
  auto green = Lambertian::new_solid_color(Color(0.12f, 0.45f, 0.15f));
  auto light = std::make_shared<DiffuseLight>(SolidColor::new_rgb(7.0f, 7.0f, 7.0f));
  HittablePtr box1 = translate(rotate_y(std::make_shared<Cuboid>(Point3(0.0f, 0.0f, 0.0f), Point3(165.0f, 330.0f, 165.0f), white), 15.0f),
                               Vec3(265.0f, 0.0f, 295.0f));
  HittablePtr box2 = translate(rotate_y(std::make_shared<Cuboid>(Point3(0.0f, 0.0f, 0.0f), Point3(165.0f, 165.0f, 165.0f), white), -18.0f),
                               Vec3(130.0f, 0.0f, 65.0f));
  box1 = std::make_shared<ConstantMedium>(box1, 0.005f, SolidColor::new_rgb(0.0f, 0.0f, 0.0f));
  box2 = std::make_shared<ConstantMedium>(box2, 0.005f, SolidColor::new_rgb(1.0f, 1.0f, 1.0f));
  w.objects.push_back(std::make_shared<YZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 555.0f, green));
  w.objects.push_back(std::make_shared<YZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 0.0f, red));
  w.objects.push_back(std::make_shared<XZRectangle>(113.0f, 443.0f, 127.0f, 432.0f, 554.0f, light));
  w.objects.push_back(std::make_shared<XZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 0.0f, white));
  w.objects.push_back(std::make_shared<XZRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 555.0f, white));
  w.objects.push_back(std::make_shared<XYRectangle>(0.0f, 555.0f, 0.0f, 555.0f, 555.0f, white));
  w.objects.push_back(box1);
  w.objects.push_back(box2);
  w.cameras.push_back(make_cam(Point3(278.0f, 278.0f, -800.0f), Point3(278.0f, 278.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = Color(0.0f, 0.0f, 0.0f);
  return w;
}

// scenes.rs:485-620
World book2_final_scene(float aspect, HostRng& rng) {
  World w;
  HittableList boxes1;
  auto ground = Lambertian::new_solid_color(Color(0.48f, 0.83f, 0.53f));
  const int boxes_per_side = 20;
  for (int ii = 0; ii < boxes_per_side; ++ii)
    for (int jj = 0; jj < boxes_per_side; ++jj) {
      float i = (float)ii, j = (float)jj;
      float wd = 100.0f;
      float x0 = -1000.0f + i * wd, z0 = -1000.0f + j * wd, y0 = 0.0f;
      float x1 = x0 + wd, y1 = rng.gen_range(1.0f, 101.0f), z1 = z0 + wd;
      boxes1.push_back(std::make_shared<Cuboid>(Point3(x0, y0, z0), Point3(x1, y1, z1), ground));
    }
  w.objects.push_back(std::make_shared<BvhNode>(boxes1, 0.0f, 1.0f));
  w.objects.push_back(std::make_shared<XZRectangle>(123.0f, 423.0f, 147.0f, 412.0f, 554.0f,
                                                    std::make_shared<DiffuseLight>(SolidColor::new_rgb(7.0f, 7.0f, 7.0f))));
  Point3 center1(400.0f, 400.0f, 200.0f);
  Point3 center2 = center1 + Vec3(30.0f, 0.0f, 0.0f);
  w.objects.push_back(std::make_shared<MovingSphere>(center1, 0.0f, center2, 1.0f, 50.0f, Lambertian::new_solid_color(Color(0.7f, 0.3f, 0.1f))));
  w.objects.push_back(std::make_shared<Sphere>(Point3(260.0f, 150.0f, 45.0f), 50.0f, std::make_shared<Dielectric>(1.5f)));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, 150.0f, 145.0f), 50.0f, std::make_shared<Metal>(Color(0.8f, 0.8f, 0.9f), 1.0f)));
  auto boundary = std::make_shared<Sphere>(Point3(360.0f, 150.0f, 145.0f), 70.0f, std::make_shared<Dielectric>(1.5f));
  w.objects.push_back(boundary);
  w.objects.push_back(std::make_shared<ConstantMedium>(boundary, 0.2f, SolidColor::new_rgb(0.2f, 0.4f, 0.9f)));
  auto fog = std::make_shared<Sphere>(Point3(0.0f, 0.0f, 0.0f), 5000.0f, std::make_shared<Dielectric>(1.5f));
  w.objects.push_back(std::make_shared<ConstantMedium>(fog, 0.0001f, SolidColor::new_rgb(1.0f, 1.0f, 1.0f)));
  w.objects.push_back(std::make_shared<Sphere>(Point3(400.0f, 200.0f, 400.0f), 100.0f,
                                               std::make_shared<Lambertian>(ImageTexture::open("models/earthmap.jpg"))));
  auto pertext = std::make_shared<Noise>(Perlin(rng), 0.1f);
  w.objects.push_back(std::make_shared<Sphere>(Point3(220.0f, 280.0f, 300.0f), 80.0f, std::make_shared<Lambertian>(pertext)));
  HittableList boxes2;
  auto white = std::make_shared<Lambertian>(SolidColor::new_rgb(0.73f, 0.73f, 0.73f));
  for (int k = 0; k < 1000; ++k) boxes2.push_back(std::make_shared<Sphere>(rng.random_min_max(0.0f, 165.0f), 10.0f, white));
  w.objects.push_back(std::make_shared<Translation>(std::make_shared<YRotation>(std::make_shared<BvhNode>(boxes2, 0.0f, 1.0f), 15.0f),
                                                    Vec3(-100.0f, 270.0f, 395.0f)));
  Point3 look_from(478.0f, 278.0f, -600.0f), look_at(278.0f, 278.0f, 0.0f);
  w.cameras.push_back(make_cam(look_from, look_at, Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, (look_at - look_from).length()));
  w.background = Color(0.0f, 0.0f, 0.0f);
  return w;
}

// scenes.rs:622-667: 30 cameras around the Book-2 scene, the world wrapped in one BvhNode
World animated_book2_final(float aspect, HostRng& rng) {
  World base = book2_final_scene(aspect, rng);
  World w;
  Point3 look_at(278.0f, 278.0f, 278.0f);
  const float len_s = 3.0f, fps = 10.0f;
  const float frames = fps * len_s;
  for (int frame = 0; frame < (int)frames; ++frame) {
    float from_x = 478.0f - (float)frame * (2.0f * 478.0f) / frames;
    Point3 look_from(from_x, 278.0f, -600.0f);
    w.cameras.push_back(Camera(look_from, look_at, Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 1.0f, (look_at - look_from).length(), 0.0f, 1.0f));
  }
  w.objects.push_back(std::make_shared<BvhNode>(base.objects, 0.0f, 1.0f));
  w.background = base.background;
  return w;
}

// scenes.rs:669-717
World simple_triangle(float aspect, HostRng&) {
  World w;
  auto material_ground = std::make_shared<Lambertian>(ground_checker());
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -10.0f, 0.0f), 10.0f, material_ground));
  const Point3 tri[3] = {Point3(-5.0f, 0.0f, 5.0f), Point3(0.0f, 7.0f, 0.0f), Point3(5.0f, 0.0f, -5.0f)};
  w.objects.push_back(TriangleMesh::new_flat_shaded(tri, std::make_shared<Lambertian>(std::make_shared<UVDebug>())));
  w.cameras.push_back(make_cam(Point3(13.0f, 2.0f, 3.0f), Point3(0.0f, 2.5f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = DEFAULT_BACKGROUND;
  return w;
}

// scenes.rs:719-771.  variant 0 = as the reference (cow without material -> magenta emitter);
// variant 1 = BASELINE config C3 "Lambertian + metal": even faces Lambertian(0.73), odd faces
// Metal((0.8,0.8,0.9), 0.1)  (SURVEY.md §8d).
World wavefront_cow(float aspect, int variant) {
  World w;
  auto material_ground = std::make_shared<Lambertian>(ground_checker());
  HittablePtr cow;
  if (variant == 0) {
    cow = load_wavefront_obj("models/cow-nonormals.obj");
  } else {
    auto grey = Lambertian::new_solid_color(Color(0.73f, 0.73f, 0.73f));
    auto metal = std::make_shared<Metal>(Color(0.8f, 0.8f, 0.9f), 0.1f);
    auto mesh = load_mesh("models/cow-nonormals.obj", grey);
    std::vector<MaterialPtr> per_face(mesh->len());
    for (size_t i = 0; i < per_face.size(); ++i) per_face[i] = (i % 2 == 0) ? MaterialPtr(grey) : MaterialPtr(metal);
    cow = std::make_shared<BvhNode>(HittableList{mesh->with_materials(std::move(per_face))}, 0.0f, 1.0f);
  }
  cow = std::make_shared<Translation>(cow, Vec3(0.0f, 2.5f, 0.0f));
  w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -10.6f, 0.0f), 10.0f, material_ground));
  w.objects.push_back(std::make_shared<XYRectangle>(1.0f, 5.0f, 1.0f, 7.0f, 5.0f,
                                                    std::make_shared<DiffuseLight>(SolidColor::new_rgb(1.4f, 1.3f, 1.3f))));
  w.objects.push_back(cow);
  w.cameras.push_back(make_cam(Point3(13.0f, 2.0f, 3.0f), Point3(0.0f, 2.5f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = Color(0.085f, 0.1f, 0.125f);
  return w;
}

// Deterministic stand-in for the monument's diffuse PNG, which is missing from the reference tree
// (.MISSING_LARGE_BLOBS:1): 2048x2048 value noise + uv grid.
std::shared_ptr<ImageTexture> substitute_monument_texture() {
  const uint32_t N = 2048;
  std::vector<uint8_t> rgb((size_t)N * N * 3);
  auto hash = [](uint32_t x, uint32_t y) {
    uint32_t h = x * 0x9E3779B1u ^ (y + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0xC2B2AE3Du; h ^= h >> 13;
    return (float)(h & 0xFFFF) * (1.0f / 65535.0f);
  };
  for (uint32_t y = 0; y < N; ++y)
    for (uint32_t x = 0; x < N; ++x) {
      float fx = (float)x / 64.0f, fy = (float)y / 64.0f;
      uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy;
      float tx = fx - (float)ix, ty = fy - (float)iy;
      float a = hash(ix, iy), b = hash(ix + 1, iy), c = hash(ix, iy + 1), d = hash(ix + 1, iy + 1);
      float n = (a * (1 - tx) + b * tx) * (1 - ty) + (c * (1 - tx) + d * tx) * ty;
      bool grid = (x % 256) < 4 || (y % 256) < 4;
      float r = 0.45f + 0.35f * n, g = 0.40f + 0.30f * n, bl = 0.33f + 0.25f * n;
      if (grid) { r *= 0.5f; g *= 0.5f; bl *= 0.6f; }
      uint8_t* px = &rgb[((size_t)y * N + x) * 3];
      px[0] = (uint8_t)(255.0f * r); px[1] = (uint8_t)(255.0f * g); px[2] = (uint8_t)(255.0f * bl);
    }
  return std::make_shared<ImageTexture>(std::move(rgb), N, N);
}

// scenes.rs:816-858.  variant 0 = as the reference (needs the missing PNG -> throws like the
// reference panics); variant 1 = BASELINE config C4: substitute texture + earthmap sphere
// (r=2, Lambertian<ImageTexture(earthmap)>) at (0,-16,2)  (SURVEY.md §8d).
World textured_monument(float aspect, int variant) {
  World w;
  HittablePtr monument;
  if (variant == 0) {
    monument = load_wavefront_obj("models/monument_downscaled_polygon_reduced.obj");
  } else {
    monument = load_wavefront_obj("models/monument_downscaled_polygon_reduced.obj",
                                  std::make_shared<Lambertian>(substitute_monument_texture()));
  }
  monument = std::make_shared<Translation>(monument, Vec3(0.0f, 0.0f, -19.0f));
  w.objects.push_back(std::make_shared<XYRectangle>(-15.0f, 15.0f, -17.0f, 17.0f, 33.0f,
                                                    std::make_shared<DiffuseLight>(SolidColor::new_rgb(1.2f, 1.0f, 1.0f))));
  w.objects.push_back(monument);
  if (variant == 1)
    w.objects.push_back(std::make_shared<Sphere>(Point3(0.0f, -16.0f, 2.0f), 2.0f,
                                                 std::make_shared<Lambertian>(ImageTexture::open("models/earthmap.jpg"))));
  w.cameras.push_back(make_cam(Point3(-5.0f, -30.0f, 25.0f), Point3(0.0f, 0.0f, 5.0f), Vec3(1.0f, 0.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = Color(0.085f, 0.1f, 0.125f);
  return w;
}

// BASELINE config C5: n_spheres spheres (centres U[-100,100]^3, radius U[0.05,0.25],
// material class by id%10: 0-6 Lambertian, 7-8 Metal(fuzz U[0,0.5]), 9 Dielectric(1.5); albedos from
// a fixed-size random palette) and n_shards x 10 triangles (vertices = shard centre + U[-0.3,0.3]^3,
// Lambertian from the palette); sky background; camera at (0,0,-260) looking at the origin, vfov 40.
World stress(float aspect, HostRng& rng, size_t n_spheres, size_t n_shards) {
  World w;
  std::vector<MaterialPtr> lambert, metal;
  for (int i = 0; i < 70; ++i) { Color a = rng.random_vec(); Color b = rng.random_vec(); lambert.push_back(Lambertian::new_solid_color(a * b + Color(0.05f, 0.05f, 0.05f))); }
  for (int i = 0; i < 20; ++i) { Color a = rng.random_min_max(0.5f, 1.0f); float fz = rng.gen_range(0.0f, 0.5f); metal.push_back(std::make_shared<Metal>(a, fz)); }
  MaterialPtr glass = std::make_shared<Dielectric>(1.5f);
  HittableList spheres;
  spheres.reserve(n_spheres);
  for (size_t i = 0; i < n_spheres; ++i) {
    Point3 c = rng.random_min_max(-100.0f, 100.0f);
    float r = rng.gen_range(0.05f, 0.25f);
    size_t cls = i % 10;
    MaterialPtr m = cls < 7 ? lambert[i % 70] : (cls < 9 ? metal[i % 20] : glass);
    spheres.push_back(std::make_shared<Sphere>(c, r, m));
  }
  if (!spheres.empty()) w.objects.push_back(std::make_shared<BvhNode>(std::move(spheres), 0.0f, 1.0f));
  if (n_shards > 0) {
    size_t ntri = n_shards * 10;
    std::vector<float> verts(ntri * 9);
    std::vector<MaterialPtr> per_face(ntri);
    for (size_t s = 0; s < n_shards; ++s) {
      Point3 c = rng.random_min_max(-100.0f, 100.0f);
      for (size_t t = 0; t < 10; ++t) {
        size_t i = s * 10 + t;
        for (int k = 0; k < 3; ++k) {
          Vec3 d = rng.random_min_max(-0.3f, 0.3f);
          Point3 p = c + d;
          verts[9 * i + 3 * k] = p.x(); verts[9 * i + 3 * k + 1] = p.y(); verts[9 * i + 3 * k + 2] = p.z();
        }
        per_face[i] = lambert[s % 70];
      }
    }
    auto mesh = std::make_shared<TriangleMesh>(std::move(verts), std::vector<float>(), std::vector<float>(), std::move(per_face));
    w.objects.push_back(std::make_shared<BvhNode>(HittableList{mesh}, 0.0f, 1.0f));
  }
  w.cameras.push_back(make_cam(Point3(0.0f, 0.0f, -260.0f), Point3(0.0f, 0.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), 40.0f, aspect, 0.0f, 10.0f));
  w.background = DEFAULT_BACKGROUND;
  return w;
}

}  // namespace

std::vector<std::string> scene_names() {
  return {"jumpy-balls", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell-box", "smokey-cornell-box",
          "book2-final-scene", "animated-book2-final-scene", "simple-triangle", "wavefront-cow-obj",
          "wavefront-suspension-obj", "textured-monument",
          // bench variants (BASELINE.json configs C3-C5)
          "cow-lambert-metal", "monument-earth", "stress", "stress-spheres", "stress-triangles"};
}

// name[:a[:b]] — the numeric suffixes size the stress scenes (spheres, shards)
World generate_scene(const std::string& full_name, float aspect_ratio, uint64_t seed) {
  std::string name = full_name;
  std::vector<size_t> args;
  size_t colon = name.find(':');
  if (colon != std::string::npos) {
    std::string rest = name.substr(colon + 1);
    name = name.substr(0, colon);
    size_t pos = 0;
    while (pos <= rest.size()) {
      size_t nx = rest.find(':', pos);
      std::string tok = rest.substr(pos, nx == std::string::npos ? std::string::npos : nx - pos);
      if (!tok.empty()) args.push_back((size_t)std::stoull(tok));
      if (nx == std::string::npos) break;
      pos = nx + 1;
    }
  }
  HostRng rng(seed);
  if (name == "jumpy-balls") return jumpy_balls(aspect_ratio, rng);
  if (name == "two-spheres") return two_spheres(aspect_ratio, rng);
  if (name == "two-perlin-spheres") return two_perlin_spheres(aspect_ratio, rng);
  if (name == "earth") return earth(aspect_ratio, rng);
  if (name == "simple-light") return simple_light(aspect_ratio, rng);
  if (name == "cornell-box") return cornell_box(aspect_ratio, rng);
  if (name == "simple-triangle") return simple_triangle(aspect_ratio, rng);
  if (name == "wavefront-cow-obj") return wavefront_cow(aspect_ratio, 0);
  if (name == "cow-lambert-metal") return wavefront_cow(aspect_ratio, 1);
  if (name == "textured-monument") return textured_monument(aspect_ratio, 0);
  if (name == "monument-earth") return textured_monument(aspect_ratio, 1);
  if (name == "stress") return stress(aspect_ratio, rng, args.size() > 0 ? args[0] : 1000000, args.size() > 1 ? args[1] : 1000000);
  if (name == "stress-spheres") return stress(aspect_ratio, rng, args.size() > 0 ? args[0] : 1000000, 0);
  if (name == "stress-triangles") return stress(aspect_ratio, rng, 0, args.size() > 0 ? args[0] : 1000000);
  if (name == "smokey-cornell-box") return smokey_cornell_box(aspect_ratio, rng);
  if (name == "book2-final-scene") return book2_final_scene(aspect_ratio, rng);
  if (name == "animated-book2-final-scene") return animated_book2_final(aspect_ratio, rng);
  if (name == "wavefront-suspension-obj")
    throw Error("scene 'wavefront-suspension-obj': Normals_Try3.obj uses usemtl without mtllib; the reference panics on it "
                "(triangular.rs:177-179)");
  throw Error("unknown scene '" + full_name + "'");
}

}  // namespace rtwh
