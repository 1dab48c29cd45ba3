// OBJ ingest: the parsed form of a Wavefront OBJ file with the semantics of the reference's
// loader (triangular.rs:170-260 over the wavefront_obj 10.0.0 crate): triangles in file order, polygons fan
// triangulated, f64 coordinates cast to f32, per-face material name from `usemtl`.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rtwh {

struct ObjData {
  std::vector<float> v, n, uv;           // per triangle, file order: 9 / 9 / 6 floats (0 where a corner has none)
  std::vector<uint8_t> has_n, has_uv;    // per triangle: all three corners carry a normal / a texture coordinate
  std::vector<int32_t> face_mtl;         // per triangle: index into mtl_names, -1 = no usemtl in effect
  std::vector<std::string> mtl_names;
  bool all_normals = true, all_uvs = true;
  std::string mtllib;                    // the last mtllib statement
  size_t triangles() const { return has_n.size(); }
};

// Reference implementation: one pass, one thread, iostreams.  Kept as the checker of the fast path.
ObjData parse_obj_simple(const std::string& path);

// The product path: the file is read once, cut into one chunk per thread at line boundaries, and parsed in three
// parallel passes (count -> prefix sums -> vertices -> faces), numbers with std::from_chars (correctly rounded, like
// the f64 parse of the crate).  Produces exactly what parse_obj_simple produces.  threads <= 0: all hardware threads.
ObjData parse_obj_fast(const std::string& path, int threads = 0);

}  // namespace rtwh
