// Parallel OBJ parser: see obj_loader.hpp.  Semantics = parse_obj_simple (rtw_host.cpp), which restates what the
// reference's loader sees through the wavefront_obj crate (triangular.rs:170-260).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>

#include "obj_loader.hpp"
#include "rtw_host.hpp"

namespace rtwh {

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

struct Cursor {
  const char* p;
  const char* e;  // end of line (exclusive)
  void skip() { while (p < e && is_space(*p)) ++p; }
  // next whitespace-delimited token; empty when the line is exhausted
  bool token(const char*& b, const char*& t) {
    skip();
    b = p;
    while (p < e && !is_space(*p)) ++p;
    t = p;
    return t > b;
  }
};

// `ss >> double`: a correctly rounded decimal parse; a missing / malformed field leaves 0 (and stops the line)
inline bool parse_double(Cursor& c, double& out) {
  const char *b, *t;
  out = 0.0;
  if (!c.token(b, t)) return false;
  if (*b == '+') ++b;
  auto r = std::from_chars(b, t, out);
  if (r.ec != std::errc() || r.ptr == b) { out = 0.0; return false; }
  return true;
}

inline bool parse_long(const char* b, const char* t, long& out) {
  if (b < t && *b == '+') ++b;
  auto r = std::from_chars(b, t, out);
  return r.ec == std::errc() && r.ptr != b;
}

enum Tag { T_OTHER, T_V, T_VT, T_VN, T_F, T_MTLLIB, T_USEMTL };

inline Tag classify(Cursor& c) {
  const char *b, *t;
  if (!c.token(b, t) || *b == '#') return T_OTHER;
  const size_t n = (size_t)(t - b);
  if (n == 1 && *b == 'v') return T_V;
  if (n == 1 && *b == 'f') return T_F;
  if (n == 2 && b[0] == 'v' && b[1] == 't') return T_VT;
  if (n == 2 && b[0] == 'v' && b[1] == 'n') return T_VN;
  if (n == 6 && !memcmp(b, "mtllib", 6)) return T_MTLLIB;
  if (n == 6 && !memcmp(b, "usemtl", 6)) return T_USEMTL;
  return T_OTHER;
}

struct Chunk {
  size_t begin = 0, end = 0;
  size_t nv = 0, nvt = 0, nvn = 0, ntri = 0;            // counted in pass A
  size_t v0 = 0, vt0 = 0, vn0 = 0, tri0 = 0;             // prefix sums
  std::vector<std::pair<size_t, std::string>> usemtl;   // (triangles of this chunk emitted before it, name)
  std::string mtllib;
  bool has_mtllib = false;
  std::string error;
};

template <class F>
void for_each_line(const char* data, size_t begin, size_t end, F&& f) {
  size_t p = begin;
  while (p < end) {
    const char* nl = (const char*)memchr(data + p, '\n', end - p);
    const size_t e = nl ? (size_t)(nl - data) : end;
    Cursor c{data + p, data + e};
    f(c);
    p = e + 1;
  }
}

}  // namespace

ObjData parse_obj_fast(const std::string& path, int threads) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw Error("cannot open OBJ file: " + path);
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)std::max(sz, 0L));
  if (sz > 0 && fread(buf.data(), 1, buf.size(), f) != buf.size()) {
    fclose(f);
    throw Error("cannot read OBJ file: " + path);
  }
  fclose(f);
  const char* data = buf.data();
  const size_t size = buf.size();

  // Vertex statements and face statements usually sit in different parts of the file, so every pass has its work
  // in a different subset of the chunks: many more chunks than threads, handed out dynamically.
  const int workers = std::max(1, threads > 0 ? threads : (int)std::thread::hardware_concurrency());
  int T = std::max(1, std::min(workers * 16, (int)(size / (1 << 16)) + 1));  // at least 64 KiB per chunk
  std::vector<Chunk> ch((size_t)T);
  for (int i = 0; i < T; ++i) {  // cut at line boundaries
    size_t b = size * (size_t)i / (size_t)T;
    if (i > 0) {
      const char* nl = (const char*)memchr(data + b - 1, '\n', size - (b - 1));
      b = nl ? (size_t)(nl - data) + 1 : size;
    }
    ch[(size_t)i].begin = b;
    if (i > 0) ch[(size_t)i - 1].end = b;
  }
  ch.back().end = size;
  for (int i = 1; i < T; ++i) ch[(size_t)i].begin = std::max(ch[(size_t)i].begin, ch[(size_t)i - 1].begin);

  auto parallel = [&](auto&& body) {
    std::atomic<int> next{0};
    auto work = [&] {
      for (int i = next.fetch_add(1); i < T; i = next.fetch_add(1)) body(ch[(size_t)i]);
    };
    std::vector<std::thread> th;
    for (int w = 1; w < std::min(workers, T); ++w) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    for (auto& c : ch)
      if (!c.error.empty()) throw Error(c.error);
  };

  // ---- pass A: count ------------------------------------------------------------------------------------
  parallel([&](Chunk& c) {
    for_each_line(data, c.begin, c.end, [&](Cursor& cur) {
      switch (classify(cur)) {
        case T_V: c.nv++; break;
        case T_VT: c.nvt++; break;
        case T_VN: c.nvn++; break;
        case T_F: {
          size_t corners = 0;
          const char *b, *t;
          while (cur.token(b, t)) corners++;
          if (corners < 3) {
            if (c.error.empty()) c.error = "OBJ points / lines are not supported (triangular.rs:186-191): " + path;
          } else {
            c.ntri += corners - 2;
          }
          break;
        }
        case T_MTLLIB: {
          const char *b, *t;
          c.mtllib = cur.token(b, t) ? std::string(b, t) : std::string();
          c.has_mtllib = true;
          break;
        }
        case T_USEMTL: {
          const char *b, *t;
          c.usemtl.emplace_back(c.ntri, cur.token(b, t) ? std::string(b, t) : std::string());
          break;
        }
        default: break;
      }
    });
  });
  size_t nv = 0, nvt = 0, nvn = 0, ntri = 0;
  ObjData out;
  std::map<std::string, int32_t> mtl_index;
  std::vector<int32_t> chunk_start_mtl((size_t)T, -1);   // material in effect when the chunk starts
  std::vector<std::vector<int32_t>> chunk_event_mtl((size_t)T);
  int32_t cur = -1;
  for (int i = 0; i < T; ++i) {
    Chunk& c = ch[(size_t)i];
    c.v0 = nv; c.vt0 = nvt; c.vn0 = nvn; c.tri0 = ntri;
    nv += c.nv; nvt += c.nvt; nvn += c.nvn; ntri += c.ntri;
    if (c.has_mtllib) out.mtllib = c.mtllib;
    chunk_start_mtl[(size_t)i] = cur;
    for (auto& ev : c.usemtl) {
      if (ev.second.empty()) {
        // `ss >> cur_mtl` with nothing to read fails before it touches the string: the material stays in effect
      } else {
        auto it = mtl_index.find(ev.second);
        if (it == mtl_index.end()) {
          it = mtl_index.emplace(ev.second, (int32_t)out.mtl_names.size()).first;
          out.mtl_names.push_back(ev.second);
        }
        cur = it->second;
      }
      chunk_event_mtl[(size_t)i].push_back(cur);
    }
  }

  // ---- pass B: vertices ---------------------------------------------------------------------------------
  std::vector<double> pos(3 * nv), tex(2 * nvt), nrm(3 * nvn);
  parallel([&](Chunk& c) {
    size_t iv = c.v0, it = c.vt0, in = c.vn0;
    for_each_line(data, c.begin, c.end, [&](Cursor& cur2) {
      switch (classify(cur2)) {
        case T_V: {
          double* d = &pos[3 * iv++];
          if (parse_double(cur2, d[0]) && parse_double(cur2, d[1])) parse_double(cur2, d[2]);
          break;
        }
        case T_VT: {
          double* d = &tex[2 * it++];
          if (parse_double(cur2, d[0])) parse_double(cur2, d[1]);
          break;
        }
        case T_VN: {
          double* d = &nrm[3 * in++];
          if (parse_double(cur2, d[0]) && parse_double(cur2, d[1])) parse_double(cur2, d[2]);
          break;
        }
        default: break;
      }
    });
  });

  // ---- pass C: faces ------------------------------------------------------------------------------------
  out.v.resize(9 * ntri);
  out.n.resize(9 * ntri);
  out.uv.resize(6 * ntri);
  out.has_n.resize(ntri);
  out.has_uv.resize(ntri);
  out.face_mtl.resize(ntri);
  std::vector<uint8_t> chunk_all_n((size_t)T, 1), chunk_all_uv((size_t)T, 1);
  parallel([&](Chunk& c) {
    const size_t ci = (size_t)(&c - ch.data());
    size_t sv = c.v0, st = c.vt0, sn = c.vn0;  // vertices defined so far: relative (negative) indices count from here
    size_t tri = c.tri0, ev = 0;
    int32_t mtl = chunk_start_mtl[ci];
    struct Corner { long v, t, n; };
    std::vector<Corner> cs;
    auto resolve = [&](long idx, size_t count, long& r) {
      r = idx > 0 ? idx - 1 : (long)count + idx;
      return idx != 0 && r >= 0 && r < (long)count;
    };
    for_each_line(data, c.begin, c.end, [&](Cursor& cur3) {
      switch (classify(cur3)) {
        case T_V: sv++; break;
        case T_VT: st++; break;
        case T_VN: sn++; break;
        case T_USEMTL: mtl = chunk_event_mtl[ci][ev++]; break;
        case T_F: {
          cs.clear();
          const char *b, *t;
          while (cur3.token(b, t)) {
            Corner k{-1, -1, -1};
            const char* s1 = (const char*)memchr(b, '/', (size_t)(t - b));
            const char* s2 = s1 ? (const char*)memchr(s1 + 1, '/', (size_t)(t - s1 - 1)) : nullptr;
            long iv = 0, it2 = 0, in2 = 0;
            bool ok = parse_long(b, s1 ? s1 : t, iv) && resolve(iv, sv, k.v);
            if (ok && s1) {
              const char* tb = s1 + 1;
              const char* te = s2 ? s2 : t;
              if (te > tb) ok = parse_long(tb, te, it2) && resolve(it2, st, k.t);
            }
            if (ok && s2 && t > s2 + 1) ok = parse_long(s2 + 1, t, in2) && resolve(in2, sn, k.n);
            if (!ok) {
              if (c.error.empty()) c.error = "OBJ index out of range in " + path;
              return;
            }
            cs.push_back(k);
          }
          for (size_t k = 2; k < cs.size(); ++k, ++tri) {  // triangle fan
            const Corner tr[3] = {cs[0], cs[k - 1], cs[k]};
            bool hn = true, ht = true;
            for (int q = 0; q < 3; ++q) {
              const Corner& co = tr[q];
              hn = hn && co.n >= 0;
              ht = ht && co.t >= 0;
              for (int a = 0; a < 3; ++a) out.v[9 * tri + 3 * q + a] = (float)pos[3 * (size_t)co.v + a];  // f64 -> `as f32`
              for (int a = 0; a < 3; ++a) out.n[9 * tri + 3 * q + a] = co.n >= 0 ? (float)nrm[3 * (size_t)co.n + a] : 0.f;
              for (int a = 0; a < 2; ++a) out.uv[6 * tri + 2 * q + a] = co.t >= 0 ? (float)tex[2 * (size_t)co.t + a] : 0.f;
            }
            out.has_n[tri] = hn;
            out.has_uv[tri] = ht;
            out.face_mtl[tri] = mtl;
            if (!hn) chunk_all_n[ci] = 0;
            if (!ht) chunk_all_uv[ci] = 0;
          }
          break;
        }
        default: break;
      }
    });
  });
  for (int i = 0; i < T; ++i) {
    out.all_normals = out.all_normals && chunk_all_n[(size_t)i];
    out.all_uvs = out.all_uvs && chunk_all_uv[(size_t)i];
  }
  return out;
}

}  // namespace rtwh
