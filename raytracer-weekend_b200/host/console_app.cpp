// console_app with a `--backend cuda` switch: the C++ mirror of console_app/src/main.rs.
//
//   console_app [--width|-w 400] [--aspect-ratio|-a 1.7777778] [--samples-per-pixel|-s 100]
//               [--backend cuda] [--gpus N] [--seed N] [--device D] [--lib PATH] [--out-dir render] <scene>
//
// Same flow as main.rs:28-96: image_height = round(width / aspect_ratio); Scene::generate with
// aspect = width/height; one Raytracer per camera; divide by spp, gamma 2, clamp, u8; save
// render/image_NNNN.png.  The only backend is the CUDA one — `--backend cpu` is the Rust
// reference itself and is refused here (no CPU fallback exists in this repo's product code).
#include <sys/stat.h>
#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rtw_host.hpp"

namespace {

void put_u32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
  put_u32(out, (uint32_t)data.size());
  size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  out.insert(out.end(), data.begin(), data.end());
  uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start));
  put_u32(out, crc);
}
// RGB8 PNG (what image::RgbImage::save writes, main.rs:92-94)
bool save_png(const std::string& path, const uint8_t* rgb, uint32_t w, uint32_t h) {
  std::vector<uint8_t> raw;
  raw.reserve((size_t)h * (1 + 3 * (size_t)w));
  for (uint32_t y = 0; y < h; ++y) {
    raw.push_back(0);
    raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  comp.resize(clen);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  put_u32(ihdr, w); put_u32(ihdr, h);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  png_chunk(out, "IHDR", ihdr);
  png_chunk(out, "IDAT", comp);
  png_chunk(out, "IEND", {});
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  fclose(f);
  return ok;
}

[[noreturn]] void usage(const char* msg) {
  if (msg) fprintf(stderr, "error: %s\n\n", msg);
  fprintf(stderr,
          "USAGE: console_app [OPTIONS] <SCENE>\n"
          "  -w, --width <WIDTH>                    [default: 400]\n"
          "  -a, --aspect-ratio <ASPECT_RATIO>      [default: 1.7777778]\n"
          "  -s, --samples-per-pixel <SPP>          [default: 100]\n"
          "      --backend <cuda>                   [default: cuda]\n"
          "      --gpus <N>                         spread every frame over N GPUs of the box [default: 1]\n"
          "      --seed <N>  --device <D>  --lib <librtw_cuda.so>  --out-dir <DIR>\n"
          "      --progress-out <FILE>   also write the frames as a ProgressMessage stream (postcard + COBS,\n"
          "                              the wire format of discovery_host_receiver)\n"
          "SCENES:");
  for (const auto& n : rtwh::scene_names()) fprintf(stderr, " %s", n.c_str());
  fprintf(stderr, "\n");
  exit(2);
}

}  // namespace

int main(int argc, char** argv) {
  uint32_t width = 400, spp = 100;
  double aspect_ratio = 1.7777778;
  std::string backend = "cuda", scene, out_dir = "render", lib, progress_out;
  uint64_t seed = 1;
  int device = 0;
  uint32_t gpus = 1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) usage(("missing value for " + a).c_str());
      return argv[++i];
    };
    if (a == "-w" || a == "--width") width = (uint32_t)std::stoul(val());
    else if (a == "-a" || a == "--aspect-ratio") aspect_ratio = std::stod(val());
    else if (a == "-s" || a == "--samples-per-pixel") spp = (uint32_t)std::stoul(val());
    else if (a == "--backend") backend = val();
    else if (a == "--seed") seed = std::stoull(val());
    else if (a == "--device") device = std::stoi(val());
    else if (a == "--gpus") gpus = (uint32_t)std::stoul(val());
    else if (a == "--lib") lib = val();
    else if (a == "--out-dir") out_dir = val();
    else if (a == "--progress-out") progress_out = val();
    else if (a == "-h" || a == "--help") usage(nullptr);
    else if (!a.empty() && a[0] == '-') usage(("unknown option " + a).c_str());
    else scene = a;
  }
  if (scene.empty()) usage("missing scene subcommand");
  if (backend != "cuda") {
    fprintf(stderr, "error: backend '%s' is not available here: the CPU backend is the Rust reference itself; "
                    "this front end only drives `--backend cuda`\n", backend.c_str());
    return 2;
  }
  if (lib.empty()) {
    // next to the executable: ../lib/librtw_cuda.so
    std::string self = argv[0];
    size_t s = self.find_last_of('/');
    lib = (s == std::string::npos ? std::string(".") : self.substr(0, s)) + "/../lib/librtw_cuda.so";
  }
  // main.rs:31-41
  uint32_t image_width = width;
  uint32_t image_height = (uint32_t)std::llround((double)image_width / aspect_ratio);
  try {
    rtwh::World world = rtwh::generate_scene(scene, (float)image_width / (float)image_height, seed);
    mkdir(out_dir.c_str(), 0755);
    // main.rs:48-95.  The reference builds a Raytracer per camera over the same world; the backend keeps that
    // world resident (flattened + built once) and overlaps the PNG encoding of frame n with the rendering of n+1.
    rtw_sink sink;
    if (rtwh_sink_open(lib.c_str(), "rtw_", device, &sink) != RTW_OK) {
      fprintf(stderr, "error: %s\n", rtwh_last_error());
      return 1;
    }
    FILE* progress = progress_out.empty() ? nullptr : fopen(progress_out.c_str(), "wb");
    if (!progress_out.empty() && !progress) {
      fprintf(stderr, "error: cannot write %s\n", progress_out.c_str());
      return 1;
    }
    bool io_ok = true;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t frames = rtwh::render_animation(
        world, image_width, image_height, spp, &sink, seed,
        [&](uint32_t frame_no, const std::vector<rtwh::Pixel>& all_pixels, const rtw_render_stats& st) {
          // main.rs:66-90
          std::vector<uint8_t> img((size_t)image_width * image_height * 3);
          float scale = 1.0f / (float)spp;
          for (size_t i = 0; i < all_pixels.size(); ++i)
            for (int c = 0; c < 3; ++c) {
              float v = std::sqrt(scale * all_pixels[i].color.e[c]);
              float cl = v < 0.0f ? 0.0f : (v > 0.999f ? 0.999f : v);
              float b = 255.999f * cl;
              img[3 * i + c] = (b != b || b <= 0.0f) ? 0 : (b >= 255.0f ? 255 : (uint8_t)b);
            }
          char name[64];
          snprintf(name, sizeof(name), "/image_%04u.png", frame_no);
          if (!save_png(out_dir + name, img.data(), image_width, image_height)) {
            fprintf(stderr, "error: cannot write %s%s\n", out_dir.c_str(), name);
            io_ok = false;
            return false;
          }
          if (progress) {  // lib.rs:128-138 as discovery_app streams it (raytracer.rs:62-111)
            rtwh::ProgressMessage m;
            m.kind = rtwh::ProgressMessage::ImageStart;
            m.width = image_width; m.height = image_height; m.samples_per_pixel = spp;
            std::vector<uint8_t> b = rtwh::to_vec_cobs(m);
            fwrite(b.data(), 1, b.size(), progress);
            m.kind = rtwh::ProgressMessage::PixelMsg;
            for (const rtwh::Pixel& px : all_pixels) {
              m.pixel = px;
              b = rtwh::to_vec_cobs(m);
              fwrite(b.data(), 1, b.size(), progress);
            }
            m.kind = rtwh::ProgressMessage::ImageEnd;
            b = rtwh::to_vec_cobs(m);
            fwrite(b.data(), 1, b.size(), progress);
          }
          fprintf(stderr, "frame %u: %ux%u, %u spp, %llu segments, %u GPU(s), render %.1f ms (%.1f Mrays/s) -> %s%s\n", frame_no,
                  image_width, image_height, spp, (unsigned long long)st.segments, st.gpus, st.ms_render,
                  st.ms_render > 0 ? st.segments / (st.ms_render * 1e3) : 0.0, out_dir.c_str(), name);
          return true;
        },
        gpus);
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (progress) fclose(progress);
    rtwh_sink_close(&sink);
    fprintf(stderr, "%u frame(s) in %.3f s (flatten + build + render + PNG)\n", frames, wall);
    if (!io_ok) return 1;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
