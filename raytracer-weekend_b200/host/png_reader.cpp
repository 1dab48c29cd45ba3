// PNG decoder for ImageTexture::open (image_texture.rs:23-30: `image::open(path)`, decoded by the png 0.17 crate).
// PNG is lossless, so the texels are EXACTLY what the reference's decoder yields — unlike JPEG, where IDCT and
// upsampling differ between decoders, which is why JPEG assets still come in decoded.
// Supported: non-interlaced, bit depth 8 for every colour type (grey, grey+alpha, RGB, RGBA, palette), plus bit
// depths 1 / 2 / 4 for grey (scaled to 8 bits like the crate's EXPAND transformation) and palette.  Alpha is
// dropped: ImageTexture::value reads r, g, b only (image_texture.rs:44-50).  16-bit and interlaced files are
// refused with a message (their 8-bit reduction / Adam7 pass order are decoder policy, not worth guessing).
#include <zlib.h>

#include <cstdio>
#include <cstring>

#include "rtw_host.hpp"

namespace rtwh {

namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

// Returns false when `path` does not exist or is not a PNG; throws Error on a PNG it cannot decode.
bool read_png(const std::string& path, std::vector<uint8_t>& rgb, uint32_t& width, uint32_t& height) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<uint8_t> file;
  uint8_t buf[1 << 16];
  for (size_t n; (n = fread(buf, 1, sizeof(buf), f)) > 0;) file.insert(file.end(), buf, buf + n);
  fclose(f);
  static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (file.size() < 8 || memcmp(file.data(), SIG, 8) != 0) return false;
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, plte;
  bool seen_end = false;
  for (size_t p = 8; p + 12 <= file.size() && !seen_end;) {
    const uint32_t len = be32(&file[p]);
    const char* type = (const char*)&file[p + 4];
    if (p + 12 + (size_t)len > file.size()) throw Error("PNG truncated: " + path);
    const uint8_t* data = &file[p + 8];
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)type, 4 + len) != be32(data + len))
      throw Error("PNG chunk CRC mismatch: " + path);
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) throw Error("PNG bad IHDR: " + path);
      w = be32(data); h = be32(data + 4);
      depth = data[8]; ctype = data[9]; interlace = data[12];
      if (data[10] != 0 || data[11] != 0) throw Error("PNG unknown compression / filter method: " + path);
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(data, data + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      seen_end = true;
    }
    p += 12 + (size_t)len;
  }
  if (ctype < 0 || w == 0 || h == 0 || idat.empty()) throw Error("PNG without IHDR / IDAT: " + path);
  if (interlace != 0) throw Error("interlaced PNG is not supported: " + path);
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;  // grey
    case 2: channels = 3; break;  // RGB
    case 3: channels = 1; break;  // palette index
    case 4: channels = 2; break;  // grey + alpha
    case 6: channels = 4; break;  // RGBA
    default: throw Error("PNG unknown colour type: " + path);
  }
  const bool small = depth == 1 || depth == 2 || depth == 4;
  if (!(depth == 8 || (small && (ctype == 0 || ctype == 3))))
    throw Error("PNG bit depth " + std::to_string(depth) + " with colour type " + std::to_string(ctype) +
                " is not supported (8-bit, or 1/2/4-bit grey / palette): " + path);
  if (ctype == 3 && plte.size() < 3) throw Error("palette PNG without PLTE: " + path);
  const size_t bits_pp = (size_t)channels * (size_t)depth;
  const size_t stride = ((size_t)w * bits_pp + 7) / 8;
  const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;  // filter unit
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  uLongf out_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size())
    throw Error("PNG inflate failed: " + path);
  // unfilter in place (PNG spec 9.2)
  std::vector<uint8_t> zero(stride, 0);
  for (uint32_t y = 0; y < h; ++y) {
    uint8_t* cur = &raw[(stride + 1) * (size_t)y + 1];
    const uint8_t* up = y ? &raw[(stride + 1) * (size_t)(y - 1) + 1] : zero.data();
    const int ft = raw[(stride + 1) * (size_t)y];
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= bpp ? cur[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
      int pred;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: pred = paeth(a, b, c); break;
        default: throw Error("PNG unknown filter type: " + path);
      }
      cur[x] = (uint8_t)(cur[x] + pred);
    }
  }
  rgb.resize((size_t)w * h * 3);
  const int maxv = (1 << depth) - 1;
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t* row = &raw[(stride + 1) * (size_t)y + 1];
    uint8_t* out = &rgb[(size_t)y * w * 3];
    for (uint32_t x = 0; x < w; ++x) {
      uint8_t r, g, b;
      if (depth == 8) {
        const uint8_t* px = row + (size_t)x * channels;
        if (ctype == 2 || ctype == 6) { r = px[0]; g = px[1]; b = px[2]; }
        else if (ctype == 3) {
          if ((size_t)px[0] * 3 + 2 >= plte.size()) throw Error("PNG palette index out of range: " + path);
          r = plte[px[0] * 3]; g = plte[px[0] * 3 + 1]; b = plte[px[0] * 3 + 2];
        } else { r = g = b = px[0]; }
      } else {  // 1 / 2 / 4 bits: samples packed MSB first
        const size_t bit = (size_t)x * depth;
        const int v = (row[bit >> 3] >> (8 - depth - (int)(bit & 7))) & maxv;
        if (ctype == 3) {
          if ((size_t)v * 3 + 2 >= plte.size()) throw Error("PNG palette index out of range: " + path);
          r = plte[v * 3]; g = plte[v * 3 + 1]; b = plte[v * 3 + 2];
        } else {
          r = g = b = (uint8_t)(v * (255 / maxv));  // 1 bit: x255, 2 bits: x85, 4 bits: x17
        }
      }
      out[3 * x] = r; out[3 * x + 1] = g; out[3 * x + 2] = b;
    }
  }
  width = w; height = h;
  return true;
}

}  // namespace rtwh
