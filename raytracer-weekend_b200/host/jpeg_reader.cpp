// JPEG decoder for ImageTexture::open (image_texture.rs:23-30: `image::open(path)`; the reference's textures
// models/earthmap.jpg (baseline) and models/capsule0.jpg (progressive) are JPEG files).
//
// Scope: 8-bit Huffman JPEG — baseline / extended sequential (SOF0, SOF1) and progressive (SOF2, spectral selection +
// successive approximation), 1 or 3 components (grey, YCbCr), restart intervals, any sampling factors.
// Arithmetic: JPEG decoders may differ by +-1 LSB (IDCT rounding, chroma upsampling: SURVEY.md §8c).  This one follows
// the arithmetic of the IJG / libjpeg-turbo defaults — the "islow" integer IDCT (Loeffler-Ligtenberg-Moschytz, 13-bit
// constants, 2 extra bits after the column pass), the 16-bit fixed-point YCbCr -> RGB tables and, for 2x1 / 2x2
// subsampled chroma, the "fancy" triangle-filter upsampling — so that its texels equal those of the decoder that
// produced the committed assets/earthmap.rtwi fixture (PIL = libjpeg-turbo); tests/test_host.py checks that bit for
// bit and cross-checks synthetic files (progressive, restart markers, 4:2:0 / 4:2:2 / grey) against PIL.  The
// reference's own decoder (zune-jpeg 0.4.13 behind image 0.25.2) uses the same IDCT family; any +-1 differences
// against it remain "texel parity unpinned" as SURVEY.md says.  Other sampling ratios are upsampled by replication.
// Not supported (refused with a message): 12-bit, arithmetic coding, lossless, CMYK / 4 components.
#include <cstdio>
#include <cstring>

#include "rtw_host.hpp"

namespace rtwh {

namespace {

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  bool present = false;
  uint8_t bits[17] = {0};
  uint8_t vals[256] = {0};
  int32_t maxcode[18];
  int32_t valptr[17];
  int32_t mincode[17];
  void build() {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l];
      k += bits[l];
      maxcode[l] = bits[l] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int blocks_w = 0, blocks_h = 0;  // allocated (MCU-padded) block grid
  int width = 0, height = 0;       // downsampled size in samples (ceil)
  std::vector<int16_t> coef;       // blocks_w * blocks_h * 64, natural order
  std::vector<uint8_t> plane;      // blocks_w*8 x blocks_h*8 samples
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int nbits = 0;
  bool hit_marker = false;
  void reset() { acc = 0; nbits = 0; hit_marker = false; }
  void fill() {
    while (nbits <= 24) {
      int byte = 0;
      if (!hit_marker && p < end) {
        byte = *p;
        if (byte == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {  // a marker: feed zeros from here on (the caller resynchronises)
            hit_marker = true;
            byte = 0;
          }
        } else {
          ++p;
        }
      }
      acc |= (uint32_t)byte << (24 - nbits);
      nbits += 8;
    }
  }
  int get(int n) {  // n <= 16
    if (n == 0) return 0;
    if (nbits < n) fill();
    int v = (int)(acc >> (32 - n));
    acc <<= n;
    nbits -= n;
    return v;
  }
  int bit() { return get(1); }
  int decode(const Huff& h) {
    if (nbits < 16) fill();
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
      code = (code << 1) | (int)(acc >> 31);
      acc <<= 1;
      nbits -= 1;
      if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    throw Error("JPEG: bad Huffman code");
  }
  static int extend(int v, int n) { return (n && v < (1 << (n - 1))) ? v - (1 << n) + 1 : v; }
};

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// jpeg_idct_islow: 8x8 block of DEQUANTISED coefficients (natural order) -> 8x8 samples
void idct_islow(const int* in, uint8_t* out, int stride) {
  const int C_BITS = 13, P1 = 2;
  const int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
            F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069, F_2_053119869 = 16819,
            F_2_562915447 = 20995, F_3_072711026 = 25172;
  int ws[64];
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < 8; ++i) {
      int x0, x1, x2, x3, x4, x5, x6, x7;
      if (pass == 0) {
        x0 = in[i]; x1 = in[8 + i]; x2 = in[16 + i]; x3 = in[24 + i]; x4 = in[32 + i]; x5 = in[40 + i]; x6 = in[48 + i]; x7 = in[56 + i];
      } else {
        const int* w = ws + 8 * i;
        x0 = w[0]; x1 = w[1]; x2 = w[2]; x3 = w[3]; x4 = w[4]; x5 = w[5]; x6 = w[6]; x7 = w[7];
      }
      // even part
      int z1 = (x2 + x6) * F_0_541196100;
      int tmp2 = z1 + x6 * (-F_1_847759065);
      int tmp3 = z1 + x2 * F_0_765366865;
      int tmp0 = (x0 + x4) * (1 << C_BITS);
      int tmp1 = (x0 - x4) * (1 << C_BITS);
      const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
      // odd part
      tmp0 = x7; tmp1 = x5; tmp2 = x3; tmp3 = x1;
      z1 = tmp0 + tmp3;
      int z2 = tmp1 + tmp2, z3 = tmp0 + tmp2, z4 = tmp1 + tmp3;
      const int z5 = (z3 + z4) * F_1_175875602;
      tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
      z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
      z3 += z5; z4 += z5;
      tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
      if (pass == 0) {
        const int n = C_BITS - P1;
        ws[i] = descale(tmp10 + tmp3, n); ws[56 + i] = descale(tmp10 - tmp3, n);
        ws[8 + i] = descale(tmp11 + tmp2, n); ws[48 + i] = descale(tmp11 - tmp2, n);
        ws[16 + i] = descale(tmp12 + tmp1, n); ws[40 + i] = descale(tmp12 - tmp1, n);
        ws[24 + i] = descale(tmp13 + tmp0, n); ws[32 + i] = descale(tmp13 - tmp0, n);
      } else {
        const int n = C_BITS + P1 + 3;
        uint8_t* o = out + (size_t)i * stride;
        o[0] = clamp8(descale(tmp10 + tmp3, n) + 128); o[7] = clamp8(descale(tmp10 - tmp3, n) + 128);
        o[1] = clamp8(descale(tmp11 + tmp2, n) + 128); o[6] = clamp8(descale(tmp11 - tmp2, n) + 128);
        o[2] = clamp8(descale(tmp12 + tmp1, n) + 128); o[5] = clamp8(descale(tmp12 - tmp1, n) + 128);
        o[3] = clamp8(descale(tmp13 + tmp0, n) + 128); o[4] = clamp8(descale(tmp13 - tmp0, n) + 128);
      }
    }
  }
}

struct Decoder {
  const uint8_t* data;
  size_t size;
  int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
  bool progressive = false, have_frame = false;
  int restart_interval = 0;
  int adobe_transform = -1;
  uint16_t qt[4][64];
  bool qt_present[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  Component comp[3];
  uint32_t eobrun = 0;

  static uint16_t be16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }

  void parse_dqt(const uint8_t* p, size_t len) {
    size_t i = 0;
    while (i < len) {
      const int pq = p[i] >> 4, tq = p[i] & 15;
      ++i;
      if (tq > 3) throw Error("JPEG: bad quantisation table id");
      for (int k = 0; k < 64; ++k) {
        if (i + (pq ? 2 : 1) > len) throw Error("JPEG: truncated DQT");
        qt[tq][ZIGZAG[k]] = pq ? be16(p + i) : p[i];
        i += pq ? 2 : 1;
      }
      qt_present[tq] = true;
    }
  }
  void parse_dht(const uint8_t* p, size_t len) {
    size_t i = 0;
    while (i < len) {
      if (i + 17 > len) throw Error("JPEG: truncated DHT");
      const int tc = p[i] >> 4, th = p[i] & 15;
      if (tc > 1 || th > 3) throw Error("JPEG: bad Huffman table id");
      Huff& h = tc ? ac[th] : dc[th];
      int total = 0;
      h.bits[0] = 0;
      for (int l = 1; l <= 16; ++l) { h.bits[l] = p[i + l]; total += h.bits[l]; }
      i += 17;
      if (total > 256 || i + total > len) throw Error("JPEG: bad DHT");
      memcpy(h.vals, p + i, total);
      i += total;
      h.present = true;
      h.build();
    }
  }
  void parse_sof(const uint8_t* p, size_t len, int marker) {
    if (len < 6) throw Error("JPEG: truncated SOF");
    if (p[0] != 8) throw Error("JPEG: only 8-bit samples are supported");
    height = be16(p + 1);
    width = be16(p + 3);
    ncomp = p[5];
    if (ncomp != 1 && ncomp != 3) throw Error("JPEG: " + std::to_string(ncomp) + " components (CMYK / YCCK) are not supported");
    if (width == 0 || height == 0 || len < 6 + 3 * (size_t)ncomp) throw Error("JPEG: bad SOF");
    progressive = marker == 0xC2;
    hmax = vmax = 1;
    for (int c = 0; c < ncomp; ++c) {
      comp[c].id = p[6 + 3 * c];
      comp[c].h = p[7 + 3 * c] >> 4;
      comp[c].v = p[7 + 3 * c] & 15;
      comp[c].tq = p[8 + 3 * c];
      if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) throw Error("JPEG: bad component");
      hmax = std::max(hmax, comp[c].h);
      vmax = std::max(vmax, comp[c].v);
    }
    if (ncomp == 1) comp[0].h = comp[0].v = hmax = vmax = 1;  // a single-component scan is never interleaved
    mcus_x = (width + 8 * hmax - 1) / (8 * hmax);
    mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
    for (int c = 0; c < ncomp; ++c) {
      Component& k = comp[c];
      k.blocks_w = mcus_x * k.h;
      k.blocks_h = mcus_y * k.v;
      k.width = (width * k.h + hmax - 1) / hmax;
      k.height = (height * k.v + vmax - 1) / vmax;
      k.coef.assign((size_t)k.blocks_w * k.blocks_h * 64, 0);
    }
    have_frame = true;
  }

  // ---- entropy decoding of one block -------------------------------------------------------------------------------
  void block_baseline(BitReader& br, Component& k, int16_t* b, int& pred) {
    const int t = br.decode(dc[k.td]);
    if (t > 11) throw Error("JPEG: bad DC size");
    const int diff = t ? BitReader::extend(br.get(t), t) : 0;
    pred += diff;
    b[0] = (int16_t)pred;
    for (int i = 1; i < 64;) {
      const int rs = br.decode(ac[k.ta]), r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r != 15) break;
        i += 16;
        continue;
      }
      i += r;
      if (i > 63) throw Error("JPEG: AC index out of range");
      b[ZIGZAG[i]] = (int16_t)BitReader::extend(br.get(s), s);
      ++i;
    }
  }
  void block_dc_first(BitReader& br, Component& k, int16_t* b, int& pred, int al) {
    const int t = br.decode(dc[k.td]);
    const int diff = t ? BitReader::extend(br.get(t), t) : 0;
    pred += diff;
    b[0] = (int16_t)(pred * (1 << al));
  }
  void block_dc_refine(BitReader& br, int16_t* b, int al) {
    if (br.bit()) b[0] = (int16_t)(b[0] | (1 << al));
  }
  void block_ac_first(BitReader& br, Component& k, int16_t* b, int ss, int se, int al) {
    if (eobrun > 0) { --eobrun; return; }
    for (int i = ss; i <= se;) {
      const int rs = br.decode(ac[k.ta]), r = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (r < 15) {
          eobrun = (1u << r) - 1u;
          if (r) eobrun += (uint32_t)br.get(r);
          break;
        }
        i += 16;
        continue;
      }
      i += r;
      if (i > 63) throw Error("JPEG: AC index out of range");
      b[ZIGZAG[i]] = (int16_t)(BitReader::extend(br.get(s), s) * (1 << al));
      ++i;
    }
  }
  void block_ac_refine(BitReader& br, Component& k, int16_t* b, int ss, int se, int al) {
    const int p1 = 1 << al, m1 = -(1 << al);
    int i = ss;
    if (eobrun == 0) {
      for (; i <= se;) {
        const int rs = br.decode(ac[k.ta]);
        int r = rs >> 4, s = rs & 15, value = 0;
        if (s == 0) {
          if (r < 15) {
            eobrun = (1u << r);
            if (r) eobrun += (uint32_t)br.get(r);
            break;
          }
        } else {
          if (s != 1) throw Error("JPEG: bad refinement code");
          value = br.bit() ? p1 : m1;
        }
        // skip r zero-history coefficients, refining the non-zero ones on the way
        for (; i <= se; ++i) {
          int16_t& c = b[ZIGZAG[i]];
          if (c != 0) {
            if (br.bit() && (c & p1) == 0) c = (int16_t)(c >= 0 ? c + p1 : c + m1);
          } else {
            if (r == 0) break;
            --r;
          }
        }
        if (value && i <= se) b[ZIGZAG[i]] = (int16_t)value;
        ++i;
      }
    }
    if (eobrun > 0) {  // refine the rest of the band of a block inside an EOB run
      for (; i <= se; ++i) {
        int16_t& c = b[ZIGZAG[i]];
        if (c != 0 && br.bit() && (c & p1) == 0) c = (int16_t)(c >= 0 ? c + p1 : c + m1);
      }
      --eobrun;
    }
  }

  // ---- one scan --------------------------------------------------------------------------------------------------------
  const uint8_t* scan(const uint8_t* p, const uint8_t* end) {
    const size_t len = be16(p);
    const int ns = p[2];
    if (ns < 1 || ns > ncomp || len < 6 + 2 * (size_t)ns) throw Error("JPEG: bad SOS");
    int idx[3];
    for (int i = 0; i < ns; ++i) {
      int c = -1;
      for (int j = 0; j < ncomp; ++j)
        if (comp[j].id == p[3 + 2 * i]) c = j;
      if (c < 0) throw Error("JPEG: SOS names an unknown component");
      comp[c].td = p[4 + 2 * i] >> 4;
      comp[c].ta = p[4 + 2 * i] & 15;
      if (comp[c].td > 3 || comp[c].ta > 3) throw Error("JPEG: bad table selector");
      idx[i] = c;
    }
    const int ss = p[3 + 2 * ns], se = p[4 + 2 * ns], ah = p[5 + 2 * ns] >> 4, al = p[5 + 2 * ns] & 15;
    if (progressive) {
      if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) throw Error("JPEG: bad progressive scan");
    } else if (ss != 0 || se != 63) {
      throw Error("JPEG: bad sequential scan parameters");
    }
    for (int i = 0; i < ns; ++i) {
      const Component& k = comp[idx[i]];
      if ((ss == 0 && !(progressive && ah) && !dc[k.td].present) || (se > 0 && !ac[k.ta].present)) throw Error("JPEG: missing Huffman table");
    }
    BitReader br{p + len, end};
    int pred[3] = {0, 0, 0};
    eobrun = 0;
    int restarts_left = restart_interval, next_rst = 0;
    // a non-interleaved scan covers only the component's real blocks (ceil(size / 8)), an interleaved one whole MCUs
    const bool interleaved = ns > 1;
    const Component& k0 = comp[idx[0]];
    const int units_x = interleaved ? mcus_x : (k0.width + 7) / 8, units_y = interleaved ? mcus_y : (k0.height + 7) / 8;
    for (int uy = 0; uy < units_y; ++uy)
      for (int ux = 0; ux < units_x; ++ux) {
        if (restart_interval && restarts_left == 0) {
          // byte-align, expect RSTn
          br.reset();
          const uint8_t* q = br.p;
          while (q + 1 < end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
          if (q + 1 >= end || q[1] != 0xD0 + next_rst) throw Error("JPEG: restart marker out of sequence");
          br.p = q + 2;
          next_rst = (next_rst + 1) & 7;
          restarts_left = restart_interval;
          pred[0] = pred[1] = pred[2] = 0;
          eobrun = 0;
        }
        for (int i = 0; i < ns; ++i) {
          Component& k = comp[idx[i]];
          const int bw = interleaved ? k.h : 1, bh = interleaved ? k.v : 1;
          for (int by = 0; by < bh; ++by)
            for (int bx = 0; bx < bw; ++bx) {
              const int X = ux * bw + bx, Y = uy * bh + by;
              int16_t* b = &k.coef[((size_t)Y * k.blocks_w + X) * 64];
              if (!progressive) block_baseline(br, k, b, pred[i]);
              else if (ss == 0) { if (ah == 0) block_dc_first(br, k, b, pred[i], al); else block_dc_refine(br, b, al); }
              else if (ah == 0) block_ac_first(br, k, b, ss, se, al);
              else block_ac_refine(br, k, b, ss, se, al);
            }
        }
        if (restart_interval) --restarts_left;
      }
    // continue after the entropy-coded segment: the next marker
    const uint8_t* q = br.hit_marker ? br.p : br.p;
    while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
    return q;
  }

  void decode(std::vector<uint8_t>& rgb) {
    if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) throw Error("JPEG: no SOI marker");
    const uint8_t* p = data + 2;
    const uint8_t* end = data + size;
    bool done = false, any_scan = false;
    while (!done && p + 4 <= end) {
      if (p[0] != 0xFF) { ++p; continue; }
      const int m = p[1];
      if (m == 0xFF) { ++p; continue; }
      if (m == 0xD9) break;
      if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { p += 2; continue; }
      const size_t len = be16(p + 2);
      if (len < 2 || p + 2 + len > end) throw Error("JPEG: truncated segment");
      const uint8_t* body = p + 4;
      switch (m) {
        case 0xDB: parse_dqt(body, len - 2); break;
        case 0xC4: parse_dht(body, len - 2); break;
        case 0xC0: case 0xC1: case 0xC2:
          if (have_frame) throw Error("JPEG: more than one frame");
          parse_sof(body, len - 2, m);
          break;
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          throw Error("JPEG: unsupported coding process (lossless / hierarchical / arithmetic)");
        case 0xDD: restart_interval = be16(body); break;
        case 0xEE:
          if (len >= 14 && memcmp(body, "Adobe", 5) == 0) adobe_transform = body[11];
          break;
        case 0xDA:
          if (!have_frame) throw Error("JPEG: scan before frame header");
          p = scan(p + 2, end);
          any_scan = true;
          continue;
        default: break;
      }
      p += 2 + len;
    }
    if (!have_frame || !any_scan) throw Error("JPEG: no image data");
    // ---- dequantise + IDCT ------------------------------------------------------------------------------------------------
    for (int c = 0; c < ncomp; ++c) {
      Component& k = comp[c];
      if (!qt_present[k.tq]) throw Error("JPEG: missing quantisation table");
      const int stride = k.blocks_w * 8;
      k.plane.assign((size_t)stride * k.blocks_h * 8, 0);
      int deq[64];
      for (int by = 0; by < k.blocks_h; ++by)
        for (int bx = 0; bx < k.blocks_w; ++bx) {
          const int16_t* b = &k.coef[((size_t)by * k.blocks_w + bx) * 64];
          for (int i = 0; i < 64; ++i) deq[i] = (int)b[i] * (int)qt[k.tq][i];
          idct_islow(deq, &k.plane[(size_t)by * 8 * stride + bx * 8], stride);
        }
    }
    // ---- upsample to full resolution ----------------------------------------------------------------------------------------
    std::vector<uint8_t> full[3];
    for (int c = 0; c < ncomp; ++c) {
      const Component& k = comp[c];
      full[c].resize((size_t)width * height);
      const int stride = k.blocks_w * 8;
      const int hs = hmax / k.h, vs = vmax / k.v;
      const bool integral = (hmax % k.h == 0) && (vmax % k.v == 0);
      auto in = [&](int x, int y) -> int {  // edge rows / columns replicate (libjpeg's context rows, real columns only)
        x = x < 0 ? 0 : (x >= k.width ? k.width - 1 : x);
        y = y < 0 ? 0 : (y >= k.height ? k.height - 1 : y);
        return k.plane[(size_t)y * stride + x];
      };
      if (integral && hs == 1 && vs == 1) {
        for (int y = 0; y < height; ++y) memcpy(&full[c][(size_t)y * width], &k.plane[(size_t)y * stride], width);
      } else if (integral && hs == 2 && vs == 1) {  // h2v1_fancy_upsample: 3/4 nearer + 1/4 further, rounding alternates
        for (int y = 0; y < height; ++y)
          for (int x = 0; x < width; ++x) {
            const int i = x >> 1;
            int v;
            if (k.width == 1) v = in(0, y);
            else if ((x & 1) == 0) v = (i == 0) ? in(0, y) : (3 * in(i, y) + in(i - 1, y) + 1) >> 2;
            else v = (i == k.width - 1) ? in(i, y) : (3 * in(i, y) + in(i + 1, y) + 2) >> 2;
            full[c][(size_t)y * width + x] = (uint8_t)v;
          }
      } else if (integral && hs == 2 && vs == 2) {  // h2v2_fancy_upsample: 9/16, 3/16, 3/16, 1/16
        for (int y = 0; y < height; ++y) {
          const int j = y >> 1, jn = (y & 1) ? j + 1 : j - 1;  // nearer row j, further row above / below
          for (int x = 0; x < width; ++x) {
            const int i = x >> 1;
            auto colsum = [&](int ii) { return 3 * in(ii, j) + in(ii, jn); };
            const int t = colsum(i);
            int v;
            if (k.width == 1) v = (t * 4 + 8) >> 4;
            else if ((x & 1) == 0) v = (i == 0) ? (t * 4 + 8) >> 4 : (t * 3 + colsum(i - 1) + 8) >> 4;
            else v = (i == k.width - 1) ? (t * 4 + 7) >> 4 : (t * 3 + colsum(i + 1) + 7) >> 4;
            full[c][(size_t)y * width + x] = (uint8_t)v;
          }
        }
      } else {  // any other ratio: replication (not decoder-exact; no asset of the reference needs it)
        for (int y = 0; y < height; ++y)
          for (int x = 0; x < width; ++x) full[c][(size_t)y * width + x] = (uint8_t)in(x * k.h / hmax, y * k.v / vmax);
      }
    }
    // ---- colour -----------------------------------------------------------------------------------------------------------
    rgb.resize((size_t)width * height * 3);
    const size_t n = (size_t)width * height;
    if (ncomp == 1) {
      for (size_t i = 0; i < n; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = full[0][i];
    } else if (adobe_transform == 0) {  // Adobe marker says the three components are already RGB
      for (size_t i = 0; i < n; ++i) { rgb[3 * i] = full[0][i]; rgb[3 * i + 1] = full[1][i]; rgb[3 * i + 2] = full[2][i]; }
    } else {  // jdcolor.c: 16-bit fixed point, FIX(x) = (int)(x * 65536 + 0.5)
      int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
      for (int i = 0; i < 256; ++i) {
        const int x = i - 128;
        cr_r[i] = (91881 * x + 32768) >> 16;
        cb_b[i] = (116130 * x + 32768) >> 16;
        cr_g[i] = -46802 * x;
        cb_g[i] = -22554 * x + 32768;
      }
      for (size_t i = 0; i < n; ++i) {
        const int y = full[0][i], cb = full[1][i], cr = full[2][i];
        rgb[3 * i] = clamp8(y + cr_r[cr]);
        rgb[3 * i + 1] = clamp8(y + ((cb_g[cb] + cr_g[cr]) >> 16));
        rgb[3 * i + 2] = clamp8(y + cb_b[cb]);
      }
    }
  }
};

}  // namespace

// Returns false when `path` does not exist or is not a JPEG; throws Error on a JPEG it cannot decode.
bool read_jpeg(const std::string& path, std::vector<uint8_t>& rgb, uint32_t& width, uint32_t& height) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<uint8_t> file;
  uint8_t buf[1 << 16];
  for (size_t n; (n = fread(buf, 1, sizeof(buf), f)) > 0;) file.insert(file.end(), buf, buf + n);
  fclose(f);
  if (file.size() < 4 || file[0] != 0xFF || file[1] != 0xD8) return false;
  Decoder d;
  d.data = file.data();
  d.size = file.size();
  memset(d.qt, 0, sizeof(d.qt));
  try {
    d.decode(rgb);
  } catch (const Error& e) {
    throw Error(path + ": " + e.what());
  }
  width = (uint32_t)d.width;
  height = (uint32_t)d.height;
  return true;
}

}  // namespace rtwh
