// TEST INFRASTRUCTURE — CPU oracle. Not part of the product; see oracle/README.md.
//
// The random stream shared (by specification, not by code) with the CUDA backend.
//
// The reference draws from rand 0.9.0-alpha.1 `ThreadRng` (lib.rs:34-35,64), an OS-seeded ChaCha
// that cannot be seeded through the trait signatures, so sampling parity with the reference
// itself is statistical only.  To make oracle-vs-GPU comparisons deterministic
// both sides use a counter-based generator instead, and only restate the *distributions* of
// rand's f32 sampling:
//   gen::<f32>()        = (u32 >> 8) * 2^-24                        (rand `Standard` for f32)
//   gen_range(lo..hi)   = v12 * (hi-lo) + (lo - (hi-lo)), v12 in [1,2) from the top 23 bits
//                                                                    (rand `UniformFloat::sample_single`)
// Generator: Philox4x32-10 (Salmon et al., SC'11), key = (pixel_index, sample_index),
// counter = (block, stage, seed_lo, seed_hi); draw i of a stage is word i%4 of block i/4.
// stage 0 = pixel jitter + camera ray (lib.rs:84-86), stage b+1 = scatter at bounce b.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
  uint32_t key[2] = {0, 0};
  uint32_t ctr[4] = {0, 0, 0, 0};
  uint32_t buf[4];
  int idx = 4;
  uint64_t draws = 0;

  Rng() {}
  Rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage) { begin(seed, pixel, sample, stage); }

  void begin(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage) {
    key[0] = pixel; key[1] = sample;
    ctr[0] = 0; ctr[1] = stage; ctr[2] = (uint32_t)seed; ctr[3] = (uint32_t)(seed >> 32);
    idx = 4;
  }
  void set_stage(uint32_t stage) { ctr[0] = 0; ctr[1] = stage; idx = 4; }

  uint32_t next_u32() {
    if (idx == 4) { philox4x32_10(ctr, key, buf); ctr[0] += 1; idx = 0; }
    ++draws;
    return buf[idx++];
  }
  // A draw that does not advance the sequential stream: word 0 of the block (id, 0x80000000 | stage).
  // Used where the reference draws INSIDE hit() (ConstantMedium, volumes.rs:58): the BVH backend tests
  // objects in another order than the reference's list, so those draws are keyed by the object instead
  // of by their position in the stream.
  float gen_f32_keyed(uint32_t id) const {
    uint32_t c[4] = {id, 0x80000000u | ctr[1], ctr[2], ctr[3]}, out[4];
    philox4x32_10(c, key, out);
    return (float)(out[0] >> 8) * (1.0f / 16777216.0f);
  }
  // rand `Standard` for f32: 24 random bits, [0,1)
  float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
  // rand `Standard` for f64: 53 random bits from a u64 (low word drawn first)
  double gen_f64() {
    uint64_t lo = next_u32(), hi = next_u32();
    return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
  }
  // rand UniformFloat<f32>::sample_single (half-open range)
  float gen_range(float lo, float hi) {
    uint32_t bits = 0x3F800000u | (next_u32() >> 9);
    float v12; std::memcpy(&v12, &bits, 4);
    float scale = hi - lo;
    float offset = lo - scale;
    float res = v12 * scale + offset;
    if (!(res < hi)) {  // rand retries with a smaller scale; equivalent for our purposes
      uint32_t hb; std::memcpy(&hb, &hi, 4);
      hb = (hi > 0.0f) ? hb - 1 : hb + 1;
      std::memcpy(&res, &hb, 4);
    }
    return res;
  }
  // integer in [0, n): widening multiply (distribution of rand's gen_range(0..n) up to bias < 2^-32 n)
  uint32_t gen_below(uint32_t n) { return (uint32_t)(((uint64_t)next_u32() * n) >> 32); }
};

}  // namespace orc
