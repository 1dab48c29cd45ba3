// tools/div_check.cu — SharedDivisor (rtw_device.cuh) against the `/` operator, bit for bit, on the GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/bin/div_check tools/div_check.cu && tools/bin/div_check
// Random operand pairs inside the fast-path window (both exponents in [2^-40, 2^41)), pairs with extreme mantissas,
// zeros of both signs, and operands OUTSIDE the window (denormals, inf, NaN, huge, tiny: they must take the fallback).
// Prints the number of mismatches (must be 0) for each class.
#include <cstdio>
#include <cstdint>
#include "../raytracer-weekend_b200/csrc/rtw_device.cuh"

__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ bool same(float x, float y) {
  return __float_as_uint(x) == __float_as_uint(y) || (x != x && y != y);
}

// mode 0: random mantissas / signs, exponents uniform in the window; 1: any 32-bit patterns (mostly outside the
// window: fallback); 2: mantissas from the adversarial set {0, 1, 0x7fffff, 0x7ffffe, 0x400000, 0x3fffff} x window
__global__ void k_check(uint64_t base, int mode, unsigned long long* bad, unsigned long long* fast_taken, uint32_t* ex) {
  const uint64_t i = base + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nbad = 0, nfast = 0;
  for (int rep = 0; rep < 64; ++rep) {
    const uint32_t ra = mix(i * 64 + rep), rb = mix((i * 64 + rep) ^ 0x9e3779b97f4a7c15ull);
    uint32_t ua, ub;
    if (mode == 1) {
      ua = ra; ub = rb;
      if ((rep & 7) == 0) ua &= 0x80000000u;             // +-0 numerators
      if ((rep & 7) == 1) ub = (ub & 0x807fffffu);       // denormal / zero divisors
    } else {
      const uint32_t ea = 87u + (ra >> 8) % 81u, eb = 87u + (rb >> 8) % 81u;
      uint32_t ma = ra & 0x7fffffu, mb = rb & 0x7fffffu;
      if (mode == 2) {
        const uint32_t adv[6] = {0u, 1u, 0x7fffffu, 0x7ffffeu, 0x400000u, 0x3fffffu};
        ma = adv[(ra >> 3) % 6u]; if (rep & 1) mb = adv[(rb >> 3) % 6u];
      }
      ua = (ra & 0x80000000u) | (ea << 23) | ma;
      ub = (rb & 0x80000000u) | (eb << 23) | mb;
      if (mode == 0 && (rep & 31) == 0) ua &= 0x80000000u;  // zero numerators of both signs
    }
    const float a = __uint_as_float(ua), b = __uint_as_float(ub);
    rtw::SharedDivisor dv;
    dv.set(b);
    const float q = dv.div(a);
    const float ref = a / b;
    if (!same(q, ref)) {
      nbad++;
      const unsigned long long slot = atomicAdd(bad + 2, 1ull);
      if (slot < 8) { ex[5 * slot] = ua; ex[5 * slot + 1] = ub; ex[5 * slot + 2] = __float_as_uint(q); ex[5 * slot + 3] = __float_as_uint(ref); ex[5 * slot + 4] = 0; }
    }
    if (dv.fast && (rtw::in_div_window(a) || ua == 0u)) nfast++;
    // the caller-guaranteed form (flat_rect_run): k and o pass div_operand_safe, k is not -0 -> k - o needs no check
    if (mode != 1) {
      float k = __uint_as_float((ra & 0x80000000u) | ((112u + (ra >> 8) % 55u) << 23) | (ra & 0x7fffffu));   // 2^-15 .. 2^40
      float o = __uint_as_float((rb & 0x80000000u) | ((87u + (rb >> 9) % 80u) << 23) | (rb & 0x7fffffu));     // 2^-40 .. 2^40
      if ((rep & 15) == 3) k = 0.0f;
      if ((rep & 3) == 0) o = k;                          // exact cancellation -> +0
      if ((rep & 15) == 1) o = __uint_as_float(ra & 0x80000000u);   // origin component +-0
      if ((rep & 15) == 2) o = __uint_as_float(__float_as_uint(k) ^ (1u + (rb & 3u)));  // a few ulps apart
      if (rtw::plane_div_safe(k) && rtw::origin_div_safe(o) && dv.fast) {
        const float num = k - o;
        if (!same(dv.div_in_window(num), num / b)) {
          nbad++;
          const unsigned long long slot = atomicAdd(bad + 2, 1ull);
          if (slot < 8) { ex[5 * slot] = __float_as_uint(num); ex[5 * slot + 1] = ub; ex[5 * slot + 2] = __float_as_uint(dv.div_in_window(num)); ex[5 * slot + 3] = __float_as_uint(num / b); ex[5 * slot + 4] = 1; }
        }
      }
    }
  }
  if (nbad) atomicAdd(bad, nbad);
  atomicAdd(fast_taken, nfast);
}

int main() {
  unsigned long long *d, h[3];
  uint32_t *dex, hex_[40];
  cudaMalloc(&d, 24);
  cudaMalloc(&dex, sizeof(hex_));
  const char* names[3] = {"window, random mantissas", "any bit patterns (fallback mostly)", "window, adversarial mantissas"};
  int rc = 0;
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(d, 0, 24);
    const int launches = mode == 0 ? 16 : 4;
    for (int l = 0; l < launches; ++l) k_check<<<16384, 256>>>((uint64_t)l * 16384 * 256, mode, d, d + 1, dex);
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    cudaMemcpy(hex_, dex, sizeof(hex_), cudaMemcpyDeviceToHost);
    for (unsigned long long i = 0; i < h[2] && i < 8; ++i) printf("   mismatch (%s): a = %08x  b = %08x  got %08x  want %08x\n", hex_[5 * i + 4] ? "k - o form" : "div()", hex_[5 * i], hex_[5 * i + 1], hex_[5 * i + 2], hex_[5 * i + 3]);
    const double n = (double)launches * 16384 * 256 * 64;
    printf("%-40s %.3g pairs, fast path taken for %.3g, mismatches vs '/': %llu\n", names[mode], n, (double)h[1], h[0]);
    if (h[0]) rc = 1;
  }
  printf(cudaGetLastError() == cudaSuccess && rc == 0 ? "DIV_CHECK_OK\n" : "DIV_CHECK_FAILED\n");
  return rc;
}
