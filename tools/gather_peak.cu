// Random 64-byte gather bandwidth of the GPU: the roofline that bounds BVH traversal of a hierarchy that does not
// fit the caches (config C5: every step fetches one 64-byte child pair = 4 x LDG.128 at a data-dependent address).
//   independent : every lane issues gathers at hashed addresses, 8 in flight per lane (maximum memory parallelism)
//   dependent   : every lane chases a chain (next index = hash of the loaded data), one fetch in flight per lane —
//                 the access pattern of a ray walking a tree
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gather_peak tools/gather_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int BYTES>
__device__ __forceinline__ float fetch(const float4* __restrict__ a, size_t rec) {
  const float4* p = a + rec * (BYTES / 16);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < BYTES / 16; ++k) { float4 v = __ldg(p + k); s += v.x + v.w; }
  return s;
}

template <int BYTES>
__global__ void k_independent(const float4* __restrict__ a, unsigned nrec, int iters, float* out) {
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  for (int i = 0; i < iters; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) s += fetch<BYTES>(a, hash32(tid * 9781u + (i + u) * 0x9E3779B9u) % nrec);
  }
  if (s == 123.456f) out[tid] = s;
}

template <int BYTES>
__global__ void k_dependent(const float4* __restrict__ a, unsigned nrec, int iters, float* out) {
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned idx = hash32(tid) % nrec;
  float s = 0.f;
  for (int i = 0; i < iters; ++i) {
    float v = fetch<BYTES>(a, idx);
    s += v;
    idx = hash32(idx + __float_as_uint(v) + i) % nrec;
  }
  if (s == 123.456f) out[tid] = s;
}

template <class K>
static double run(K kernel, const float4* a, unsigned nrec, int iters, int grid, int block, float* out, int bytes) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kernel<<<grid, block>>>(a, nrec, iters, out);  // warm-up
  cudaEventRecord(e0);
  kernel<<<grid, block>>>(a, nrec, iters, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return (double)grid * block * iters * bytes / (ms * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
  size_t mb = argc > 1 ? atol(argv[1]) : 1400;  // working set in MB (C5: 700 MB of pairs + 530 MB of primitives)
  size_t n4 = mb * (1 << 20) / 16;
  float4* a; float* out;
  cudaMalloc(&a, n4 * 16); cudaMalloc(&out, 1 << 26);
  cudaMemset(a, 0x3c, n4 * 16);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int sms = pr.multiProcessorCount;
  printf("%s, %d SMs, working set %zu MB (L2 %d MB)\n", pr.name, sms, mb, pr.l2CacheSize >> 20);
  for (int wps : {16, 32, 64}) {  // resident warps per SM
    int block = 128, grid = sms * wps * 32 / block;
    printf("  %2d warps/SM:  64 B independent %7.1f GB/s   64 B dependent %7.1f GB/s   128 B dependent %7.1f GB/s   32 B dependent %7.1f GB/s\n", wps,
           run(k_independent<64>, a, (unsigned)(n4 / 4), 256, grid, block, out, 64), run(k_dependent<64>, a, (unsigned)(n4 / 4), 256, grid, block, out, 64),
           run(k_dependent<128>, a, (unsigned)(n4 / 8), 256, grid, block, out, 128), run(k_dependent<32>, a, (unsigned)(n4 / 2), 256, grid, block, out, 32));
  }
  return 0;
}
