// x3_synth.cu -- on-device generators for the synthetic signals of SURVEY.md section 8(d).
// Bench / test utility (config 5 is 118 GB of PCM, more than host RAM, so it is generated per shard on
// the device).  Integer-only and bit-identical to oracle/x3o_synth; sample n depends only on (kind, seed, fs, n).
#include <cuda_runtime.h>

#include "x3_kernels.h"
#include "x3_sin1024.h"

namespace x3 {
namespace {

__constant__ short c_sin[1024];

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long hh(uint32_t seed, unsigned long long n) {
  return splitmix64(((unsigned long long)seed << 32) ^ n);
}
__device__ __forceinline__ int uni(uint32_t seed, unsigned long long n, int a) {
  return (int)(hh(seed, n) % (unsigned long long)(2 * a + 1)) - a;
}
__device__ __forceinline__ int colored(uint32_t seed, unsigned long long n, int a) {
  int s = 0;
  for (unsigned long long j = 0; j < 16 && j <= n; j++) s += uni(seed, n - j, a);
  return s >> 2;
}
__device__ __forceinline__ int isin(unsigned long long p, int amp) { return ((int)c_sin[p & 1023] * amp) >> 15; }

__global__ void synth_kernel(int kind, uint32_t seed, uint32_t fs, unsigned long long n0, unsigned long long count,
                             int16_t *out) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const unsigned long long n = n0 + i;
    int v;
    if (kind == 1) {
      v = -3460 + colored(seed, n, 8) + isin((n * 50ull * 1024ull) / fs, 200);
    } else if (kind == 2) {
      const int A[4] = {2, 8, 24, 40};
      v = colored(seed, n, A[(n / fs) % 4]);
      if (n % 196608ull < 64) v += isin((n * 48000ull * 1024ull) / fs, 6000);
    } else {
      uint32_t k = (uint32_t)(hh(seed ^ 0xABCDu, n / 4096) % 6);
      if (k == 5) k = (uint32_t)(hh(seed ^ 0x1234u, n / 20) % 5);
      v = k == 0 ? uni(seed, n, 32767) : k == 1 ? uni(seed, n, 300) : k == 2 ? uni(seed, n, 12) : k == 3 ? 32767 : -32768;
    }
    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
    out[i] = (int16_t)v;
  }
}

}  // namespace

cudaError_t launch_synth(int kind, uint32_t seed, uint32_t fs, unsigned long long n0, unsigned long long count,
                         int16_t *out, cudaStream_t stream) {
  static bool table_ready[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !table_ready[dev]) {
    cudaError_t e = cudaMemcpyToSymbol(c_sin, X3_SIN1024, sizeof(short) * 1024);
    if (e != cudaSuccess) return e;
    table_ready[dev] = true;
  }
  if (count == 0) return cudaSuccess;
  unsigned long long g = (count + 255) / 256;
  if (g > 148ull * 32ull) g = 148ull * 32ull;
  synth_kernel<<<(unsigned)g, 256, 0, stream>>>(kind, seed, fs, n0, count, out);
  return cudaGetLastError();
}

}  // namespace x3
