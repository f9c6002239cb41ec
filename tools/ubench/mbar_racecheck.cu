// Does compute-sanitizer's racecheck follow mbarrier arrive (release) -> try_wait (acquire) ordering?
// Warp 0 writes shared memory and arrives; warp 1 waits on the mbarrier's phase and reads.  The program is race free
// under the PTX memory model; run it under `compute-sanitizer --tool racecheck` with MODE=0 (mbarrier) and MODE=1
// (__syncthreads) to see what the tool reports for each.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DMODE=0 -o mbar_racecheck mbar_racecheck.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#ifndef MODE
#define MODE 0
#endif
__global__ void k(uint32_t *out, int iters) {
  __shared__ uint32_t data[32];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
  __syncthreads();
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    if (wid == 0) {
      data[lane] = (uint32_t)it * 32u + (uint32_t)lane;
      __syncwarp();
#if MODE == 0
      if (lane == 0) asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(mb) : "memory");
#endif
    }
#if MODE == 1
    __syncthreads();
#endif
    if (wid == 1) {
#if MODE == 0
      asm volatile(
          "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mb),
          "r"((uint32_t)(it & 1))
          : "memory");
#endif
      acc += data[31 - lane];
    }
    __syncthreads();   // data[] is rewritten in the next iteration
  }
  if (wid == 1) out[lane] = acc;
}
int main() {
  uint32_t *d, h[32];
  cudaMalloc(&d, 128);
  k<<<1, 64>>>(d, 64);
  cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  uint32_t want = 0;
  for (int it = 0; it < 64; it++) want += (uint32_t)it * 32u + 31u;
  printf("mode %d: lane 0 sum %u (expected %u), %s\n", MODE, h[0], want, cudaGetLastError() == cudaSuccess ? "ok" : "CUDA error");
  return h[0] != want;
}
