// Dependent-chain latency microbenchmark for sm_100a: cycles per link of the decoder's bit-position chain
// (shift -> count leading zeros -> add) and of candidate replacements.  One warp, one chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu ; run: ./latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNR 16
#define DEF(name, body)                                                                      \
  __global__ void k_##name(uint32_t *out, long long *cyc, uint32_t s0, uint32_t s1) {        \
    uint32_t x = threadIdx.x + s0, lo = s0 * 77u + 5u, hi = s0 * 991u + 7u;                  \
    const uint32_t one = s1, nbk = s1 + 1u;                                                  \
    (void)one; (void)nbk; (void)lo; (void)hi;                                                \
    long long t0 = clock64();                                                                \
    _Pragma("unroll 1") for (int i = 0; i < ITERS; i++) {                                    \
      _Pragma("unroll") for (int r = 0; r < UNR; r++) { body }                               \
    }                                                                                        \
    long long t1 = clock64();                                                                \
    out[threadIdx.x] = x;                                                                    \
    if (threadIdx.x == 0) *cyc = t1 - t0;                                                    \
  }

DEF(iadd, x = x + one;)
DEF(lop3, x = (x ^ lo) & (x | one);)
DEF(shf, x = __funnelshift_l(lo, hi, x);)
DEF(imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(lo));)
DEF(flo, x = __clz(x | one);)
DEF(popc, x = __popc(x | lo);)
DEF(i2f_rz, { float f; asm volatile("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(x | 0x10000u)); x = __float_as_uint(f); })
DEF(prmt, x = __byte_perm(x, lo, 0x5432);)
DEF(isetp_sel, x = (x > lo) ? one : x + 1;)
DEF(vimnmx, x = max(x, lo) ^ one;)
// the decoder's chain: t = funnel(lo,hi,cum); z = clz(t); cum += z + nbk      (cum kept small by & 31)
DEF(chain_flo, { const uint32_t t = __funnelshift_l(lo, hi, x); const uint32_t z = __clz(t | 0x100u); x = (x + z + nbk) & 31u; })
// same with cvt.rz: z = 158 - exponent
DEF(chain_i2f, { const uint32_t t = __funnelshift_l(lo, hi, x); float f; asm volatile("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(t | 0x100u)); const uint32_t e = __float_as_uint(f) >> 23; x = (x + nbk + 158u - e) & 31u; })
// compare ladder for short runs: z = (t < 2^31) + (t < 2^30) + (t < 2^29) (then a fallback, not timed)
DEF(chain_cmp, { const uint32_t t = __funnelshift_l(lo, hi, x); const uint32_t z = (t < 0x80000000u) + (t < 0x40000000u) + (t < 0x20000000u); x = (x + z + nbk) & 31u; })
DEF(lds, { extern __shared__ uint32_t sm[]; x = sm[x & 255u] + one; })

#define RUN(name, shm)                                                                       \
  {                                                                                          \
    k_##name<<<1, 32, shm>>>(d_out, d_cyc, 1, 1);                                            \
    k_##name<<<1, 32, shm>>>(d_out, d_cyc, 1, 1);                                            \
    long long c;                                                                             \
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);                                        \
    printf("%-12s %7.2f cycles per link\n", #name, (double)c / ((double)ITERS * UNR));      \
  }

int main() {
  uint32_t *d_out;
  long long *d_cyc;
  cudaMalloc(&d_out, 4096);
  cudaMalloc(&d_cyc, 8);
  RUN(iadd, 0) RUN(lop3, 0) RUN(shf, 0) RUN(imad, 0) RUN(flo, 0) RUN(popc, 0) RUN(i2f_rz, 0) RUN(prmt, 0)
  RUN(isetp_sel, 0) RUN(vimnmx, 0) RUN(chain_flo, 0) RUN(chain_i2f, 0) RUN(chain_cmp, 0) RUN(lds, 1024)
  return cudaDeviceSynchronize() != cudaSuccess;
}
