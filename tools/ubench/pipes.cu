// Instruction-throughput microbenchmark for sm_100a integer pipes (which ops share the ALU pipe, which run on FMA / XU).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define DEF(name, body)                                                                         \
  __global__ void k_##name(uint32_t *out, uint32_t s0, uint32_t s1) {                           \
    uint32_t a0 = threadIdx.x + s0, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;          \
    uint32_t a4 = a0 * 11 + 1, a5 = a0 * 13 + 1, a6 = a0 * 17 + 2, a7 = a0 * 19 + 3;            \
    const uint32_t one = s1, k = s0;                                                            \
    _Pragma("unroll 1") for (int i = 0; i < ITERS; i++) {                                       \
      _Pragma("unroll") for (int r = 0; r < 4; r++) { body }                                    \
    }                                                                                           \
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;         \
  }
#define ALL8(OP) OP(a0, a1) OP(a1, a2) OP(a2, a3) OP(a3, a4) OP(a4, a5) OP(a5, a6) OP(a6, a7) OP(a7, a0)

#define OP_LOP(x, y) x = (x ^ y) & (x | one);
DEF(lop3, ALL8(OP_LOP))
#define OP_SHF(x, y) x = __funnelshift_l(x, y, k);
DEF(shf, ALL8(OP_SHF))
#define OP_SHL(x, y) x = y << (x & 31);
DEF(shl_var, ALL8(OP_SHL))
#define OP_IADD3(x, y) x = x + y;
DEF(iadd, ALL8(OP_IADD3))
#define OP_IMAD(x, y) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y));
DEF(imad, ALL8(OP_IMAD))
#define OP_IMADHI(x, y) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y));
DEF(imad_hi, ALL8(OP_IMADHI))
#define OP_MULHI(x, y) x = __umulhi(x, y);
DEF(mul_hi, ALL8(OP_MULHI))
#define OP_PRMT(x, y) x = __byte_perm(x, y, 0x5432);
DEF(prmt, ALL8(OP_PRMT))
#define OP_FLO(x, y) x = __clz(y);
DEF(flo_add, ALL8(OP_FLO))
#define OP_BFIND(x, y) asm volatile("bfind.shiftamt.u32 %0, %1;" : "=r"(x) : "r"(y));
DEF(bfind, ALL8(OP_BFIND))
#define OP_POPC(x, y) x = __popc(y);
DEF(popc, ALL8(OP_POPC))
#define OP_SEL(x, y) x = (x > y) ? one : x;
DEF(setp_sel, ALL8(OP_SEL))
#define OP_MNMX(x, y) x = max(x, y) ^ one;
DEF(mnmx, ALL8(OP_MNMX))
#define OP_VADD2(x, y) x = __vadd2(x, y);
DEF(viadd16x2, ALL8(OP_VADD2))
#define OP_VMAX2(x, y) x = __vmaxu2(x, y) ^ one;
DEF(vimnmx16x2, ALL8(OP_VMAX2))
#define OP_DP2A(x, y) x = __dp2a_lo(x, y, one);
DEF(idp2a, ALL8(OP_DP2A))
#define OP_DP4A(x, y) x = __dp4a(x, y, one);
DEF(idp4a, ALL8(OP_DP4A))
#define OP_FFMA(x, y) { float f = __uint_as_float(x); f = fmaf(f, __uint_as_float(one), __uint_as_float(y)); x = __float_as_uint(f); }
DEF(ffma, ALL8(OP_FFMA))
#define OP_I2F(x, y) x = __float_as_uint((float)(int)y);
DEF(i2f, ALL8(OP_I2F))
#define OP_MIX(x, y) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y)); x = (x ^ y) & (x | one);
DEF(imad_plus_lop3, ALL8(OP_MIX))
#define OP_MIX2(x, y) x = __clz(x); x = (x ^ y) & (x | one);
DEF(flo_plus_lop3, ALL8(OP_MIX2))
#define OP_LEA(x, y) x = (x << 3) + y;
DEF(lea, ALL8(OP_LEA))
#define OP_SHR(x, y) x = (y >> 3) ^ one;
DEF(shr_const, ALL8(OP_SHR))
#define OP_SHLC(x, y) x = (y << 3) ^ x;
DEF(shl_const_xor, ALL8(OP_SHLC))


#define OP_IDP_LOP(x, y) x = __dp2a_lo(x, y, one); x = (x ^ y) & (x | one);
DEF(idp2a_plus_lop3, ALL8(OP_IDP_LOP))
#define OP_IDP_IMAD(x, y) x = __dp2a_lo(x, y, one); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y));
DEF(idp2a_plus_imad, ALL8(OP_IDP_IMAD))
#define OP_IDP4_LOP(x, y) x = __dp4a(x, y, one); x = (x ^ y) & (x | one);
DEF(idp4a_plus_lop3, ALL8(OP_IDP4_LOP))
#define OP_I2F_LOP(x, y) x = __float_as_uint((float)(int)y); x = (x ^ y) & (x | one);
DEF(i2f_plus_lop3, ALL8(OP_I2F_LOP))
#define OP_I2F_IMAD(x, y) x = __float_as_uint((float)(int)y); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y));
DEF(i2f_plus_imad, ALL8(OP_I2F_IMAD))
#define OP_FFMA_LOP(x, y) { float f = __uint_as_float(x); f = fmaf(f, __uint_as_float(one), __uint_as_float(y)); x = __float_as_uint(f); } x = (x ^ y) & (x | one);
DEF(ffma_plus_lop3, ALL8(OP_FFMA_LOP))
#define OP_FFMA_IMAD(x, y) { float f = __uint_as_float(x); f = fmaf(f, __uint_as_float(one), __uint_as_float(y)); x = __float_as_uint(f); } asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y));
DEF(ffma_plus_imad, ALL8(OP_FFMA_IMAD))
#define OP_FADD_LOP(x, y) { float f = __uint_as_float(x); f = f + __uint_as_float(y); x = __float_as_uint(f); } x = (x ^ y) & (x | one);
DEF(fadd_plus_lop3, ALL8(OP_FADD_LOP))
#define OP_IMADW(x, y) { unsigned long long w = (unsigned long long)x * y + one; x = (uint32_t)(w >> 32) ^ (uint32_t)w; }
DEF(imad_wide_xor, ALL8(OP_IMADW))
#define OP_HADD2_LOP(x, y) { asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y)); } x = (x ^ y) & (x | one);
DEF(hadd2_plus_lop3, ALL8(OP_HADD2_LOP))

template <class K>
void run(const char *name, K kern, int ops_per_iter, uint32_t *d) {
  int dev_sms = 148;
  dim3 grid(dev_sms * 2), block(512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<grid, block>>>(d, 3, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<grid, block>>>(d, 3, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_instr = (double)grid.x * (block.x / 32) * ITERS * 4.0 * ops_per_iter;
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-16s %8.3f ms  %6.3f source-ops/cycle/SMSP (%d per body)\n", name, ms, warp_instr / cycles / (dev_sms * 4), ops_per_iter);
}

int main() {
  uint32_t *d;
  cudaMalloc(&d, 148 * 2 * 512 * 4);
#define RUN(name, n) run(#name, k_##name, n, d);
  RUN(lop3, 8) RUN(shf, 8) RUN(shl_var, 8) RUN(iadd, 8) RUN(imad, 8) RUN(imad_hi, 8) RUN(mul_hi, 8) RUN(prmt, 8)
  RUN(flo_add, 8) RUN(bfind, 8) RUN(popc, 8) RUN(setp_sel, 8) RUN(mnmx, 8) RUN(viadd16x2, 8) RUN(vimnmx16x2, 8)
  RUN(idp2a, 8) RUN(idp4a, 8) RUN(ffma, 8) RUN(i2f, 8) RUN(imad_plus_lop3, 8) RUN(flo_plus_lop3, 8) RUN(lea, 8)
  RUN(shr_const, 8) RUN(shl_const_xor, 8)
  RUN(idp2a_plus_lop3, 8) RUN(idp2a_plus_imad, 8) RUN(idp4a_plus_lop3, 8) RUN(i2f_plus_lop3, 8) RUN(i2f_plus_imad, 8)
  RUN(ffma_plus_lop3, 8) RUN(ffma_plus_imad, 8) RUN(fadd_plus_lop3, 8) RUN(imad_wide_xor, 8) RUN(hadd2_plus_lop3, 8)
  return 0;
}
