// x3_lookback.cuh -- single-pass decoupled look-back over 62-bit values (one 64-bit status word per
// item: 2 flag bits + value), used for frame byte offsets (encoder) and frame ordinals (index scan).
#pragma once

#include <stdint.h>

namespace x3 {

constexpr unsigned long long kFlagAgg = 1ull << 62;     // value = this item's own contribution
constexpr unsigned long long kFlagPrefix = 2ull << 62;  // value = inclusive prefix through this item
constexpr unsigned long long kValueMask = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

// Called by one full warp.  Item `idx` has already published kFlagAgg|own (or kFlagPrefix|own for idx 0).
// Returns the exclusive prefix (sum of items 0..idx-1) in every lane and publishes the inclusive prefix.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *status, unsigned long long idx,
                                                                unsigned long long own) {
  const int lane = threadIdx.x & 31;
  unsigned long long excl = 0;
  if (idx == 0) return 0;
  long long at = (long long)idx - 1;
  for (;;) {
    const long long j = at - lane;
    const unsigned long long v = j >= 0 ? ld_status(status + j) : kFlagPrefix;
    const unsigned flag = (unsigned)(v >> 62);
    const unsigned pmask = __ballot_sync(0xffffffffu, flag == 2u);
    const unsigned zmask = __ballot_sync(0xffffffffu, flag == 0u);
    unsigned upto;  // lanes [0, upto] contribute
    if (pmask) {
      upto = (unsigned)__ffs((int)pmask) - 1u;
      if (zmask & ((2u << upto) - 1u)) { __nanosleep(32); continue; }
    } else {
      if (zmask) { __nanosleep(32); continue; }
      upto = 31u;
    }
    unsigned long long val = (unsigned)lane <= upto ? (v & kValueMask) : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
    excl += val;
    if (pmask) break;
    at -= 32;
  }
  if (lane == 0) st_status(status + idx, kFlagPrefix | ((excl + own) & kValueMask));
  return excl;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

}  // namespace x3
