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
//
// The window is WIDE: every lane reads 8 predecessors, 256 per round trip.  With a few hundred items in
// flight (one per resident CTA) whose prefixes are all still pending, a 32-wide window needs ~10 dependent
// round trips per item, which is slower than items are produced -- the look-back depth then grows until it
// spans everything in flight.  256 entries per trip covers the in-flight set in one or two.
constexpr int kLookG = 8;
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *status, unsigned long long idx,
                                                                unsigned long long own) {
  const int lane = threadIdx.x & 31;
  unsigned long long excl = 0;
  if (idx == 0) return 0;
  long long at = (long long)idx - 1;  // nearest predecessor of this round
  for (;;) {
    // lane handles items at - 8*lane - g, g = 0..7 (nearest first)
    unsigned long long v[kLookG];
#pragma unroll
    for (int g = 0; g < kLookG; g++) {
      const long long j = at - (long long)(kLookG * lane + g);
      v[g] = j >= 0 ? ld_status(status + j) : kFlagPrefix;  // virtual prefix 0 before item 0
    }
    // per lane: sum up to and including its first prefix.  An unpublished entry that comes before the lane's
    // first prefix is needed, so the lane polls that one word until it appears (entries older than any
    // published prefix are always published themselves, so lanes past the warp's first prefix never wait).
    unsigned long long sum = 0;
    bool has_prefix = false;
#pragma unroll
    for (int g = 0; g < kLookG; g++) {
      if (!has_prefix) {
        while ((v[g] >> 62) == 0ull) {
          __nanosleep(64);
          v[g] = ld_status(status + (at - (long long)(kLookG * lane + g)));
        }
        sum += v[g] & kValueMask;
        if ((unsigned)(v[g] >> 62) == 2u) has_prefix = true;
      }
    }
    __syncwarp();
    const unsigned pmask = __ballot_sync(0xffffffffu, has_prefix);
    const unsigned upto = pmask ? (unsigned)__ffs((int)pmask) - 1u : 31u;  // lanes [0, upto] contribute
    unsigned long long val = (unsigned)lane <= upto ? sum : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
    excl += val;
    if (pmask) break;
    at -= 32 * kLookG;
  }
  if (lane == 0) st_status(status + idx, kFlagPrefix | ((excl + own) & kValueMask));
  return excl;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

}  // namespace x3
