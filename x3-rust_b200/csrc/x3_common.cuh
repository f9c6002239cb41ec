// x3_common.cuh -- shared definitions for the B200 X3 codec kernels.
//
// The per-thread codec logic (block classification, bit packing, bit parsing) lives in
// __host__ __device__ functions so that tests/sim can run the exact same source on the CPU,
// phase by phase, against the oracle before any GPU time is spent.  The cooperative parts
// (scans, look-back, copies) are in the .cu files.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define X3_HD __host__ __device__ __forceinline__
#define X3_D __device__ __forceinline__
#else
#define X3_HD inline
#define X3_D inline
#endif

#if !defined(__CUDACC__)
struct alignas(16) uint4 { uint32_t x, y, z, w; };  // host-only stand-in for the CUDA vector type (tests/sim)
#endif

namespace x3 {

// ---- format constants (x3.rs:136-184) ------------------------------------------------------
constexpr int kFrameHeaderLen = 20;        // FrameHeader::LENGTH, x3.rs:166
constexpr uint32_t kFrameKey = 0x7833;     // "x3", x3.rs:169
constexpr uint32_t kFrameMaxLength = 0x7fe0; // Frame::MAX_LENGTH, x3.rs:145
constexpr int kMaxBlockLen = 60;           // Parameters::MAX_BLOCK_LENGTH, x3.rs:90
constexpr uint32_t kReadBufferSize = 1024 * 24; // X3_READ_BUFFER_SIZE, decodefile.rs:44

// inv_len of RICE0..3 (x3.rs:214,222,236,250) and table offsets (x3.rs:210,218,226,240)
X3_HD uint32_t rice_inv_len(uint32_t code) { return code == 0 ? 16u : code == 1 ? 26u : code == 2 ? 44u : 60u; }
X3_HD uint32_t rice_offset(uint32_t code) { return code == 0 ? 6u : code == 1 ? 11u : code == 2 ? 20u : 28u; }

// ---- portable bit intrinsics ------------------------------------------------------------------
X3_HD uint32_t clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)x);
#else
  return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
X3_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return __builtin_bswap32(x);
#endif
}
// (hi:lo) << s, upper word; s in [0,32]
X3_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_lc(lo, hi, s);
#else
  if (s == 0) return hi;
  if (s >= 32) return lo;
  return (hi << s) | (lo >> (32 - s));
#endif
}
// (hi:lo) >> s, lower word; s in [0,32]
X3_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, s);
#else
  if (s == 0) return lo;
  if (s >= 32) return hi;
  return (lo >> s) | (hi << (32 - s));
#endif
}
// shifts that are defined for any amount: PTX shl/shr clamp the amount to 32 (one SHF), C++ needs the select
X3_HD uint32_t shl_safe(uint32_t v, uint32_t s) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
#else
  return s >= 32 ? 0u : v << s;
#endif
}
X3_HD uint32_t shr_safe(uint32_t v, uint32_t s) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
#else
  return s >= 32 ? 0u : v >> s;
#endif
}
// leading zeros as a shift amount: clz for x != 0; for x == 0 the device returns 0xffffffff (bfind.shiftamt, one
// FLO) and the host 32 -- callers only rely on "x == 0 gives a value >= 32"
X3_HD uint32_t clz_shift(uint32_t x) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
#else
  return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

// zig-zag fold of the first difference: u = d<0 ? -2d-1 : 2d  (closed form of the `offset` indexing of
// x3.rs:207-252, verified against the four tables by tests/test_oracle_golden.py)
// (a widening multiply -- mul.wide.s32 d,2 -> 2d and the sign mask in one IMAD.WIDE -- was measured: much slower)
X3_HD uint32_t fold(int32_t d) { return ((uint32_t)d << 1) ^ (uint32_t)(d >> 31); }
// INV_RICE_CODE[i], x3.rs:200-204
X3_HD int32_t unfold(uint32_t i) { return (int32_t)(i >> 1) ^ -(int32_t)(i & 1u); }

// ---- CRC-16/CCITT-FALSE (crc.rs) --------------------------------------------------------------
// Table bank layout (uint16 entries), built on the host by x3_build_crc_tables():
//   [0..4)   T_k[b] = b * x^(8k+16) mod P, k = 0..3       (slicing-by-4; T_0 is crc.rs:22-42)
//   [4..6)   multiply by x^4096 (lo byte, hi byte)         (lane Horner step: 32 chunks of 16 bytes)
//   [6..16)  multiply by x^(128*2^k), k = 0..4 (lo, hi)    (warp tree combine)
constexpr int kCrcTables = 16;
constexpr int kCrcTableEntries = kCrcTables * 256;
// A second bank of the same 16 tables with every entry byte-swapped follows the first one in device memory; the
// fast encoder uses it (crc16_word_sw / crc16_mulc_sw below).
constexpr int kCrcBankEntries2 = 2 * kCrcTableEntries;
// A third part follows: NIBBLE tables (4 x 16 entries, byte-swapped state form) of "multiply by a constant mod P" for
// the strip encoder, which keeps them in shared memory (128 bytes per constant instead of 1 KB).  Constant c:
//   0: 1 (identity), 1..3: x^(256 c), 4..10: x^(1024 (c-3)), 11: x^128, 12: x^8192.
// Entry [c*64 + t*16 + nib]: the product for a swapped state whose only non-zero nibble is `nib` at bits 4p..4p+3,
// p = 2, 3, 0, 1 for t = 0..3 (the order the lookup code wants: crc16_mul_nib).
constexpr int kCrcMulConsts = 13;
constexpr int kCrcMulEntries = kCrcMulConsts * 64;
constexpr int kCrcMulX128 = 11, kCrcMulX8192 = 12;
constexpr int kCrcBankEntries3 = kCrcBankEntries2 + kCrcMulEntries;

// one 32-bit big-endian word (first stream byte in bits 31..24) through the CRC state
X3_HD uint32_t crc16_word(const uint16_t *T, uint32_t s, uint32_t w) {
  uint32_t v = (s << 16) ^ w;
  return (uint32_t)T[768 + (v >> 24)] ^ (uint32_t)T[512 + ((v >> 16) & 0xff)] ^
         (uint32_t)T[256 + ((v >> 8) & 0xff)] ^ (uint32_t)T[v & 0xff];
}
// one 16-bit big-endian halfword
X3_HD uint32_t crc16_half(const uint16_t *T, uint32_t s, uint32_t h) {
  uint32_t v = (s ^ h) & 0xffffu;
  return (uint32_t)T[256 + (v >> 8)] ^ (uint32_t)T[v & 0xff];
}
X3_HD uint32_t crc16_byte(const uint16_t *T, uint32_t s, uint32_t b) {
  return ((s << 8) & 0xffffu) ^ (uint32_t)T[((s >> 8) ^ b) & 0xff];
}
// Table-free variants for kernels that are bound by shared-memory lookups.  With P = x^16+x^12+x^5+1,
// a*x^16 mod P for a 16-bit a is r = low16(q<<12 ^ q<<5 ^ q), where the quotient q solves q = a ^ q>>4 ^ q>>11,
// i.e. q = a ^ a>>4 ^ a>>8 ^ a>>12 ^ a>>11 (the q>>15 terms cancel).  Ten ALU operations per 16 bits.
X3_HD uint32_t crc16_mulx16(uint32_t a) {  // a < 2^16
  const uint32_t q = a ^ (a >> 4) ^ (a >> 8) ^ (a >> 12) ^ (a >> 11);
  return ((q << 12) ^ (q << 5) ^ q) & 0xffffu;
}
X3_HD uint32_t crc16_word_alu(uint32_t s, uint32_t w) {  // same result as crc16_word
  const uint32_t s1 = crc16_mulx16((s ^ (w >> 16)) & 0xffffu) ^ (w & 0xffffu);
  return crc16_mulx16(s1);
}
X3_HD uint32_t crc16_half_alu(uint32_t s, uint32_t h) { return crc16_mulx16((s ^ h) & 0xffffu); }
// multiply a 16-bit state by the constant whose (lo,hi) table pair starts at table index t
X3_HD uint32_t crc16_mulc(const uint16_t *T, int t, uint32_t s) {
  return (uint32_t)T[t * 256 + (s & 0xff)] ^ (uint32_t)T[(t + 1) * 256 + ((s >> 8) & 0xff)];
}

// ---- "swapped" form: the CRC state is kept byte-swapped (s_sw = bswap16(s)) and data words are taken as they lie in
// memory (little-endian load of stream bytes), with the byte-swapped table bank T2.  No bswap per word, and on the
// device the table addresses come from IDP.4A (byte * 2 + table base in one FMA-pipe instruction) instead of
// shift/mask pairs on the ALU pipe, which is the encoder's bottleneck.
#if defined(__CUDA_ARCH__)
template <int BYTE_OFF>
__device__ __forceinline__ uint32_t lds_u16_off(uint32_t addr) {
  uint32_t r;
  asm("ld.shared.u16 %0, [%1+%2];" : "=r"(r) : "r"(addr), "n"(BYTE_OFF));
  return r;
}
#endif
// stream bytes b0..b3 = bits 0..7, 8..15, 16..23, 24..31 of w_le
X3_HD uint32_t crc16_word_sw(const uint16_t *T2, uint32_t s_sw, uint32_t w_le) {
  const uint32_t v = s_sw ^ w_le;
#if defined(__CUDA_ARCH__)
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(T2);
  return lds_u16_off<1536>(__dp4a(v, 0x00000002u, tb)) ^ lds_u16_off<1024>(__dp4a(v, 0x00000200u, tb)) ^
         lds_u16_off<512>(__dp4a(v, 0x00020000u, tb)) ^ lds_u16_off<0>(__dp4a(v, 0x02000000u, tb));
#else
  return (uint32_t)T2[768 + (v & 0xff)] ^ (uint32_t)T2[512 + ((v >> 8) & 0xff)] ^
         (uint32_t)T2[256 + ((v >> 16) & 0xff)] ^ (uint32_t)T2[v >> 24];
#endif
}
// multiply a swapped 16-bit state (bits above 15 are ignored) by the constant of table pair TBL
template <int TBL>
X3_HD uint32_t crc16_mulc_sw(const uint16_t *T2, uint32_t s_sw) {
#if defined(__CUDA_ARCH__)
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(T2);
  return lds_u16_off<TBL * 512>(__dp4a(s_sw, 0x00000200u, tb)) ^ lds_u16_off<(TBL + 1) * 512>(__dp4a(s_sw, 0x00000002u, tb));
#else
  return (uint32_t)T2[TBL * 256 + ((s_sw >> 8) & 0xff)] ^ (uint32_t)T2[(TBL + 1) * 256 + (s_sw & 0xff)];
#endif
}
X3_HD uint32_t bswap16(uint32_t s) { return ((s & 0xffu) << 8) | ((s >> 8) & 0xffu); }
// swapped 16-bit state times the constant whose 64-entry nibble table is N (shared memory on the device): the four
// nibbles are spread into the bytes of one register (two LOP3 and an IMAD), the four table addresses are IDP.4A
X3_HD uint32_t crc16_mul_nib(const uint16_t *N, uint32_t s_sw) {
  const uint32_t x = (s_sw & 0xf0f0u) * 4096u + (s_sw & 0x0f0fu);   // bytes: n0, n2, n1, n3
#if defined(__CUDA_ARCH__)
  const uint32_t tb = (uint32_t)__cvta_generic_to_shared(N);
  return lds_u16_off<64>(__dp4a(x, 0x00000002u, tb)) ^ lds_u16_off<0>(__dp4a(x, 0x00000200u, tb)) ^
         lds_u16_off<96>(__dp4a(x, 0x00020000u, tb)) ^ lds_u16_off<32>(__dp4a(x, 0x02000000u, tb));
#else
  return (uint32_t)N[32 + (x & 15u)] ^ (uint32_t)N[(x >> 8) & 15u] ^ (uint32_t)N[48 + ((x >> 16) & 15u)] ^ (uint32_t)N[16 + (x >> 24)];
#endif
}

// frame header bytes 0..16 -> header CRC (encoder.rs:153); words are big-endian images of the bytes
X3_HD uint32_t header_crc(const uint16_t *T, uint32_t id, uint32_t num_samples, uint32_t payload_len) {
  uint32_t s = 0xffffu;
  s = crc16_word(T, s, (kFrameKey << 16) | ((id & 0xff) << 8) | (id & 0xff));
  s = crc16_word(T, s, ((num_samples & 0xffff) << 16) | (payload_len & 0xffff));
  s = crc16_word(T, s, 0u);
  s = crc16_word(T, s, 0u);
  return s;
}

// ---- encoder parameters, preprocessed on the host ------------------------------------------------
struct CodecParams {
  uint32_t block_len;
  uint32_t spf;            // samples per frame = block_len * blocks_per_frame
  uint32_t codes[3];
  uint32_t thresholds[3];
};

}  // namespace x3
