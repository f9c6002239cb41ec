// x3_crc_host.h -- host-side construction of the CRC-16/CCITT-FALSE table bank (layout: x3_common.cuh).
#pragma once

#include <stdint.h>

#include "x3_common.cuh"

namespace x3 {

inline uint16_t crc_shift1(uint16_t s) { return (uint16_t)((s & 0x8000) ? ((s << 1) ^ 0x1021) : (s << 1)); }
inline uint16_t crc_mul16(uint16_t a, uint16_t K) {  // a(x) * K(x) mod x^16+x^12+x^5+1
  uint16_t r = 0;
  for (int i = 15; i >= 0; i--) {
    r = crc_shift1(r);
    if ((a >> i) & 1) r ^= K;
  }
  return r;
}
inline uint16_t crc_xpow(unsigned n) {  // x^n mod P
  uint16_t s = 1;
  for (unsigned i = 0; i < n; i++) s = crc_shift1(s);
  return s;
}
inline void build_crc_bank(uint16_t *T /*[kCrcBankEntries3]*/) {
  for (int b = 0; b < 256; b++) {
    uint16_t c = (uint16_t)(b << 8);
    for (int j = 0; j < 8; j++) c = crc_shift1(c);  // b * x^16, the table of crc.rs:22-42
    for (int k = 0; k < 4; k++) {
      T[k * 256 + b] = c;  // b * x^(8k+16)
      for (int j = 0; j < 8; j++) c = crc_shift1(c);
    }
  }
  auto fill_mul = [&](int t, uint16_t K) {
    for (int b = 0; b < 256; b++) {
      T[t * 256 + b] = crc_mul16((uint16_t)b, K);
      T[(t + 1) * 256 + b] = crc_mul16((uint16_t)(b << 8), K);
    }
  };
  fill_mul(4, crc_xpow(4096));                                    // 32 chunks of 16 bytes
  for (int k = 0; k < 5; k++) fill_mul(6 + 2 * k, crc_xpow(128u << k));  // 2^k chunks of 16 bytes
  for (int i = 0; i < kCrcTableEntries; i++) T[kCrcTableEntries + i] = (uint16_t)((T[i] << 8) | (T[i] >> 8));  // swapped bank
  // nibble tables of the strip encoder's constants (layout: x3_common.cuh)
  auto sw = [](uint16_t v) { return (uint16_t)((v << 8) | (v >> 8)); };
  for (int c = 0; c < kCrcMulConsts; c++) {
    const unsigned e = c == 0 ? 0u : c <= 3 ? 256u * c : c <= 10 ? 1024u * (c - 3) : c == kCrcMulX128 ? 128u : 8192u;
    const uint16_t K = crc_xpow(e);
    const int pos[4] = {2, 3, 0, 1};
    for (int t = 0; t < 4; t++)
      for (int nib = 0; nib < 16; nib++)
        T[kCrcBankEntries2 + c * 64 + t * 16 + nib] = sw(crc_mul16(sw((uint16_t)(nib << (4 * pos[t]))), K));
  }
}

}  // namespace x3
