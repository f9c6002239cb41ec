// x3_enc_core.cuh -- per-block encoder logic (one thread = one block of <= 60 samples).
//
// Replaces, per block: encoder::diff (encoder.rs:222-225), x3_encode_block (encoder.rs:289-315),
// encode_rice_block (:233-267), encode_bfp_block (:269-276), encode_literal (:278-285) and the
// BitPacker (bitpacker.rs:142-163).  Host+device so that tests/sim can run it on the CPU.
//
// Bit layout.  A frame payload is one MSB-first bit string (bitpacker.rs).  The kernel keeps it in
// shared memory as a byte image addressed as 32-bit words.  A block whose bit range is
// [o, o+n) writes every word it owns exactly once:
//   * words that start inside the block and end inside it  -> plain store while packing
//   * the word holding its first bit when o % 32 != 0       -> "head" partial, published to Hs[b]
//   * the word holding its last bit when (o+n) % 32 != 0    -> "tail" partial Ts[b]; after a barrier
//     the owner ORs in the heads of the following blocks that start in that word (at most two when
//     block_len is 20; a loop handles shorter blocks).
// No atomics and no pre-zeroing are needed; every output word has exactly one writer.
#pragma once

#include "x3_common.cuh"

namespace x3 {

enum BlockKind : uint32_t { kRice = 0, kBfp = 1, kLiteral = 2 };

struct BlockMode {
  uint32_t kind;   // BlockKind
  uint32_t k;      // Rice: number of suffix bits (RiceCode.nsubs); BFP: bits per sample - 1 (num_bits)
  uint32_t hdr;    // header value: ftype+1 in 2 bits (encoder.rs:250) or num_bits / 15 in 6 bits (:270,:280)
  uint32_t stat;   // index into stats[6] (encoder.rs:266,275,284)
};

// The selection rule of encoder.rs:304-314 + :241-247.  It is a threshold rule on max|d|.
X3_HD BlockMode classify(uint32_t max_abs, const CodecParams &P) {
  BlockMode m;
  if (max_abs <= P.thresholds[2]) {
    uint32_t ftype = (max_abs > P.thresholds[0]) + (max_abs > P.thresholds[1]) + (max_abs > P.thresholds[2]);
    m.kind = kRice;
    m.k = P.codes[ftype];
    m.hdr = ftype + 1;
    m.stat = m.k;
  } else {
    uint32_t nb = 32u - clz32(max_abs);  // count_bits, encoder.rs:229-231
    if (nb >= 15) {
      m.kind = kLiteral; m.k = 15; m.hdr = 15; m.stat = 5;
    } else {
      m.kind = kBfp; m.k = nb; m.hdr = nb; m.stat = 4;
    }
  }
  return m;
}

// bits a block of `len` samples occupies, given its mode and (for Rice) the sum of u>>k
X3_HD uint32_t block_bits(const BlockMode &m, uint32_t len, uint32_t sum_q) {
  if (m.kind == kRice) return 2u + len * (m.k + 1u) + sum_q;
  if (m.kind == kBfp) return 6u + len * (m.k + 1u);
  return 6u + len * 16u;
}

// MSB-first bit sink over the shared-memory byte image (see file comment).
struct BitSink {
  uint64_t acc;
  uint32_t cnt;       // valid low bits of acc not yet stored
  uint32_t *dst;      // where the next completed word goes
  uint32_t *nxt;      // the word after that
  uint32_t *head;     // Hs[b]

  X3_HD void init(uint32_t bit_off, uint32_t *out_words, uint32_t *head_slot) {
    acc = 0;
    cnt = bit_off & 31u;
    head = head_slot;
    uint32_t *w = out_words + (bit_off >> 5);
    nxt = w + 1;
    if (cnt) {
      dst = head_slot;
    } else {
      dst = w;
      *head_slot = 0u;
    }
  }
  // append the low n bits of v (v < 2^n, n <= 32)
  X3_HD void put(uint32_t v, uint32_t n) {
    acc = (acc << n) | (uint64_t)v;
    cnt += n;
  }
  // store one completed word if there is one (call often enough that cnt never exceeds 64)
  X3_HD void flush() {
    if (cnt >= 32u) {
      cnt -= 32u;
      *dst = bswap32((uint32_t)(acc >> cnt));
      dst = nxt;
      nxt = nxt + 1;
    }
  }
  // end of block: returns the tail partial word image (0 if none) and sets has_tail
  X3_HD uint32_t finish(bool &has_tail) {
    flush();
    has_tail = false;
    if (cnt == 0u) return 0u;
    uint32_t img = bswap32((uint32_t)(acc << (32u - cnt)));
    if (dst == head) {  // never completed a word and started mid-word: all bits belong to a predecessor's word
      *head = img;
      return 0u;
    }
    has_tail = true;
    return img;
  }
};

// ---------------------------------------------------------------------------------------------
// Generic block path: any block_len <= 60, any codes / thresholds.  Reads the samples three times
// from (shared) memory; used for non-default Parameters and for short tail blocks.
// `s` points at the frame's samples, block covers s[start .. start+len), predecessor s[start-1].
// ---------------------------------------------------------------------------------------------
X3_HD BlockMode block_measure_generic(const int16_t *s, uint32_t start, uint32_t len, const CodecParams &P,
                                      uint32_t &nbits) {
  uint32_t maxu = 0;
  int32_t prev = s[start - 1];
  for (uint32_t i = 0; i < len; i++) {
    int32_t x = s[start + i];
    uint32_t u = fold(x - prev);
    prev = x;
    maxu = u > maxu ? u : maxu;
  }
  BlockMode m = classify((maxu + 1u) >> 1, P);
  uint32_t sum_q = 0;
  if (m.kind == kRice) {
    prev = s[start - 1];
    for (uint32_t i = 0; i < len; i++) {
      int32_t x = s[start + i];
      sum_q += fold(x - prev) >> m.k;
      prev = x;
    }
  }
  nbits = block_bits(m, len, sum_q);
  return m;
}

template <class Sink>
X3_HD void block_pack_generic(const int16_t *s, uint32_t start, uint32_t len, const BlockMode &m, Sink &sink) {
  int32_t prev = s[start - 1];
  if (m.kind == kRice) {
    sink.put(m.hdr, 2);
    sink.flush();
    const uint32_t k = m.k, marker = 1u << k, mask = marker - 1u;
    for (uint32_t i = 0; i < len; i++) {
      int32_t x = s[start + i];
      uint32_t u = fold(x - prev);
      prev = x;
      uint32_t q = u >> k;
      while (q >= 32u) {  // only reachable with thresholds far beyond the reference's table domains
        sink.put(0u, 32u);
        sink.flush();
        q -= 32u;
      }
      sink.put(0u, q);
      sink.flush();
      sink.put(marker | (u & mask), k + 1u);
      sink.flush();
    }
  } else if (m.kind == kBfp) {
    sink.put(m.hdr, 6);
    sink.flush();
    const uint32_t w = m.k + 1u, mask = (1u << w) - 1u;
    for (uint32_t i = 0; i < len; i++) {
      int32_t x = s[start + i];
      sink.put((uint32_t)(x - prev) & mask, w);
      sink.flush();
      prev = x;
    }
  } else {
    sink.put(15u, 6);
    sink.flush();
    for (uint32_t i = 0; i < len; i++) {
      sink.put((uint32_t)(uint16_t)s[start + i], 16);
      sink.flush();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path: Parameters::default() (block_len 20, codes 0/1/3, thresholds 3/8/20) and a block of
// 19 or 20 samples.  One pass over shared memory; the folded differences stay in registers.
// ---------------------------------------------------------------------------------------------
constexpr int kFastBL = 20;

struct FastBlock {
  uint32_t u[kFastBL];  // folded first differences
  int32_t pred;         // sample before the block
};

X3_HD bool params_are_default(const CodecParams &P) {
  return P.block_len == 20 && P.codes[0] == 0 && P.codes[1] == 1 && P.codes[2] == 3 && P.thresholds[0] == 3 &&
         P.thresholds[1] == 8 && P.thresholds[2] == 20;
}

// d + a.lo16 * b.byte0 + a.hi16 * b.byte1 (all signed): IDP.2A on sm_100a, which issues on the FMA pipe.  The
// encoder is bound by the ALU pipe (LOP3/SHF/PRMT/ISETP), so sample extraction and differencing go through here.
X3_HD int32_t dp2a_s16s8(uint32_t a, uint32_t b, int32_t c) {
#if defined(__CUDA_ARCH__)
  return __dp2a_lo((int)a, (int)b, c);
#else
  return c + (int32_t)(int16_t)(a & 0xffffu) * (int32_t)(int8_t)(b & 0xffu) +
         (int32_t)(int16_t)(a >> 16) * (int32_t)(int8_t)((b >> 8) & 0xffu);
#endif
}
X3_HD uint32_t mulhi_add(uint32_t a, uint32_t b, uint32_t c) {  // IMAD.HI.U32: hi32(a*b) + c
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b) + c;
#else
  return (uint32_t)(((uint64_t)a * b) >> 32) + c;
#endif
}

// Reads s[start-1 .. start+20] as eleven aligned 32-bit words (start is odd: 1 + 20*b); sample 19 is ignored
// (u = 0) when len == 19.  `neg1` is -1 in a register the compiler cannot see through (a kernel argument), so that
// ~p is computed as p * neg1 + neg1 on the FMA pipe.
// With p = 2d: fold(d) = max(p, ~p)  (p >= 0: 2d > -2d-1; p < 0: -2d-1 > 2d), one IDP/IMAD pair + one VIMNMX.
X3_HD BlockMode block_measure_fast(const int16_t *s, uint32_t start, uint32_t len, FastBlock &fb, uint32_t &nbits,
                                   int32_t neg1) {
  uint32_t W[kFastBL / 2 + 1];
#if defined(__CUDA_ARCH__)
  {
    const uint2 *w2 = reinterpret_cast<const uint2 *>(s + (start - 1));  // byte offset 40*b: 8-byte aligned
#pragma unroll
    for (int j = 0; j < kFastBL / 4; j++) {
      const uint2 v = w2[j];
      W[2 * j] = v.x;
      W[2 * j + 1] = v.y;
    }
    W[kFastBL / 2] = reinterpret_cast<const uint32_t *>(s + (start - 1))[kFastBL / 2];
  }
#else
  for (int j = 0; j <= kFastBL / 2; j++)
    W[j] = (uint32_t)(uint16_t)s[start - 1 + 2 * j] | ((uint32_t)(uint16_t)s[start + 2 * j] << 16);
#endif
  constexpr uint32_t kHiMinusLo = 0x02feu;  // byte0 = -2, byte1 = +2
  constexpr uint32_t kMinusHi = 0xfe00u;    // byte0 =  0, byte1 = -2
  constexpr uint32_t kPlusLo = 0x0002u;     // byte0 = +2, byte1 =  0
  fb.pred = dp2a_s16s8(W[0], 0x0001u, 0);   // s[start-1]
  uint32_t maxu = 0, sumu = 0;
#pragma unroll
  for (int j = 0; j < kFastBL / 2; j++) {
    const int32_t p0 = dp2a_s16s8(W[j], kHiMinusLo, 0);                                   // 2*(s[2j+1] - s[2j])
    const int32_t p1 = dp2a_s16s8(W[j + 1], kPlusLo, dp2a_s16s8(W[j], kMinusHi, 0));      // 2*(s[2j+2] - s[2j+1])
    const int32_t n0 = p0 * neg1 + neg1, n1 = p1 * neg1 + neg1;                           // ~p
    uint32_t u0 = (uint32_t)(p0 > n0 ? p0 : n0), u1 = (uint32_t)(p1 > n1 ? p1 : n1);
    if (j == kFastBL / 2 - 1 && len != (uint32_t)kFastBL) u1 = 0;
    fb.u[2 * j] = u0;
    fb.u[2 * j + 1] = u1;
    const uint32_t m01 = u0 > u1 ? u0 : u1;
    maxu = m01 > maxu ? m01 : maxu;
    sumu += u0 + u1;
  }
  const uint32_t max_abs = (maxu + 1u) >> 1;
  BlockMode m;
  uint32_t sum_q = sumu;
  if (max_abs <= 20u) {
    uint32_t ftype = (max_abs > 3u) + (max_abs > 8u);
    m.kind = kRice;
    m.k = ftype == 0 ? 0u : (ftype == 1 ? 1u : 3u);
    m.hdr = ftype + 1;
    m.stat = m.k;
    if (m.k) {
      // sum of u >> k as hi32(u * 2^(32-k)) accumulated by IMAD.HI (FMA pipe)
      const uint32_t mult = 0x80000000u >> (m.k - 1u);
      sum_q = 0;
#pragma unroll
      for (int i = 0; i < kFastBL; i++) sum_q = mulhi_add(fb.u[i], mult, sum_q);
    }
  } else {
    uint32_t nb = 32u - clz32(max_abs);
    if (nb >= 15) { m.kind = kLiteral; m.k = 15; m.hdr = 15; m.stat = 5; }
    else { m.kind = kBfp; m.k = nb; m.hdr = nb; m.stat = 4; }
  }
  nbits = block_bits(m, len, sum_q);
  return m;
}

// shared-memory word address: a 32-bit shared-window address on the device (so address arithmetic is one
// instruction), a plain pointer in the CPU simulation
#if defined(__CUDA_ARCH__)
typedef uint32_t smaddr_t;
X3_HD smaddr_t sm_addr(const uint32_t *p) { return (smaddr_t)__cvta_generic_to_shared(p); }
X3_HD void sm_store_if(bool pred, smaddr_t a, uint32_t v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u32 [%0], %1; }" ::"r"(a), "r"(v), "r"((uint32_t)pred) : "memory");
}
X3_HD smaddr_t sm_advance(smaddr_t a, uint32_t words) { return a + 4u * words; }
X3_HD void sm_or(smaddr_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#else
typedef uint32_t *smaddr_t;
X3_HD smaddr_t sm_addr(uint32_t *p) { return p; }
X3_HD void sm_store_if(bool pred, smaddr_t a, uint32_t v) { if (pred) *a = v; }
X3_HD smaddr_t sm_advance(smaddr_t a, uint32_t words) { return a + words; }
X3_HD void sm_or(smaddr_t a, uint32_t v) { *a |= v; }
#endif

// MSB-first bit sink of the fast kernel.  Words are written while packing to their final position in the image by
// the block that holds their LAST bit (a plain store of the whole word; the bits before the block's first bit are
// zero in it).  A block's last, partial word is not stored: finish_tail() returns it (zero padded) and the caller
// ORs it into place with one shared-memory reduction after every plain store is done -- by then the word has been
// stored by the block that completes it, or zeroed by the caller if it is the frame's last.  No second pointer and
// no select: flush() is a compare, a funnel shift, a byte swap, a predicated store and two updates.
struct FastSink {
  uint64_t acc;
  uint32_t cnt;
  smaddr_t dst;
  X3_HD void init(uint32_t bit_off, uint32_t *out_words) {
    acc = 0;
    cnt = bit_off & 31u;
    dst = sm_addr(out_words + (bit_off >> 5));
  }
  X3_HD void put(uint32_t v, uint32_t n) {  // v < 2^n, n <= 32, cnt + n <= 64
    acc = (acc << n) | (uint64_t)v;
    cnt += n;
  }
  X3_HD void flush() {
    const uint32_t w = bswap32((uint32_t)(acc >> (cnt & 31u)));   // cnt in [32,63]: acc >> (cnt - 32)
#if defined(__CUDA_ARCH__)
    // one compare; the store and the pointer step are both predicated on it
    asm volatile("{ .reg .pred p; setp.ge.u32 p, %1, 32; @p st.shared.u32 [%0], %2; @p add.u32 %0, %0, 4; }"
                 : "+r"(dst)
                 : "r"(cnt), "r"(w)
                 : "memory");
#else
    const uint32_t full = cnt >> 5;         // 0 or 1
    sm_store_if(full != 0u, dst, w);
    dst = sm_advance(dst, full);
#endif
    cnt &= 31u;
  }
  X3_HD uint32_t finish_tail() {            // the word at dst, valid if cnt != 0 afterwards
    flush();
    return cnt ? bswap32((uint32_t)(acc << (32u - cnt))) : 0u;
  }
};

// Fast-path packer (no finish: the caller takes the tail).  Rice codewords are at most 10 bits at the default
// thresholds, so three of them are first merged in a 32-bit register -- Horner with the powers of two 2^len as
// multipliers, i.e. two IMADs on the FMA pipe instead of shifts and ORs on the ALU pipe -- and appended with one
// 64-bit shift; BFP / literal samples go in pairs.
template <class Sink>
X3_HD void block_pack_fast(const FastBlock &fb, uint32_t len, const BlockMode &m, Sink &sink) {
  const bool full = len == (uint32_t)kFastBL;
  if (m.kind == kRice) {
    sink.put(m.hdr, 2);
    const uint32_t k = m.k, marker = 1u << k, mask = marker - 1u, k1 = k + 1u, K1 = 2u << k, n3 = 3u * k1;
#pragma unroll
    for (int t = 0; t < 6; t++) {
      const uint32_t u0 = fb.u[3 * t], u1 = fb.u[3 * t + 1], u2 = fb.u[3 * t + 2];
      const uint32_t q0 = u0 >> k, q1 = u1 >> k, q2 = u2 >> k;
      const uint32_t b0 = marker | (u0 & mask), b1 = marker | (u1 & mask), b2 = marker | (u2 & mask);
      const uint32_t v = (b0 * (K1 << q1) + b1) * (K1 << q2) + b2;   // b0 : b1 : b2 with lengths q+k1
      sink.put(v, q0 + q1 + q2 + n3);
      sink.flush();
    }
    {
      const uint32_t u0 = fb.u[18], u1 = fb.u[19];
      const uint32_t q0 = u0 >> k, q1 = u1 >> k;
      const uint32_t b0 = marker | (u0 & mask), b1 = marker | (u1 & mask);
      const uint32_t p1 = full ? (K1 << q1) : 1u, c1 = full ? b1 : 0u, l1 = full ? q1 + k1 : 0u;
      sink.put(b0 * p1 + c1, q0 + k1 + l1);
    }
  } else if (m.kind == kBfp) {
    sink.put(m.hdr, 6);
    sink.flush();
    const uint32_t w = m.k + 1u, mask = (1u << w) - 1u;  // w <= 15
#pragma unroll
    for (int t = 0; t < 10; t++) {
      const uint32_t d0 = (uint32_t)unfold(fb.u[2 * t]) & mask, d1 = (uint32_t)unfold(fb.u[2 * t + 1]) & mask;
      if (t < 9 || full) sink.put((d0 << w) | d1, 2u * w);
      else sink.put(d0, w);
      sink.flush();
    }
  } else {
    sink.put(15u, 6);
    sink.flush();
    int32_t x = fb.pred;
#pragma unroll
    for (int t = 0; t < 10; t++) {
      x += unfold(fb.u[2 * t]);
      const uint32_t s0 = (uint32_t)x & 0xffffu;
      x += unfold(fb.u[2 * t + 1]);
      if (t < 9 || full) sink.put((s0 << 16) | ((uint32_t)x & 0xffffu), 32);
      else sink.put(s0, 16);
      sink.flush();
    }
  }
}

// payload bytes for a payload of total_bits bits: pad to a byte, then to an even length
// (BitPacker::word_align, bitpacker.rs:124-132; the payload starts at an even stream position)
X3_HD uint32_t payload_bytes(uint32_t total_bits) { return ((total_bits + 15u) >> 4) << 1; }

}  // namespace x3
