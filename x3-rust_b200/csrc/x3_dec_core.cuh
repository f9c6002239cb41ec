// x3_dec_core.cuh -- per-frame decoder logic (one thread = one frame).
//
// Replaces decoder::decode_frame (decoder.rs:36-58), decode_block (:132-145), the Rice / BFP / literal
// block decoders (:147-235), BitReader (bitreader.rs:29-176) and the payload CRC check of
// decodefile.rs:93-103.  Host+device so tests/sim can run the same source on the CPU.
//
// Three paths:
//  * decode_frame_fast  -- Parameters::default() streams with frames of a multiple of 80 samples.  64-bit
//    shifting bit window refilled once per three Rice codes, zero runs from the exponent of a float conversion
//    (see the table comment below), output staged 160 bytes at a time so every lane writes whole 32-byte sectors.
//    The payload CRC is checked by a separate kernel (crc16_fold below, x3_decode.cu) beside the decode, as
//    decodefile.rs:93-103 does before it.  It never trusts a frame it cannot prove well formed: any zero run the
//    32-bit peek cannot see the end of, out-of-range Rice index or bad BFP header makes it give up ...
//  * decode_frame_generic -- every frame the tuned path does not cover (other Parameters, short last frame,
//    unaligned output): same reader, any block_len / codes, closed-form inverse fold; gives up the same way ...
//  * decode_frame_exact -- ... and the frame is decoded again by a literal restatement of the reference's
//    word-structured BitReader, which reproduces its behaviour on malformed payloads bit for bit
//    (short-tail refill, one-word look-ahead in count_zero_bits, zero fill past the end).
#pragma once

#include "x3_common.cuh"

namespace x3 {

enum : int {
  kDecOk = 0,
  kDecErrOutOfBoundsInverse = -2,  // X3Error::OutOfBoundsInverse
  kDecErrPayloadLen = -9,          // X3Error::FrameHeaderInvalidPayloadLen
  kDecErrPayloadCrc = -11,         // X3Error::FrameHeaderInvalidPayloadCRC
  kDecErrInvalidBpf = -13,         // X3Error::FrameDecodeInvalidBPF
  kDecErrNoSpace = -15,            // output capacity exceeded
  kDecErrPanic = -104,             // the reference would panic (samples == 0, payload shorter than 2 bytes)
  kDecRetryExact = 1               // fast path gave up; not an error
};

// ------------------------------------------------------------------------------------------------
// Exact path: bitreader.rs restated over a byte pointer.
// ------------------------------------------------------------------------------------------------
struct ExactReader {
  const uint8_t *a;
  uint32_t len, idx;
  uint32_t leading_word, rem_bit;

  X3_HD void read_word(uint32_t at, uint32_t &word, uint32_t &nbytes) const {  // bitreader.rs:29-48
    if (len - at >= 4u) {
      word = ((uint32_t)a[at] << 24) | ((uint32_t)a[at + 1] << 16) | ((uint32_t)a[at + 2] << 8) | (uint32_t)a[at + 3];
      nbytes = 4;
    } else {
      const uint32_t r = len - at;
      uint32_t w = 0;
      if (r >= 1) w |= (uint32_t)a[at] << 24;
      if (r >= 2) w |= (uint32_t)a[at + 1] << 16;
      if (r == 3) w |= (uint32_t)a[at + 2] << 8;
      word = w;
      nbytes = r;
    }
  }
  X3_HD void init(const uint8_t *arr, uint32_t n) {  // bitreader.rs:65-74
    a = arr; len = n;
    uint32_t w, nb;
    read_word(0, w, nb);
    idx = nb; leading_word = w; rem_bit = nb * 8u;
  }
  X3_HD void get_next() {  // bitreader.rs:149-164
    if (idx >= len) { leading_word = 0; rem_bit = 0; return; }
    uint32_t w, nb;
    read_word(idx, w, nb);
    leading_word = w; idx += nb; rem_bit = nb * 8u;
  }
  X3_HD void inc_bits(uint32_t n) {  // bitreader.rs:77-92
    if (n < rem_bit) {
      leading_word = shl_safe(leading_word, n);
      rem_bit -= n;
    } else if (n > rem_bit) {
      const uint32_t rem = n - rem_bit;
      get_next();
      rem_bit = 32u - rem;
      leading_word = shl_safe(leading_word, rem);
    } else {
      get_next();
    }
  }
  X3_HD uint32_t read_nbits(uint32_t n) {  // bitreader.rs:105-118
    if (n <= rem_bit) {
      const uint32_t r = leading_word >> (32u - n);
      inc_bits(n);
      return r;
    }
    const uint32_t rem = n - rem_bit;
    uint32_t r = leading_word >> (32u - n);
    inc_bits(rem_bit);
    r |= leading_word >> (32u - rem);
    inc_bits(rem);
    return r;
  }
  X3_HD uint32_t count_zero_bits() {  // bitreader.rs:127-139
    uint32_t count = clz32(leading_word);
    if (count > rem_bit) {
      if (idx < len) {
        uint32_t w, nb;
        read_word(idx, w, nb);
        count = rem_bit + clz32(w);
      } else {
        count = rem_bit;
      }
    }
    inc_bits(count);
    return count;
  }
};

// decoder.rs:36-58 with :132-235.  `out` gets `samples` values (2-byte stores).
X3_HD int decode_frame_exact(const uint8_t *payload, uint32_t payload_len, int16_t *out, uint32_t samples,
                             const CodecParams &P) {
  if (payload_len < 2u || samples == 0u) return kDecErrPanic;
  int32_t lw = (int16_t)(((uint32_t)payload[0] << 8) | (uint32_t)payload[1]);  // decoder.rs:42
  uint32_t p_wav = 0;
  out[p_wav++] = (int16_t)lw;
  ExactReader br;
  br.init(payload + 2, payload_len - 2u);
  uint32_t remaining = samples - 1u;
  while (remaining > 0u) {
    const uint32_t bl = remaining < P.block_len ? remaining : P.block_len;  // decoder.rs:50
    const uint32_t ftype = br.read_nbits(2);                                // decoder.rs:138
    if (ftype == 0u) {                                                      // decode_bpf_block :209-235
      const uint32_t nb = br.read_nbits(4) + 1u;
      if (nb <= 5u) return kDecErrInvalidBpf;
      if (nb == 16u) {
        for (uint32_t i = 0; i < bl; i++) {
          lw = (int16_t)br.read_nbits(16);
          out[p_wav + i] = (int16_t)lw;
        }
      } else {
        for (uint32_t i = 0; i < bl; i++) {
          int32_t v = (int32_t)(br.read_nbits(nb) & 0xffffu);
          if (v > (1 << (nb - 1u))) v -= (1 << nb);  // unsigned_to_i16, decoder.rs:198-207 (strictly greater)
          lw = (int16_t)(lw + v);
          out[p_wav + i] = (int16_t)lw;
        }
      }
    } else {
      const uint32_t code = P.codes[ftype - 1u];
      const uint32_t inv_len = rice_inv_len(code);
      if (ftype == 1u) {  // decode_ricecode_block_r1 :147-170
        for (uint32_t i = 0; i < bl; i++) {
          const uint32_t z = br.count_zero_bits();
          br.read_nbits(1);
          if (z >= inv_len) return kDecErrOutOfBoundsInverse;
          lw = (int16_t)(lw + unfold(z));
          out[p_wav + i] = (int16_t)lw;
        }
      } else {  // decode_ricecode_block_r2r3 :172-196
        const uint32_t nb = ftype == 2u ? 2u : 4u;
        const int32_t level = 1 << code;  // 1 << nsubs
        for (uint32_t i = 0; i < bl; i++) {
          const int32_t nz = (int16_t)br.count_zero_bits();
          const int32_t r = (int16_t)br.read_nbits(nb);
          const int32_t iv = (int16_t)(r + (int16_t)(level * (int16_t)(nz - 1)));
          if (iv < 0 || (uint32_t)iv >= inv_len) return kDecErrOutOfBoundsInverse;
          lw = (int16_t)(lw + unfold((uint32_t)iv));
          out[p_wav + i] = (int16_t)lw;
        }
      }
    }
    remaining -= bl;
    p_wav += bl;
  }
  return kDecOk;
}

X3_HD uint32_t crc16_bytes(const uint16_t *T, const uint8_t *d, uint32_t n) {
  uint32_t s = 0xffffu;
  for (uint32_t i = 0; i < n; i++) s = crc16_byte(T, s, d[i]);
  return s;
}

// ------------------------------------------------------------------------------------------------
// Payload CRC by folding (decodefile.rs:93-103, crc.rs:44-52), one thread per payload.
//
// With P = x^16 + x^12 + x^5 + 1 and a = x^32 mod P:  squaring is an automorphism of GF(2)[x]/P (P is square free), so
// a = x^(2^5) has the same minimal polynomial as x, i.e.  a^16 = a^12 + a^5 + 1.  A message of 32-bit words
// w_0 .. w_(n-1) is the polynomial sum w_j a^(n-1-j); keeping it as sixteen words S_15 .. S_0 (sum S_i a^i) and taking
// in one more word is
//     S <- S * a + w :   fb = S_15;  S_i <- S_(i-1);  S_12 ^= fb;  S_5 ^= fb;  S_0 = fb ^ w
// -- the CRC's own shift register, but on whole words: THREE xors per 32 bits of payload, no tables, no shifts (the
// bytes of a word never mix, so the words are taken as they lie in memory).  The sixteen words left at the end are a
// 64-byte message with the same remainder, finished by the ordinary word-at-a-time CRC.  The shift S_i <- S_(i-1) is
// a renaming: the loop is unrolled by sixteen words (64 bytes, four 16-byte loads) and every register index is static.
// Zero words in front of a message do not change it, so the fold starts at the 16-byte boundary at or before the
// payload with everything before the payload's first byte masked to zero; the initial value 0xffff of CRC-16/CCITT-
// FALSE is xored into the first two payload bytes.  What follows the last whole 64-byte block (up to 15 words and
// a halfword) goes through the ordinary CRC.
//
// `payload` 2-byte aligned, `len` even; reads whole aligned 16-byte vectors that contain payload bytes, and nothing
// outside [payload & ~15, (payload + len + 15) & ~15) -- the caller guarantees those lie inside its buffer.
// ------------------------------------------------------------------------------------------------
X3_HD uint32_t crc16_fold_load(const uint8_t *p) {   // little-endian 32-bit word at a 4-byte aligned address
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const uint32_t *>(p);
#else
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}
struct CrcVec { uint32_t w[4]; };
X3_HD CrcVec crc16_fold_load16(const uint8_t *p) {   // 16-byte aligned
  CrcVec v;
#if defined(__CUDA_ARCH__)
  const uint4 q = *reinterpret_cast<const uint4 *>(p);
  v.w[0] = q.x; v.w[1] = q.y; v.w[2] = q.z; v.w[3] = q.w;
#else
  for (int k = 0; k < 4; k++) v.w[k] = crc16_fold_load(p + 4 * k);
#endif
  return v;
}
// Where the 64-byte blocks come from.  Concept: void begin(const uint8_t *first_block, uint32_t n_blocks);
// void block(uint32_t b, CrcVec v[4]) -- block b's four vectors, called for b = 0, 1, .. n_blocks-1 in order.
struct CrcMemorySource {   // straight from memory (host simulation; the device kernel streams through a ring)
  const uint8_t *base;
  X3_HD void begin(const uint8_t *first_block, uint32_t) { base = first_block; }
  X3_HD void block(uint32_t b, CrcVec v[4]) const {
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = crc16_fold_load16(base + 64u * b + 16 * q);
  }
};
template <class Source>
X3_HD uint32_t crc16_fold(Source &src, const uint8_t *payload, uint32_t len) {
  const uintptr_t pa = (uintptr_t)payload;
  const uint8_t *A16 = payload - (pa & 15u);                 // start of the fold
  const uint32_t k0 = (uint32_t)(pa & 15u) >> 2;             // word (of the first block) that holds the first payload byte
  const uint32_t lead = (uint32_t)(pa & 3u);                 // 0 or 2 bytes of that word are not payload
  const uint32_t span = (uint32_t)(pa & 15u) + len;          // bytes from A16 to the payload's end
  const uint32_t nwords = span >> 2;                         // whole words from A16 that end inside the payload
  const uint32_t nb = nwords >> 4;                           // whole 64-byte blocks
  // the first payload word as the fold wants it: bytes before the payload cleared, 0xffff xored into payload bytes 0, 1
  const uint32_t first_and = lead ? 0xffff0000u : 0xffffffffu, first_xor = lead ? 0xffff0000u : 0x0000ffffu;
  const bool has_first_word = nwords > k0;                    // else the payload is a single halfword (len 2, lead 0)
  uint32_t acc = has_first_word ? 0u : 0xffffu;
  uint32_t k = k0;                                            // next word (index from A16) for the ordinary CRC
  if (nb > 0u) {
    uint32_t R[16];
    src.begin(A16, nb);
    {
      // block 0 into the empty state: S_i = w_(15-i)
      CrcVec v[4];
      src.block(0u, v);
#pragma unroll
      for (int t = 0; t < 16; t++) {
        uint32_t w = v[t >> 2].w[t & 3];
        if ((uint32_t)t < k0) w = 0u;
        if ((uint32_t)t == k0) w = (w & first_and) ^ first_xor;
        R[15 - t] = w;
      }
    }
    for (uint32_t b = 1; b < nb; b++) {
      CrcVec v[4];
      src.block(b, v);
#pragma unroll
      for (int t = 0; t < 16; t++) {
        const int pi = 15 - t;                                // holds S_15 now, S_0 after this step
        const uint32_t fb = R[pi];
        R[(pi + 12) & 15] ^= fb;                              // becomes S_12
        R[(pi + 5) & 15] ^= fb;                               // becomes S_5
        R[pi] = fb ^ v[t >> 2].w[t & 3];
      }
    }
#pragma unroll
    for (int i = 15; i >= 0; i--) acc = crc16_word_alu(acc, bswap32(R[i]));
    k = 16u * nb;
  }
  for (; k < nwords; k++) {
    uint32_t w = crc16_fold_load(A16 + 4u * k);
    if (k == k0) w = (w & first_and) ^ first_xor;
    acc = crc16_word_alu(acc, bswap32(w));
  }
  if (span & 2u) {                                            // the payload ends with half a word
    const uint8_t *h = A16 + 4u * nwords;
#if defined(__CUDA_ARCH__)
    const uint32_t v = *reinterpret_cast<const uint16_t *>(h);
#else
    const uint32_t v = (uint32_t)h[0] | ((uint32_t)h[1] << 8);
#endif
    acc = crc16_half_alu(acc, ((v & 0xffu) << 8) | (v >> 8));
  }
  return acc & 0xffffu;
}

// ------------------------------------------------------------------------------------------------
// Fast path
// ------------------------------------------------------------------------------------------------

// Reader concept (fast path).  The decoder only ever moves forward, at most 32 bits at a time:
//   void block_begin()                   called once per block (the device reader tops up its ring here)
//   void window(hi, lo)                  the 64 bits that start at the current bit position (big-endian numeric)
//   void advance(n)                      consume n <= 32 bits
//   uint32_t bits_used()                 bits consumed since the start of the payload
// window() may show bytes that lie after the payload; decode_frame_fast checks at the end that it never consumed
// a bit past the payload (the reference zero-fills there) and retries exactly if so.

// Plain reader over memory (host simulation): assembles the window bytewise, zero past the buffer end.
struct PlainBitReader {
  const uint8_t *payload;
  const uint8_t *end;
  uint32_t pos;
  X3_HD void init(const uint8_t *p, const uint8_t *stream_end) { payload = p; end = stream_end; pos = 0; }
  X3_HD void block_begin() {}
  X3_HD void advance(uint32_t n) { pos += n; }
  X3_HD uint32_t bits_used() const { return pos; }
  X3_HD void window(uint32_t &hi, uint32_t &lo) const {
    const uint8_t *p = payload + (pos >> 3);
    uint32_t w[3];
    for (int k = 0; k < 3; k++) {
      uint32_t v = 0;
      for (int b = 0; b < 4; b++) v = (v << 8) | ((p + 4 * k + b < end) ? (uint32_t)p[4 * k + b] : 0u);
      w[k] = v;
    }
    const uint32_t s = pos & 7u;
    hi = funnel_l(w[1], w[0], s);
    lo = funnel_l(w[2], w[1], s);
  }
};

constexpr uint32_t kFastGroup = 80;        // 4 blocks of 20 samples = 160 B = 5 whole 32-byte sectors of output
constexpr uint32_t kStageWords = 16;       // staged words per thread; word j of thread t lives at stage[j*stride] with
                                           // stride = threads per CTA on the device (conflict-free 32-bit accesses), 1 on the host.
                                           // A block adds 10 words; whole sectors (8 words) leave after every block -- one
                                           // after each of the first three blocks of a group, two after the fourth -- and the
                                           // 2, 4 or 6 words left over move to the front.  (64 bytes instead of 160 per thread:
                                           // shared memory is what limits the number of resident frames.)

// Fast-path eligibility of a frame (Parameters::default() is checked by the caller).
X3_HD bool frame_fast_eligible(uint32_t samples, uint32_t payload_len, uintptr_t payload_addr, uintptr_t out_addr) {
  return samples >= kFastGroup && samples % kFastGroup == 0u && payload_len >= 2u && (payload_len & 1u) == 0u &&
         (payload_addr & 1u) == 0u && (out_addr & 31u) == 0u;
}

// Inverse-fold tables of the fast path.  A Rice code is z zeros, then nbk = k+1 bits r whose first bit is the
// terminator; the reference's index is i = r + level*(z-1) with level = 2^k (decoder.rs:157-165, :184-191) and the
// delta is INV_RICE_CODE[i] (x3.rs:200-204).
//
// The kernel never computes z.  The 32 bits t that start at the code are converted to a float, rounding toward zero
// (I2FP.F32.U32.RZ, an ALU-pipe instruction of ~5 cycles; counting leading zeros is an XU instruction of ~23): the
// float's exponent field is e = 158 - z and its mantissa starts with the bits that follow the terminator, so
//   Q = bits(float(t)) >> (24 - nbk) = e * level + (r - level)          (one shift)
// names the code, and the bit position moves on by z + nbk = nbk + 158 - (bits >> 23), which is one IMAD.HI
// (bits * 512 >> 32, plus the rest as the addend) on the otherwise idle FMA pipe.  The chain from one code to the next
// (the compiler turns the constant multiply into LEA.HI; a real IMAD.HI, and Q as a multiply-high with the table offset
// as addend, were measured slower: 1.21 -> 1.38 and 1.23 ms).  The chain from one code to the next is shift -> convert
// -> shift-add -> add: 24 cycles instead of the 39 of shift -> count -> add (tools/ubench/latency.cu).
// The delta is looked up by Q: entry [inv_tab_off(f) + Q] of the bank, f = ftype 1..3 (nbk 1, 2, 4).  i = (r - level)
// + level * z.  Every Q that is not a code of the table (index >= inv_len = 16 / 26 / 60: decoder.rs:161,187; an
// all-zero peek, for which the float is 0 and Q = 0) holds kInvBad, which no delta equals (|INV_RICE_CODE[i]| <= 30
// for i < 60): the block keeps the minimum of its deltas and a frame that saw kInvBad goes to the exact path.
//
// Entries are ONE BYTE, and the offsets of the three tables are chosen so that the entries of valid codes lie in
// disjoint banks: f = 3 (Q 1208..1271) in banks 0..15, f = 2 (Q 292..317) in banks 16..23, f = 1 (Q 143..158) in
// banks 24..28.  Lanes decode unrelated frames, so their lookups hit different tables and different codes at the same
// time; this layout keeps them free of bank conflicts.
typedef int8_t inv_entry_t;
constexpr int kInvBad = -128;
constexpr int kInvExpBias = 158;                    // exponent field of float(t) for a t with bit 31 set
#ifndef X3_DEC_FIXSHIFT
#define X3_DEC_FIXSHIFT 0   // 1 was measured: one instruction less per code, bank conflicts between the tables, no change in time
#endif
X3_HD uint32_t inv_nbk(uint32_t f) { return f == 1u ? 1u : (f == 2u ? 2u : 4u); }
X3_HD uint32_t inv_len_of(uint32_t f) { return f == 1u ? 16u : (f == 2u ? 26u : 60u); }   // x3.rs:214,222,250
#if X3_DEC_FIXSHIFT
// X3_DEC_FIXSHIFT: every table is indexed by Q = bits >> 20 (exponent and THREE mantissa bits) whatever nbk is, and the
// mantissa bits a code does not own are don't-cares (its entries are repeated 4 or 2 times).  The shift is then a
// constant, and shift + table offset become ONE LEA.HI instead of a shift and an add.  The tables are 1272 bytes each
// and can no longer sit in disjoint banks; their offsets put the entries of the frequent short zero runs (z <= 3 of
// every table) in different banks.
constexpr int kInvQShift = 20;
constexpr int kInvTabLen = (kInvExpBias + 1) * 8;   // 1272
constexpr int kInvOff1 = 0, kInvOff2 = 1312, kInvOff3 = 2624;   // = 0, 32, 64 mod 128
constexpr int kInvTabEntries = kInvOff3 + kInvTabLen;
#else
constexpr int kInvOff1 = 84, kInvOff2 = 288, kInvOff3 = 712;
constexpr int kInvTabEntries = kInvOff3 + (kInvExpBias + 1) * 8;   // 1984
#endif
X3_HD uint32_t inv_tab_off(uint32_t f) { return f == 1u ? (uint32_t)kInvOff1 : (f == 2u ? (uint32_t)kInvOff2 : (uint32_t)kInvOff3); }
X3_HD uint32_t inv_q_shift(uint32_t f) {
#if X3_DEC_FIXSHIFT
  (void)f;
  return (uint32_t)kInvQShift;
#else
  return 24u - inv_nbk(f);
#endif
}
X3_HD inv_entry_t inv_tab_entry(int j /* 0..kInvTabEntries */) {
  for (uint32_t f = 1; f <= 3; f++) {
    const int per_e = 1 << (23 - (int)inv_q_shift(f));          // entries per exponent value
    const int level = 1 << (inv_nbk(f) - 1u);                    // of which `level` are distinct codes
    const int Q = j - (int)inv_tab_off(f);
    if (Q < 0 || Q >= (kInvExpBias + 1) * per_e) continue;
    const int z = kInvExpBias - Q / per_e, i = (Q % per_e) / (per_e / level) + level * z;
    return (z > 31 || i >= (int)inv_len_of(f)) ? (inv_entry_t)kInvBad : (inv_entry_t)unfold((uint32_t)i);
  }
  return (inv_entry_t)kInvBad;   // between the tables
}

// bits of (float)t rounded toward zero
X3_HD uint32_t f32_rz_bits(uint32_t t) {
#if defined(__CUDA_ARCH__)
  float f;
  asm("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(t));
  return __float_as_uint(f);
#else
  if (t == 0u) return 0u;
  const uint32_t z = (uint32_t)__builtin_clz(t);
  return (((uint32_t)kInvExpBias - z) << 23) | (((t << z) >> 8) & 0x7fffffu);
#endif
}

// per-ftype constants of a Rice block, fetched with one 16-byte load
struct alignas(16) RiceBlockPar {
  uint32_t rc;        // -(158 + nbk): what a code adds to the bits left in the window, apart from its exponent field
  uint32_t sh;        // inv_q_shift: 24 - nbk, or 20 for every table (X3_DEC_FIXSHIFT)
  uint32_t one;       // 1 (opaque to the compiler; entry 0 only)
  uint32_t tab_off;   // entry of Q = 0 in the table bank
};
X3_HD RiceBlockPar rice_block_par(uint32_t f) {
  RiceBlockPar p;
  const uint32_t nbk = inv_nbk(f ? f : 1u);
  p.rc = 0u - ((uint32_t)kInvExpBias + nbk);
  p.sh = inv_q_shift(f ? f : 1u);
  p.one = 1u;
  p.tab_off = inv_tab_off(f ? f : 1u);
  return p;
}

// hi32(a * b) + c and a * b + c as single FMA-pipe instructions (IMAD.HI.U32 / IMAD): the decoder is bound by the
// ALU pipe (shifts, logic, min/max), so the shift by a per-block amount is done as a multiply by a power of two.
X3_HD uint32_t mad_hi_u32(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return (uint32_t)(((uint64_t)a * b) >> 32) + c;
#endif
}
X3_HD uint32_t mad_lo_u32(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return a * b + c;
#endif
}
// one 32-byte sector (p is 32-byte aligned): a single 256-bit store on sm_100 (STG.E.256).  Each lane writes into its
// own frame, so every store instruction touches 32 different lines; halving the number of store instructions halves
// that part of the L1 data-pipe load, which is what bounds the decode kernel.
X3_HD void store_sector(uint4 *p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                        uint32_t w6, uint32_t w7) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2), "r"(w3),
               "r"(w4), "r"(w5), "r"(w6), "r"(w7)
               : "memory");
#else
  uint4 a, b;
  a.x = w0; a.y = w1; a.z = w2; a.w = w3;
  b.x = w4; b.y = w5; b.z = w6; b.w = w7;
  p[0] = a;
  p[1] = b;
#endif
}
X3_HD uint32_t max3u(uint32_t a, uint32_t b, uint32_t c) {  // VIMNMX3
  const uint32_t m = a > b ? a : b;
  return m > c ? m : c;
}
X3_HD int32_t min3s(int32_t a, int32_t b, int32_t c) {  // VIMNMX3
  const int32_t m = a < b ? a : b;
  return m < c ? m : c;
}
X3_HD uint32_t pack_lo16(uint32_t lo, uint32_t hi) {  // (lo & 0xffff) | (hi << 16)
#if defined(__CUDA_ARCH__)
  return __byte_perm(lo, hi, 0x5410);
#else
  return (lo & 0xffffu) | (hi << 16);
#endif
}

// one Rice code of the 64-bit window hi:lo, `left` bits before the window's end (left = 32 - offset): see the table
// comment.  left only ever decreases; once it is negative the (clamped) shift shows hi again and the block is flagged.
#if X3_DEC_FIXSHIFT
#define X3_RICE_LOOKUP(fb) inv_tab[((fb) >> kInvQShift) + bp.tab_off]
#else
#define X3_RICE_LOOKUP(fb) tab[(fb) >> bp.sh]
#endif
#define X3_RICE_SAMPLE(dv)                                                            \
  {                                                                                   \
    const uint32_t fb = f32_rz_bits(funnel_r(lo, hi, left));                          \
    left = mad_hi_u32(fb, 512u, left + bp.rc);                                        \
    dv = (int32_t)X3_RICE_LOOKUP(fb);                                                 \
    lw += dv;                                                                         \
  }

// Decode one frame.  `stage` = this thread's staging area: 40 words, `ss` words apart.
// Returns kDecOk or kDecRetryExact.
// `inv_tab` = kInvTabEntries entries built with inv_tab_entry, `par` = rice_block_par(0..3) (shared memory on the device).
template <class Reader>
X3_HD int decode_frame_fast(Reader &rd, uint32_t payload_len, int16_t *out, uint32_t samples, uint32_t *stage,
                            const uint32_t ss, const inv_entry_t *inv_tab, const RiceBlockPar *par) {
  uint32_t hi, lo;
  rd.block_begin();
  rd.window(hi, lo);
  int32_t lw = (int32_t)(hi >> 16);        // first sample, decoder.rs:42 (only the low 16 bits of lw matter)
  rd.advance(16);
  uint32_t prev = (uint32_t)lw;            // last sample not yet written (low half of the next output word)
  bool bad = false;

  const uint32_t nblk = samples / 20u;     // the last block has 19 samples (encoder.rs:194)
  uint4 *out4 = reinterpret_cast<uint4 *>(out);

  for (uint32_t b = 0; b < nblk; b++) {
    rd.block_begin();
    const bool tail = (b == nblk - 1u);
    const uint32_t t4 = b & 3u;               // position in the group of four blocks; 2*t4 words are waiting in the stage
    uint32_t *st = stage + 2u * t4 * ss;

    rd.window(hi, lo);
    const uint32_t ftype = hi >> 30;
    if (ftype != 0u) {
      // ---- Rice block: z zeros, then nbk bits of which the first is the terminator ----
      rd.advance(2);
      const RiceBlockPar bp = par[ftype];
      const inv_entry_t *tab = inv_tab + bp.tab_off;
      (void)tab;
      int32_t dmin = 0, mprev = 0;
      uint32_t lneg = 0, lprev = 0;   // sign bit: some group ran past its 32 bits
      // samples x0..x19; output words (prev,x0) (x1,x2) ... (x17,x18); x19 becomes prev.
      // X3_DEC_GROUP codes per reader step.  The codes of a group but the last must end inside the first 32 bits of
      // the window (the peek of the next one starts there); the last may run into the second word, and the reader
      // then moves in two steps.  Codes the encoder writes are at most 10 bits (20 > thresholds, x3.rs:33-40), a
      // valid code at most 16: three of the former always fit, longer ones send the frame to the exact path.
      // (Four per step execute fewer instructions but were measured slower, 1.21 -> 1.26 ms on C2 and 0.29 -> 0.37 ms
      // on C1: the next window waits for a longer chain.)
#ifndef X3_DEC_GROUP
#define X3_DEC_GROUP 3
#endif
      constexpr int G = X3_DEC_GROUP, NG = (20 + G - 1) / G;
      static_assert(G == 3 || G == 4, "codes per reader step");
#pragma unroll
      for (int g = 0; g < NG; g++) {
        rd.window(hi, lo);
        uint32_t left = 32u;
        int32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
        for (int j = 0; j < G; j++) {
          const int i = G * g + j;
          if (i < 20 && (i < 19 || !tail)) {
            if (G == 4 && j == 3) lneg |= left;
            if (j == 0) X3_RICE_SAMPLE(d0)
            else if (j == 1) X3_RICE_SAMPLE(d1)
            else if (j == 2) X3_RICE_SAMPLE(d2)
            else X3_RICE_SAMPLE(d3)
            if ((i & 1) == 0) st[(i >> 1) * ss] = pack_lo16(prev, (uint32_t)lw);
            else prev = (uint32_t)lw;
          }
        }
        // validity is tracked on the ALU pipe, the kernel's busiest, where a min / max costs two slots: one three-way
        // minimum per group, folded into the block's minimum every second group; `left` only needs its sign, which an
        // OR keeps (every second group, three-way as well)
        const int32_t m = G == 4 ? min3s(d0 < d1 ? d0 : d1, d2, d3) : min3s(d0, d1, d2);
        if (G == 3) {
          if (g & 1) {
            dmin = min3s(dmin, mprev, m);
            lneg |= lprev | left;
          } else {
            mprev = m;
            lprev = left;
          }
          rd.advance(32u - left);  // more than 32 bits is malformed for this path (`bad` below); the reader stays
                                   // inside its ring whatever it is given
        } else {
          dmin = dmin < m ? dmin : m;
          uint32_t cum = 32u - left;
          if (cum > 32u) { rd.advance(32u); cum -= 32u; }
          rd.advance(cum);
        }
      }
      if (G == 3 && (NG & 1)) {
        dmin = dmin < mprev ? dmin : mprev;
        lneg |= lprev;
      }
      // out-of-range index (decoder.rs:161,187; includes every zero run the 32-bit peek cannot see the end of)
      // or a group of codes longer than 32 bits (its later codes were parsed from the wrong place) -> the exact
      // path decides
      if (dmin == kInvBad || (int32_t)lneg < 0) bad = true;
    } else {
      const uint32_t nb = ((hi >> 26) & 15u) + 1u;  // decoder.rs:211
      rd.advance(6);
      if (nb <= 5u) { bad = true; break; }           // FrameDecodeInvalidBPF, decoder.rs:213-216
      if (nb == 16u) {
        // ---- literal block: raw 16-bit samples ----
        for (int j = 0; j < 10; j++) {
          rd.window(hi, lo);
          lw = (int32_t)(hi >> 16);
          st[j * ss] = pack_lo16(prev, (uint32_t)lw);
          if (j < 9 || !tail) {
            lw = (int32_t)(hi & 0xffffu);
            prev = (uint32_t)lw;
            rd.advance(32);
          } else {
            rd.advance(16);
          }
        }
      } else {
        // ---- BFP block: nb-bit two's complement differences (decoder.rs:224-231) ----
        const int32_t half = 1 << (nb - 1u), full = 1 << nb;
        for (int j = 0; j < 10; j++) {
          rd.window(hi, lo);
          int32_t v = (int32_t)(hi >> (32u - nb));
          if (v > half) v -= full;                    // unsigned_to_i16: strictly greater, decoder.rs:203
          lw += v;
          st[j * ss] = pack_lo16(prev, (uint32_t)lw);
          if (j < 9 || !tail) {
            v = (int32_t)(funnel_l(lo, hi, nb) >> (32u - nb));
            if (v > half) v -= full;
            lw += v;
            prev = (uint32_t)lw;
            rd.advance(2u * nb);
          } else {
            rd.advance(nb);
          }
        }
      }
    }
    if (bad) break;
    // ---- whole 32-byte sectors leave after every block (never a partial sector: half-written sectors were
    // measured to cost ~10 % extra DRAM traffic) ----
    {
      uint4 *o = out4 + ((size_t)(b >> 2) * 5u + t4) * 2u;
      store_sector(o, stage[0], stage[ss], stage[2u * ss], stage[3u * ss], stage[4u * ss], stage[5u * ss], stage[6u * ss],
                   stage[7u * ss]);
      if (t4 == 3u) {
        store_sector(o + 2, stage[8u * ss], stage[9u * ss], stage[10u * ss], stage[11u * ss], stage[12u * ss],
                     stage[13u * ss], stage[14u * ss], stage[15u * ss]);
      } else {
        const uint32_t left = 2u * t4 + 2u;
#pragma unroll
        for (uint32_t k = 0; k < 6u; k++)
          if (k < left) stage[k * ss] = stage[(8u + k) * ss];
      }
    }
  }
  if (bad) return kDecRetryExact;
  // bits consumed must lie inside the payload (the reference zero-fills past the end; we may have read
  // the next frame's bytes there instead)
  if (rd.bits_used() > 8u * payload_len) return kDecRetryExact;
  return kDecOk;
}

// ------------------------------------------------------------------------------------------------
// Generic fast path: any Parameters (block_len, codes) and any number of samples, same Reader concept, one sample at a
// time: no unrolling by block shape, no sector staging (samples leave in 4-byte stores where the address allows, else
// 2-byte stores).  Slower than decode_frame_fast, an order of magnitude faster than the exact path (which reads the
// payload bytewise and is only meant to settle malformed frames).  Same contract: whatever it cannot prove well
// formed makes it give up (kDecRetryExact) and the exact path decides what the reference would have reported.
// The Rice index is computed, not looked up: i = z (ftype 1) or r + level * (z - 1) (decoder.rs:157-165, :184-191),
// delta = INV_RICE_CODE[i] in closed form (x3.rs:200-204).
// ------------------------------------------------------------------------------------------------
template <class Reader>
X3_HD int decode_frame_generic(Reader &rd, uint32_t payload_len, int16_t *out, uint32_t samples, const CodecParams &P) {
  uint32_t hi, lo;
  rd.block_begin();
  rd.window(hi, lo);
  int32_t lw = (int32_t)(int16_t)(hi >> 16);   // first sample, decoder.rs:42
  rd.advance(16);
  // pairing for 4-byte stores: `held` is the sample at the even (4-byte aligned) position before the current one
  uint32_t p = 0;
  uint32_t held = 0;
  bool have_held = false;
#define X3_GEN_EMIT(v)                                                                         \
  {                                                                                            \
    const uint32_t s16 = (uint32_t)(v) & 0xffffu;                                              \
    if ((((uintptr_t)(out + p)) & 2u) == 0u) {                                                 \
      held = s16;                                                                              \
      have_held = true;                                                                        \
    } else if (have_held) {                                                                    \
      *reinterpret_cast<uint32_t *>(out + p - 1) = held | (s16 << 16);                         \
      have_held = false;                                                                       \
    } else {                                                                                   \
      out[p] = (int16_t)s16;                                                                   \
    }                                                                                          \
    p++;                                                                                       \
  }
  X3_GEN_EMIT(lw)
  uint32_t remaining = samples - 1u;
  // The reader is topped up every 16 .. 18 samples (<= 37 bytes of payload), at a block start or inside a long block,
  // and never twice in a row: a top-up waits for the copies of the one before it.
  uint32_t since = 16u;
  while (remaining > 0u) {
    const uint32_t bl = remaining < P.block_len ? remaining : P.block_len;  // decoder.rs:50
    if (since >= 16u) { rd.block_begin(); since = 0u; }
    rd.window(hi, lo);
    const uint32_t ftype = hi >> 30;
    if (ftype == 0u) {
      const uint32_t nb = ((hi >> 26) & 15u) + 1u;  // decoder.rs:211
      rd.advance(6);
      if (nb <= 5u) return kDecRetryExact;          // FrameDecodeInvalidBPF: the exact path reports it
      const int32_t half = 1 << (nb - 1u), full = 1 << nb;
      // two fields per reader step (2 * 16 bits at most), every lane in step: the trip counts depend on the block
      // length only, so the 32 frames of a warp do not diverge
      for (uint32_t i = 0; i < bl; i += 2u) {
        if (since >= 16u) { rd.block_begin(); since = 0u; }
        since += 2u;
        rd.window(hi, lo);
        const bool two = i + 1u < bl;
        int32_t v = (int32_t)(hi >> (32u - nb));
        if (nb == 16u) {
          lw = (int32_t)(int16_t)v;                  // literal block: raw samples
        } else {
          if (v > half) v -= full;                   // unsigned_to_i16: strictly greater, decoder.rs:203
          lw = (int32_t)(int16_t)(lw + v);
        }
        X3_GEN_EMIT(lw)
        if (two) {
          v = (int32_t)(funnel_l(lo, hi, nb) >> (32u - nb));
          if (nb == 16u) {
            lw = (int32_t)(int16_t)v;
          } else {
            if (v > half) v -= full;
            lw = (int32_t)(int16_t)(lw + v);
          }
          X3_GEN_EMIT(lw)
        }
        rd.advance(two ? 2u * nb : nb);
      }
    } else {
      rd.advance(2);
      const uint32_t code = P.codes[ftype - 1u];
      const int32_t inv_len = (int32_t)rice_inv_len(code);
      const uint32_t nbk = ftype == 1u ? 1u : (ftype == 2u ? 2u : 4u);   // decoder.rs:158,180
      const int32_t level = 1 << code;
      // two codes per reader step, every lane in step (see above); a pair longer than 32 bits -- possible only for
      // codes beyond the default thresholds, or malformed -- sends the frame to the exact path
      bool bad = false;
      for (uint32_t i = 0; i < bl; i += 2u) {
        if (since >= 16u) { rd.block_begin(); since = 0u; }
        since += 2u;
        rd.window(hi, lo);
        const bool two = i + 1u < bl;
        uint32_t z = clz32(hi);
        uint32_t cum = z + nbk;
        bad = bad || cum > 32u;
        uint32_t r = shl_safe(hi, z) >> (32u - nbk);
        int32_t iv = ftype == 1u ? (int32_t)z : (int32_t)r + level * ((int32_t)z - 1);
        bad = bad || iv < 0 || iv >= inv_len;        // OutOfBoundsInverse: the exact path reports it
        lw = (int32_t)(int16_t)(lw + unfold((uint32_t)iv));
        X3_GEN_EMIT(lw)
        if (two) {
          const uint32_t t = funnel_l(lo, hi, cum > 32u ? 32u : cum);
          z = clz32(t);
          cum += z + nbk;
          bad = bad || cum > 32u;
          r = shl_safe(t, z) >> (32u - nbk);
          iv = ftype == 1u ? (int32_t)z : (int32_t)r + level * ((int32_t)z - 1);
          bad = bad || iv < 0 || iv >= inv_len;
          lw = (int32_t)(int16_t)(lw + unfold((uint32_t)iv));
          X3_GEN_EMIT(lw)
        }
        if (bad) return kDecRetryExact;
        rd.advance(cum);
      }
    }
    remaining -= bl;
  }
  if (have_held) out[p - 1] = (int16_t)held;
#undef X3_GEN_EMIT
  if (rd.bits_used() > 8u * payload_len) return kDecRetryExact;
  return kDecOk;
}

}  // namespace x3
