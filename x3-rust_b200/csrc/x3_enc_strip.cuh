// x3_enc_strip.cuh -- per-thread logic of the strip encoder (encode_frames_strip_kernel, x3_encode.cu).
//
// One thread owns a STRIP of four consecutive 20-sample blocks (80 samples) of a frame, so a 10 000-sample frame is
// 125 threads = 4 warps, and everything that is per thread and frame (scan, barriers, copy-out bookkeeping) is paid
// once per four blocks.  The frame is encoded in ONE pass over the samples:
//   1. local pack  -- the thread codes its blocks (encoder.rs:289-315, same per-block functions as the other
//                     kernels) into a bit string that starts at bit 0 of ITS OWN ROW of the staging buffer, in place:
//                     the row (176 bytes = 16 bytes of padding + the strip's 160 bytes of PCM) ends up holding the
//                     strip's code bits (at most 16 + 4*326 bits = 168 bytes).  The bit string starts in the padding
//                     and grows more slowly than the samples are consumed (a block of 40 bytes codes to at most 41),
//                     so it never reaches a block that has not been loaded yet.  No other thread touches the row, so
//                     there is nothing to merge and nothing to wait for; the bit count T falls out of the packing.
//   2. scan        -- CTA exclusive scan of T -> the strip's bit offset O in the frame payload; frame size published.
//   3. relocate    -- the thread shifts its local words by (O + 8*a) mod 32 into the frame image ("window"), where a
//                     is the payload's global byte address mod 16: the image is then byte-for-byte what the stream
//                     holds at a 16-byte aligned global address, and goes out with one bulk async store (TMA).
// Words of the local bit string are kept as big-endian VALUES (first bit = bit 31); the byte swap to stream order
// happens once per output word in step 3.
#pragma once

#include "x3_enc_core.cuh"

namespace x3 {

constexpr uint32_t kStripBlocks = 4;
constexpr uint32_t kStripSamples = kStripBlocks * (uint32_t)kFastBL;  // 80
constexpr uint32_t kRowWords = 44;        // 176-byte rows: 16-byte vector loads of 32 consecutive rows hit 32 distinct bank groups
constexpr uint32_t kRowPadWords = 4;      // the strip's samples start 16 bytes into the row
constexpr uint32_t kStripMaxRows = 128;   // <= 512 blocks per frame

// MSB-first bit sink into the thread's own row (big-endian valued words, no byte swap).  Same contract as FastSink:
// put() needs cnt + n <= 64, flush() stores at most one word.
struct RowSink {
  uint64_t acc;
  uint32_t cnt;
  smaddr_t dst;
  X3_HD void init(uint32_t *row) {
    acc = 0;
    cnt = 0;
    dst = sm_addr(row);
  }
  X3_HD void put(uint32_t v, uint32_t n) {
    acc = (acc << n) | (uint64_t)v;
    cnt += n;
  }
  X3_HD void flush() {
    const uint32_t w = (uint32_t)(acc >> (cnt & 31u));   // cnt in [32,63]: acc >> (cnt - 32)
#if defined(__CUDA_ARCH__)
    asm volatile("{ .reg .pred p; setp.ge.u32 p, %1, 32; @p st.shared.u32 [%0], %2; @p add.u32 %0, %0, 4; }"
                 : "+r"(dst)
                 : "r"(cnt), "r"(w)
                 : "memory");
#else
    const uint32_t full = cnt >> 5;
    sm_store_if(full != 0u, dst, w);
    dst = sm_advance(dst, full);
#endif
    cnt &= 31u;
  }
  // end of the strip: the last partial word goes out zero padded, and one zero word after the bit string (the
  // relocation reads one word past the end); returns the strip's bit count
  X3_HD uint32_t finish(uint32_t *row) {
    flush();
    const uint32_t w = cnt ? (uint32_t)(acc << (32u - cnt)) : 0u;
    sm_store_if(true, dst, w);
    sm_store_if(cnt != 0u, sm_advance(dst, 1), 0u);
#if defined(__CUDA_ARCH__)
    return 8u * (dst - sm_addr(row)) + cnt;
#else
    return 32u * (uint32_t)(dst - row) + cnt;
#endif
  }
};

// the eleven words (s[20b], s[20b+1]) ... (s[20b+20], s[20b+21]) of block j of the strip (8-byte aligned in the row);
// word 10 of the last block is the first word of the next strip's samples, read ahead of time by the caller (`nxt`)
X3_HD void strip_load_words(const uint32_t *row, uint32_t j, uint32_t nxt, uint32_t W[kFastBL / 2 + 1]) {
  const uint32_t *p = row + kRowPadWords + 10u * j;
#if defined(__CUDA_ARCH__)
  const uint2 *p2 = reinterpret_cast<const uint2 *>(p);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint2 v = p2[i];
    W[2 * i] = v.x;
    W[2 * i + 1] = v.y;
  }
#else
  for (int i = 0; i < 10; i++) W[i] = p[i];
#endif
  W[10] = j == kStripBlocks - 1u ? nxt : p[10];
}

// Folded differences and mode of one block of 20 (full) or 19 samples from its eleven words: block_measure_fast
// without the bit count (the strip encoder packs right away).  fold(d) = max(2d, ~2d), 2d by IDP.2A, ~x as x*(-1)-1
// against an opaque -1: see block_measure_fast.
X3_HD BlockMode block_fold_fast(const uint32_t W[kFastBL / 2 + 1], bool full, FastBlock &fb, int32_t neg1) {
  constexpr uint32_t kHiMinusLo = 0x02feu, kMinusHi = 0xfe00u, kPlusLo = 0x0002u;
  fb.pred = dp2a_s16s8(W[0], 0x0001u, 0);
  uint32_t maxu = 0;
#pragma unroll
  for (int j = 0; j < kFastBL / 2; j++) {
    const int32_t p0 = dp2a_s16s8(W[j], kHiMinusLo, 0);
    const int32_t p1 = dp2a_s16s8(W[j + 1], kPlusLo, dp2a_s16s8(W[j], kMinusHi, 0));
    const int32_t n0 = p0 * neg1 + neg1, n1 = p1 * neg1 + neg1;
    uint32_t u0 = (uint32_t)(p0 > n0 ? p0 : n0), u1 = (uint32_t)(p1 > n1 ? p1 : n1);
    if (j == kFastBL / 2 - 1 && !full) u1 = 0;
    fb.u[2 * j] = u0;
    fb.u[2 * j + 1] = u1;
    const uint32_t m01 = u0 > u1 ? u0 : u1;
    maxu = m01 > maxu ? m01 : maxu;
  }
  const uint32_t max_abs = (maxu + 1u) >> 1;
  BlockMode m;
  if (max_abs <= 20u) {            // thresholds 3 / 8 / 20, codes 0 / 1 / 3 (encoder.rs:304-314, x3.rs:93-96)
    const uint32_t ftype = (max_abs > 3u) + (max_abs > 8u);
    m.kind = kRice;
    m.k = ftype == 0 ? 0u : (ftype == 1 ? 1u : 3u);
    m.hdr = ftype + 1;
    m.stat = m.k;
  } else {
    const uint32_t nb = 32u - clz32(max_abs);
    if (nb >= 15) { m.kind = kLiteral; m.k = 15; m.hdr = 15; m.stat = 5; }
    else { m.kind = kBfp; m.k = nb; m.hdr = nb; m.stat = 4; }
  }
  return m;
}

// Local pack of a strip whose four blocks are all there: 20, 20, 20 and 20 (`full`) or 19 samples.
// `first`: the strip starts the frame, its bit string begins with the <Audio State> (encoder.rs:189).
// stat_acc: six 10-bit counters of full blocks per mode; len19_stat: stats index of the 19-sample block, if any.
// One copy of the block coder, four trips (the unrolled form is 58 KB of code: the kernel then waits for instructions;
// two blocks per trip with 16-byte loads was measured 2.5 % slower than this).
X3_HD uint32_t strip_pack_fast(uint32_t *row, uint32_t nxt, bool full, bool first, int32_t neg1,
                               unsigned long long &stat_acc, uint32_t &len19_stat) {
  RowSink sink;
  sink.init(row);
  if (first) {
    sink.put(row[kRowPadWords] & 0xffffu, 16);
    sink.flush();
  }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (uint32_t j = 0; j < kStripBlocks; j++) {
    uint32_t W[kFastBL / 2 + 1];
    FastBlock fb;
    strip_load_words(row, j, nxt, W);
    const bool fullj = full || j != kStripBlocks - 1u;
    const BlockMode m = block_fold_fast(W, fullj, fb, neg1);
    if (fullj) stat_acc += 1ull << (10u * m.stat);
    else len19_stat = m.stat;
    block_pack_fast(fb, fullj ? 20u : 19u, m, sink);
    sink.flush();
  }
  return sink.finish(row);
}

// Local pack of any other strip (the stream's short last frame, frames whose block count is not a multiple of four):
// the strip's samples are copied out of the row first, because the generic block coder re-reads them while packing.
// n = samples of the frame, strip = index of the strip; short_stats[6] += block lengths per mode.
X3_HD uint32_t strip_pack_generic(uint32_t *row, uint32_t nxt, uint32_t strip, uint32_t n, uint32_t nblk,
                                  const CodecParams &P, uint32_t short_stats[6]) {
  int16_t loc[kStripSamples + 2];
  for (uint32_t i = 0; i < kStripSamples / 2; i++) {
    const uint32_t w = row[kRowPadWords + i];
    loc[2 * i] = (int16_t)(w & 0xffffu);
    loc[2 * i + 1] = (int16_t)(w >> 16);
  }
  loc[kStripSamples] = (int16_t)(nxt & 0xffffu);
  loc[kStripSamples + 1] = 0;
  RowSink sink;
  sink.init(row);
  if (strip == 0) {
    sink.put((uint32_t)(uint16_t)loc[0], 16);
    sink.flush();
  }
  for (uint32_t j = 0; j < kStripBlocks; j++) {
    const uint32_t b = kStripBlocks * strip + j;
    if (b >= nblk) break;
    const uint32_t start = 1u + b * (uint32_t)kFastBL;
    if (n <= start) break;
    const uint32_t len = (n - start) < (uint32_t)kFastBL ? (n - start) : (uint32_t)kFastBL;
    uint32_t nbits;
    const BlockMode m = block_measure_generic(loc, 1u + j * (uint32_t)kFastBL, len, P, nbits);
    short_stats[m.stat] += len;
    block_pack_generic(loc, 1u + j * (uint32_t)kFastBL, len, m, sink);
  }
  return sink.finish(row);
}

// Step 3 for one strip and one window.  The strip's T bits start at bit `start` of the window (negative: the strip
// began in an earlier window).  Every window word that holds its own LAST bit inside the strip is stored plainly by
// this thread (bits of other strips in it are zero: they OR themselves in afterwards); a last, partial word is
// returned in `tail` / `tail_idx` for the caller to OR in after the plain stores of all threads are done.
X3_HD void strip_relocate(const uint32_t *row, uint32_t T, int32_t start, uint32_t *win, uint32_t win_words,
                          uint32_t &tail, int32_t &tail_idx) {
  tail = 0;
  tail_idx = -1;
  const int32_t end = start + (int32_t)T;
  if (T == 0u || end <= 0 || start >= (int32_t)(32u * win_words)) return;
  const int32_t nw = (int32_t)((T + 31u) >> 5);
  const int32_t d0 = start > 0 ? (start >> 5) : 0;
  int32_t dend = (end + 31) >> 5;
  if (dend > (int32_t)win_words) dend = (int32_t)win_words;
  const uint32_t s = (uint32_t)(-start) & 31u;
  int32_t j = (d0 * 32 - start) >> 5;                    // floor: -1 when the window word starts before the strip
  uint32_t prev = (j >= 0 && j < nw) ? row[j] : 0u;
  for (int32_t d = d0; d < dend; d++, j++) {
    const uint32_t cur = (j + 1 < nw) ? row[j + 1] : 0u;
    const uint32_t out = bswap32(funnel_l(cur, prev, s));
    prev = cur;
    if (end >= 32 * (d + 1)) {
      win[d] = out;
    } else {
      tail = out;
      tail_idx = d;
    }
  }
}

// Step 3, regular frames (every strip but the last holds at least 32 bits, the payload fits the window): no merging.
// A window word holds bits of at most two strips; it is stored by the strip that holds its last bit, which fetches the
// other strip's bits -- the last (start & 31) bits of its predecessor -- from the predecessor's row itself.  The last
// strip of the frame also stores its final partial word (the window is 16-bit aligned, so the padding to an even
// byte count never leaves that word).
// start = bit offset of the strip in the window, Tp = bit count of the predecessor (row - kRowWords).
X3_HD void strip_relocate_fast(const uint32_t *row, uint32_t T, uint32_t start, uint32_t Tp, bool last, uint32_t *win) {
  const uint32_t s = start & 31u;
  uint32_t *dst = win + (start >> 5);
  uint32_t nst = (s + T) >> 5;                       // words whose last bit is this strip's
  if (last && ((s + T) & 31u)) nst += 1u;            // + the final partial word (its zero bits are the even-byte padding)
  uint32_t pred = 0;
  if (s) {                                           // the predecessor's last s bits, left aligned
    const uint32_t *prow = row - kRowWords;
    const uint32_t e = Tp - s;                       // Tp >= 32 > s
    pred = funnel_l(prow[(e >> 5) + 1u], prow[e >> 5], e & 31u) & ~(0xffffffffu >> s);
  }
  uint32_t prevw = 0;
#if defined(__CUDA_ARCH__)
  const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
#pragma unroll 1
  for (uint32_t k = 0; k < nst; k += 4u) {
    const uint4 v = row4[k >> 2];
    uint32_t w0 = funnel_r(v.x, prevw, s), w1 = funnel_r(v.y, v.x, s), w2 = funnel_r(v.z, v.y, s), w3 = funnel_r(v.w, v.z, s);
    prevw = v.w;
    if (k == 0u) w0 |= pred;
    if (k + 4u <= nst) {
      dst[k] = bswap32(w0); dst[k + 1] = bswap32(w1); dst[k + 2] = bswap32(w2); dst[k + 3] = bswap32(w3);
    } else {
      dst[k] = bswap32(w0);
      if (k + 1u < nst) dst[k + 1] = bswap32(w1);
      if (k + 2u < nst) dst[k + 2] = bswap32(w2);
    }
  }
#else
  for (uint32_t k = 0; k < nst; k++) {
    const uint32_t cur = k < kRowWords ? row[k] : 0u;
    uint32_t w = funnel_r(cur, prevw, s);
    prevw = cur;
    if (k == 0u) w |= pred;
    dst[k] = bswap32(w);
  }
#endif
}

// swapped-state CRC of one big-endian halfword given as it lies in memory (see crc16_word_sw)
X3_HD uint32_t crc16_half_sw(const uint16_t *T2, uint32_t s_sw, uint32_t h_le) {
  const uint32_t v = (s_sw ^ h_le) & 0xffffu;
  return (uint32_t)T2[256 + (v & 0xffu)] ^ (uint32_t)T2[v >> 8];
}
// frame header bytes 0..16 -> header CRC, swapped state / swapped tables (encoder.rs:153)
X3_HD uint32_t header_crc_sw(const uint16_t *T2, uint32_t id, uint32_t num_samples, uint32_t payload_len) {
  uint32_t s = 0xffffu;
  s = crc16_word_sw(T2, s, 0x3378u | ((id & 0xffu) << 16) | ((id & 0xffu) << 24));
  s = crc16_word_sw(T2, s, ((num_samples >> 8) & 0xffu) | ((num_samples & 0xffu) << 8) | (((payload_len >> 8) & 0xffu) << 16) |
                               ((payload_len & 0xffu) << 24));
  s = crc16_word_sw(T2, s, 0u);
  s = crc16_word_sw(T2, s, 0u);
  return bswap16(s);
}

}  // namespace x3
