// sim_x3.cpp -- CPU simulation of the CUDA kernels' algorithms, phase by phase, using the SAME
// host+device source (x3_enc_core.cuh / x3_dec_core.cuh / x3_common.cuh / x3_crc_host.h).
//
// TEST HARNESS ONLY (built by tests/test_sim_kernels.py into tests/sim/libx3sim.so).  It lets the bit
// packing, owner-merge, chunked-CRC combine, frame header, fast/exact decode logic be checked against
// the oracle on the CPU box before GPU time is spent.  It is not part of the product library and the
// product never falls back to it.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../x3-rust_b200/csrc/x3_crc_host.h"
#include "../../x3-rust_b200/csrc/x3_dec_core.cuh"
#include "../../x3-rust_b200/csrc/x3_enc_core.cuh"
#include "../../x3-rust_b200/csrc/x3_enc_strip.cuh"

using namespace x3;

namespace {
constexpr int NT = 512;
uint16_t g_T[kCrcBankEntries3];
bool g_T_ready = false;
const uint16_t *T() {
  if (!g_T_ready) { build_crc_bank(g_T); g_T_ready = true; }
  return g_T;
}

// chunked CRC exactly as encode phase A/B and crc_frames_kernel do it (32 simulated lanes)
uint32_t crc_chunked(const uint32_t *words_img, uint32_t payload_len) {
  const uint16_t *t = T();
  const uint32_t m = payload_len >> 4;
  std::vector<uint16_t> chunk(m + 1);
  for (uint32_t c = 0; c < m; c++) {
    uint32_t s = c == 0 ? 0xffffu : 0u;
    for (int w = 0; w < 4; w++) s = crc16_word(t, s, bswap32(words_img[4 * c + w]));
    chunk[c] = (uint16_t)s;
  }
  uint32_t h[32];
  for (uint32_t lane = 0; lane < 32; lane++) {
    h[lane] = 0;
    if (m > lane)
      for (int i = (int)((m - 1u - lane) >> 5); i >= 0; i--) {
        const uint32_t c = m - 1u - (32u * (uint32_t)i + lane);
        h[lane] = crc16_mulc(t, 4, h[lane]) ^ chunk[c];
      }
  }
  for (int k = 0; k < 5; k++) {
    uint32_t o[32];
    for (int lane = 0; lane < 32; lane++) o[lane] = lane + (1 << k) < 32 ? h[lane + (1 << k)] : h[lane];  // shfl_down
    for (int lane = 0; lane < 32; lane++) h[lane] ^= crc16_mulc(t, 6 + 2 * k, o[lane]);
  }
  uint32_t s = m ? h[0] : 0xffffu;
  const uint32_t rem = payload_len & 15u;
  uint32_t wi = m * 4u;
  for (uint32_t done = 0; done + 4u <= rem; done += 4u) s = crc16_word(t, s, bswap32(words_img[wi++]));
  if (rem & 2u) s = crc16_half(t, s, bswap32(words_img[wi]) >> 16);
  return s & 0xffffu;
}

// CRC as the fast encode kernel does it: per-warp slices of 32 chunks (shuffle tree), then Horner over slices
uint32_t crc_sliced(const uint32_t *words_img, uint32_t payload_len) {
  const uint16_t *t = T(), *t2 = T() + kCrcTableEntries;
  const uint32_t m = payload_len >> 4, nslices = (m + 31u) >> 5;
  std::vector<uint32_t> V(nslices + 1);
  for (uint32_t j = 0; j < nslices; j++) {
    uint32_t h[32];
    for (uint32_t lane = 0; lane < 32; lane++) {
      const uint32_t e = 32u * j + lane;
      h[lane] = 0;
      if (e < m) {
        const uint32_t c = m - 1u - e;
        uint32_t s = c == 0 ? 0xffffu : 0u;  // byte-swapped state form, swapped table bank, words as they lie
        for (int w = 0; w < 4; w++) s = crc16_word_sw(t2, s, words_img[4 * c + w]);
        h[lane] = s;
      }
    }
    for (int k = 0; k < 5; k++) {
      uint32_t o[32];
      for (int lane = 0; lane < 32; lane++) o[lane] = lane + (1 << k) < 32 ? h[lane + (1 << k)] : h[lane];
      for (int lane = 0; lane < 32; lane++) {
        const uint32_t x = o[lane] | 0xabcd0000u;  // bits above 15 must be ignored
        h[lane] ^= k == 0 ? crc16_mulc_sw<6>(t2, x) : k == 1 ? crc16_mulc_sw<8>(t2, x) : k == 2 ? crc16_mulc_sw<10>(t2, x)
                 : k == 3 ? crc16_mulc_sw<12>(t2, x) : crc16_mulc_sw<14>(t2, x);
      }
    }
    V[j] = bswap16(h[0]);
  }
  uint32_t s = 0;
  for (int j = (int)nslices - 1; j >= 0; j--) s = crc16_mulc(t, 4, s) ^ V[j];
  if (m == 0) s = 0xffffu;
  const uint32_t rem = payload_len & 15u;
  uint32_t wi = m * 4u;
  for (uint32_t done = 0; done + 4u <= rem; done += 4u) s = crc16_word(t, s, bswap32(words_img[wi++]));
  if (rem & 2u) s = crc16_half(t, s, bswap32(words_img[wi]) >> 16);
  return s & 0xffffu;
}

// one frame, mirroring encode_frames_fast_kernel (16 worker warps; FastSink, one atomicOr per unaligned block)
size_t sim_encode_frame_fast(const int16_t *pcm, uint32_t n, const CodecParams &P, bool last_frame, uint8_t *out,
                             uint64_t stats[6]) {
  const uint32_t BL = 20;
  const uint32_t nblk = n > 1 ? (n - 2u) / BL + 1u : 1u;
  std::vector<int16_t> s_in(n + 64, (int16_t)0x5a5a);
  memcpy(s_in.data(), pcm, n * sizeof(int16_t));
  std::vector<uint32_t> words((16u + nblk * 330u) / 32u + 16u, 0xdeadbeefu);
  struct Th { bool active, use_fast; uint32_t len, nbits, start, bit_off; FastBlock fb; BlockMode mode; };
  std::vector<Th> th(512);
  uint32_t run = 0;
  for (int tid = 0; tid < 512; tid++) {
    Th &t = th[tid];
    const uint32_t b = tid;
    t.active = b < nblk;
    t.start = 1u + b * BL;
    t.len = 0;
    if (t.active && n > t.start) t.len = (n - t.start) < BL ? (n - t.start) : BL;
    t.mode.kind = kRice; t.mode.k = 0; t.mode.hdr = 0; t.mode.stat = 0;
    t.nbits = 0; t.use_fast = false;
    if (t.active) {
      if (t.len >= BL - 1) { t.use_fast = true; t.mode = block_measure_fast(s_in.data(), t.start, t.len, t.fb, t.nbits, -1); }
      else if (t.len > 0) t.mode = block_measure_generic(s_in.data(), t.start, t.len, P, t.nbits);
      if (b == 0) t.nbits += 16;
      if (t.len > 0) stats[t.mode.stat] += t.len;
    }
    t.bit_off = run;
    run += t.nbits;
  }
  const uint32_t total_bits = run, payload_len = payload_bytes(total_bits);
  std::vector<int16_t> s_in_pack = s_in;
  if (!last_frame) std::fill(s_in_pack.begin(), s_in_pack.end(), (int16_t)0x7b7b);  // next frame's prefetch
  std::vector<uint32_t> tail(512, 0u);
  std::vector<uint32_t *> tail_at(512, nullptr);
  for (int tid = 0; tid < 512; tid++) {
    Th &t = th[tid];
    if (!t.active) continue;
    FastSink sink;
    sink.init(t.bit_off, words.data());
    if (tid == 0) { sink.put((uint32_t)(uint16_t)(t.use_fast ? t.fb.pred : (int32_t)s_in_pack[0]), 16); sink.flush(); }
    if (t.use_fast) block_pack_fast(t.fb, t.len, t.mode, sink);
    else if (t.len > 0) block_pack_generic(s_in_pack.data(), t.start, t.len, t.mode, sink);
    tail[tid] = sink.finish_tail();
    if (sink.cnt) tail_at[tid] = sink.dst;
  }
  if (total_bits & 31u) words[total_bits >> 5] = 0u;   // tid 0, before the barrier
  for (int tid = 0; tid < 512; tid++)
    if (tail_at[tid]) *tail_at[tid] |= tail[tid];        // shared-memory OR after the barrier
  const uint32_t crc = crc_sliced(words.data(), payload_len);
  const uint32_t hc = header_crc(T(), 1u, n, payload_len);
  uint32_t hdr[5] = {bswap32((kFrameKey << 16) | 0x0101u), bswap32(((n & 0xffffu) << 16) | (payload_len & 0xffffu)), 0u, 0u,
                     bswap32((hc << 16) | crc)};
  memcpy(out, hdr, 20);
  memcpy(out + 20, words.data(), payload_len);
  return 20 + payload_len;
}


// one frame, mirroring encode_frames_strip_kernel: 128 simulated threads, one strip of four blocks each; local pack in
// place in 176-byte rows, scan, relocation into a window aligned like the stream modulo 16 bytes (gpay = stream offset
// of the payload), windowed rounds, sliced CRC over 32-byte chunks.  win_bytes = window size (multiple of 1024).
size_t sim_encode_frame_strip(const int16_t *pcm, uint32_t n, const CodecParams &P, unsigned long long gpay,
                              uint32_t win_bytes, uint8_t *out, uint64_t stats[6]) {
  const uint16_t *t2 = T() + kCrcTableEntries;
  const uint32_t nblk = n > 1 ? (n - 2u) / 20u + 1u : 1u;
  std::vector<uint32_t> rows((kStripMaxRows + 1) * kRowWords, 0x5a5a5a5au);
  uint32_t s_next[4][4];
  memset(s_next, 0x6b, sizeof s_next);
  // stage_rows: chunk c of warp w -> byte 16 * (c + c / 10) of the warp's rows (the kernel's multiply-shift division)
  for (uint32_t w = 0; w < 4; w++) {
    const uint32_t w0 = w * 32u * kStripSamples;
    const uint32_t avail = n > w0 ? n - w0 : 0u, chunks = avail >> 3;
    unsigned char *dst = reinterpret_cast<unsigned char *>(rows.data() + w * 32u * kRowWords + kRowPadWords);
    for (uint32_t c = 0; c < 320u && c < chunks; c++) {
      if (((c * 205u) >> 11) != c / 10u) return 0;
      memcpy(dst + 16u * (c + ((c * 205u) >> 11)), pcm + w0 + 8u * c, 16);
    }
    if (chunks > 320u) memcpy(s_next[w], pcm + w0 + 2560u, 16);
    if (avail < 32u * kStripSamples + 8u)
      for (uint32_t i = chunks << 3; i < avail && i < 32u * kStripSamples + 8u; i++) {
        if (i < 32u * kStripSamples)
          reinterpret_cast<int16_t *>(rows.data() + (w * 32u + i / kStripSamples) * kRowWords + kRowPadWords)[i % kStripSamples] = pcm[w0 + i];
        else
          reinterpret_cast<int16_t *>(s_next[w])[i - 32u * kStripSamples] = pcm[w0 + i];
      }
  }
  uint32_t nxt[128], Tb[128], O[128];
  for (uint32_t t = 0; t < 128; t++) nxt[t] = (t & 31u) == 31u ? s_next[t >> 5][0] : rows[(t + 1) * kRowWords + kRowPadWords];
  uint32_t short_stats[6] = {0, 0, 0, 0, 0, 0};
  unsigned long long stat_acc_sum[6] = {0, 0, 0, 0, 0, 0};
  for (uint32_t t = 0; t < 128; t++) {
    uint32_t *row = rows.data() + t * kRowWords;
    Tb[t] = 0;
    const uint32_t b0 = kStripBlocks * t;
    if (b0 >= nblk) continue;
    if (b0 + kStripBlocks <= nblk && n >= kStripSamples * (t + 1u)) {
      const bool full = n > kStripSamples * (t + 1u);
      unsigned long long acc = 0;
      uint32_t s19 = 0;
      Tb[t] = strip_pack_fast(row, nxt[t], full, t == 0, -1, acc, s19);
      if (!full) short_stats[s19] += 19u;
      for (int m = 0; m < 6; m++) stat_acc_sum[m] += ((acc >> (10 * m)) & 1023u) * 20u;
    } else {
      Tb[t] = strip_pack_generic(row, nxt[t], t, n, nblk, P, short_stats);
    }
  }
  for (int m = 0; m < 6; m++) stats[m] += stat_acc_sum[m] + short_stats[m];
  uint32_t total_bits = 0;
  for (uint32_t t = 0; t < 128; t++) { O[t] = total_bits; total_bits += Tb[t]; }
  const uint32_t payload_len = payload_bytes(total_bits);
  const uint32_t a_off = (uint32_t)(gpay & 15u), end_bytes = a_off + payload_len;
  const uint32_t win_words = win_bytes / 4u, win_chunks = win_bytes / 32u;
  const uint32_t nrounds = (end_bytes + win_bytes - 1u) / win_bytes, nch = end_bytes >> 5;
  std::vector<uint32_t> win(win_words + 8u);
  std::vector<uint32_t> V(64, 0u);
  // regular frames (the kernel's common path): every strip but the last has >= 32 bits, the payload fits the window,
  // the window is 16-byte aligned to the payload: relocation without merging (strip_relocate_fast)
  const uint32_t nstrips = (nblk + kStripBlocks - 1u) / kStripBlocks;
  bool regular = payload_len <= win_bytes && a_off == 0u;
  for (uint32_t t = 0; t + 1u < nstrips; t++) regular = regular && Tb[t] >= 32u;
  std::vector<uint8_t> image(end_bytes + 64u, 0xEE);   // window-space bytes as written to the stream
  for (uint32_t r = 0; r < nrounds; r++) {
    std::fill(win.begin(), win.end(), 0xdeadbeefu);     // stale data of the previous round / frame
    const int32_t wbit0 = (int32_t)(8u * r * win_bytes);
    if (r == 0) for (uint32_t k = 0; k < (a_off >> 2); k++) win[k] = 0u;
    const int32_t zt = (int32_t)(8u * a_off + total_bits) - wbit0;
    if (zt >= 0 && (zt >> 5) < (int32_t)win_words) { win[zt >> 5] = 0u; win[(zt >> 5) + 1] = 0u; }
    if (regular) {
      std::fill(win.begin(), win.end(), 0xdeadbeefu);
      for (uint32_t t = 0; t < nstrips; t++)
        strip_relocate_fast(rows.data() + t * kRowWords, Tb[t], O[t], t ? Tb[t - 1] : 32u, t + 1u == nstrips, win.data());
    } else {
      uint32_t tail[128];
      int32_t tail_idx[128];
      for (uint32_t t = 0; t < 128; t++)
        strip_relocate(rows.data() + t * kRowWords, Tb[t], (int32_t)(8u * a_off + O[t]) - wbit0, win.data(), win_words, tail[t], tail_idx[t]);
      for (uint32_t t = 0; t < 128; t++)
        if (tail_idx[t] >= 0) win[tail_idx[t]] |= tail[t];
    }
    const uint32_t vb0 = r == 0 ? a_off : 0u;
    const uint32_t vb1 = end_bytes - r * win_bytes < win_bytes ? end_bytes - r * win_bytes : win_bytes;
    memcpy(image.data() + r * win_bytes + vb0, reinterpret_cast<const uint8_t *>(win.data()) + vb0, vb1 - vb0);
    const uint32_t c_lo = r * win_chunks, c_hi = nch < (r + 1u) * win_chunks ? nch : (r + 1u) * win_chunks;
    if (c_hi > c_lo) {
      const uint32_t j_lo = (nch - c_hi) >> 5, j_hi = (nch - 1u - c_lo) >> 5;
      for (uint32_t j = j_lo; j <= j_hi; j++) {
        uint32_t h[32];
        for (uint32_t lane = 0; lane < 32; lane++) {
          const uint32_t e = 32u * j + lane;
          h[lane] = 0;
          if (e >= nch) continue;
          const uint32_t c = nch - 1u - e;
          if (c < c_lo || c >= c_hi) continue;
          uint32_t q[8];
          memcpy(q, win.data() + 8u * (c - c_lo), 32);
          if (c == 0) q[a_off >> 2] ^= 0xffffu << (8u * (a_off & 2u));
          // two 16-byte halves, then the lane's own power of x (as crc_slices does)
          uint32_t h0 = 0, h1 = 0;
          for (int w = 0; w < 4; w++) { h0 = crc16_word_sw(t2, h0, q[w]); h1 = crc16_word_sw(t2, h1, q[4 + w]); }
          const uint16_t *N = T() + kCrcBankEntries2;
          uint32_t hh = crc16_mul_nib(N + 64 * kCrcMulX128, h0) ^ h1;
          hh = crc16_mul_nib(N + 64 * ((lane >> 2) ? 3 + (lane >> 2) : 0), crc16_mul_nib(N + 64 * (lane & 3), hh));
          h[lane] = hh;
        }
        for (int lane = 1; lane < 32; lane++) h[0] ^= h[lane];
        V[j] ^= h[0] & 0xffffu;
      }
    }
    if (r == nrounds - 1u) {
      const uint32_t nsl = (nch + 31u) >> 5;
      uint32_t s = 0;
      for (int j = (int)nsl - 1; j >= 0; j--) s = crc16_mul_nib(T() + kCrcBankEntries2 + 64 * kCrcMulX8192, s) ^ V[j];
      const uint32_t base = r * win_bytes;
      uint32_t pos = 32u * nch;
      while (pos + 4u <= end_bytes) {
        uint32_t w = win[(pos - base) >> 2];
        if (nch == 0 && (pos >> 2) == (a_off >> 2)) w ^= 0xffffu << (8u * (a_off & 2u));
        s = crc16_word_sw(t2, s, w);
        pos += 4u;
      }
      if (pos < end_bytes) {
        uint32_t hv = win[(pos - base) >> 2] & 0xffffu;
        if (nch == 0 && (pos >> 2) == (a_off >> 2) && (a_off & 2u) == 0u) hv ^= 0xffffu;
        s = crc16_half_sw(t2, s, hv);
      }
      const uint32_t hc = header_crc_sw(t2, 1u, n, payload_len), crc = bswap16(s);
      uint32_t hdr[5] = {bswap32((kFrameKey << 16) | 0x0101u), bswap32(((n & 0xffffu) << 16) | (payload_len & 0xffffu)), 0u, 0u,
                         bswap32((hc << 16) | crc)};
      memcpy(out, hdr, 20);
    }
  }
  memcpy(out + 20, image.data() + a_off, payload_len);
  return 20 + payload_len;
}

// one frame, mirroring encode_frames_generic_kernel (single CTA, NT simulated threads)
size_t sim_encode_frame(const int16_t *pcm, uint32_t n, const CodecParams &P, bool fast_kernel, bool last_frame,
                        uint8_t *out, uint64_t stats[6]) {
  const uint32_t BL = P.block_len;
  const uint32_t nblk = n > 1 ? (n - 2u) / BL + 1u : 1u;
  const uint32_t rounds = (nblk + NT - 1) / NT;
  std::vector<int16_t> s_in(n + 64, (int16_t)0x5a5a);  // junk after the frame, as in shared memory
  memcpy(s_in.data(), pcm, n * sizeof(int16_t));
  const uint32_t cap_words = (16u + nblk * (2u + BL * 64u)) / 32u + 16u;
  std::vector<uint32_t> words(cap_words, 0xdeadbeefu);  // NOT zeroed: every word must be written exactly once
  std::vector<uint32_t> offs(nblk + 2), Hs(nblk + 2, 0xdeadbeefu), Ts(nblk + 2, 0xdeadbeefu);
  uint32_t bit_base = 0;
  for (uint32_t r = 0; r < rounds; r++) {
    struct Th { bool active, use_fast; uint32_t len, nbits, start; FastBlock fb; BlockMode mode; };
    std::vector<Th> th(NT);
    for (int tid = 0; tid < NT; tid++) {  // measure
      Th &t = th[tid];
      const uint32_t b = r * NT + tid;
      t.active = b < nblk;
      t.start = 1u + b * BL;
      t.len = 0;
      if (t.active && n > t.start) t.len = (n - t.start) < BL ? (n - t.start) : BL;
      t.mode.kind = kRice; t.mode.k = 0; t.mode.hdr = 0; t.mode.stat = 0;
      t.nbits = 0;
      t.use_fast = false;
      if (t.active) {
        if (t.len > 0) {
          t.mode = block_measure_generic(s_in.data(), t.start, t.len, P, t.nbits);
        }
        if (b == 0) t.nbits += 16;
        if (t.len > 0) stats[t.mode.stat] += t.len;
      }
    }
    uint32_t run = bit_base;  // scan
    std::vector<uint32_t> bit_off(NT);
    for (int tid = 0; tid < NT; tid++) { bit_off[tid] = run; run += th[tid].nbits; }
    bit_base = run;
    // in the fast kernel the input buffer is overwritten by the next frame's prefetch at this point
    // (there is a next frame unless this is the stream's last one)
    std::vector<int16_t> s_in_pack = s_in;
    if (fast_kernel && !last_frame && r == rounds - 1) std::fill(s_in_pack.begin(), s_in_pack.end(), (int16_t)0x7b7b);
    for (int tid = 0; tid < NT; tid++) {  // pack
      Th &t = th[tid];
      const uint32_t b = r * NT + tid;
      if (!t.active) continue;
      offs[b] = bit_off[tid];
      BitSink sink;
      sink.init(bit_off[tid], words.data(), &Hs[b]);
      if (b == 0) {
        sink.put((uint32_t)(uint16_t)(t.use_fast ? t.fb.pred : (int32_t)s_in_pack[0]), 16);
        sink.flush();
      }
      if (t.len > 0) block_pack_generic(s_in_pack.data(), t.start, t.len, t.mode, sink);
      bool has_tail;
      Ts[b] = sink.finish(has_tail);
    }
  }
  const uint32_t total_bits = bit_base;
  const uint32_t payload_len = payload_bytes(total_bits);
  offs[nblk] = total_bits; offs[nblk + 1] = 0xffffffffu; Hs[nblk] = 0; Hs[nblk + 1] = 0;
  for (uint32_t b = 0; b < nblk; b++) {  // merge
    const uint32_t end = offs[b + 1];
    if (end & 31u) {
      const uint32_t lw = end >> 5, o = offs[b];
      if ((o >> 5) < lw || (o & 31u) == 0u) {
        uint32_t v = Ts[b] | Hs[b + 1];
        for (uint32_t j = b + 2; (offs[j] >> 5) == lw; j++) v |= Hs[j];
        words[lw] = v;
      }
    }
  }
  const uint32_t crc = crc_chunked(words.data(), payload_len);
  const uint32_t hc = header_crc(T(), 1u, n, payload_len);
  uint32_t hdr[5] = {bswap32((kFrameKey << 16) | 0x0101u), bswap32(((n & 0xffffu) << 16) | (payload_len & 0xffffu)), 0u, 0u,
                     bswap32((hc << 16) | crc)};
  memcpy(out, hdr, 20);
  memcpy(out + 20, words.data(), payload_len);
  return 20 + payload_len;
}
}  // namespace

extern "C" {

int sim_encode(const int16_t *pcm, size_t n, const uint32_t *params /*bl,bpf,c0,c1,c2,t0,t1,t2*/, uint8_t *out,
               size_t cap, size_t *out_len, uint64_t *stats, int force_generic) {
  CodecParams P;
  P.block_len = params[0];
  P.spf = params[0] * params[1];
  for (int k = 0; k < 3; k++) { P.codes[k] = params[2 + k]; P.thresholds[k] = params[5 + k]; }
  // force_generic: low 4 bits 0 = strip kernel, 1 = generic kernel, 2 = round-1 fast kernel; bits 4..7 = (stream base
  // address mod 16) / 2 and bits 8.. = window bytes / 1024 (0 = 7) for the strip kernel
  const int mode = force_generic & 15;
  const bool fast = params_are_default(P) && params[1] <= 512 && mode != 1 && (P.spf % 8u) == 0u;
  const unsigned long long base_mod = 2ull * ((force_generic >> 4) & 7);
  const uint32_t win_bytes = 1024u * ((force_generic >> 8) ? (uint32_t)(force_generic >> 8) : 7u);
  size_t pos = 0;
  for (size_t s0 = 0; s0 < n; s0 += P.spf) {
    const uint32_t fn = (uint32_t)((n - s0) < P.spf ? (n - s0) : P.spf);
    std::vector<uint8_t> tmp(64 + 2 * (size_t)fn * 9);
    const size_t L = !fast ? sim_encode_frame(pcm + s0, fn, P, false, s0 + fn >= n, tmp.data(), stats)
                     : mode == 2 ? sim_encode_frame_fast(pcm + s0, fn, P, s0 + fn >= n, tmp.data(), stats)
                                 : sim_encode_frame_strip(pcm + s0, fn, P, base_mod + pos + 20u, win_bytes, tmp.data(), stats);
    if (L == 0) return -200;
    if (pos + L > cap) return -15;
    memcpy(out + pos, tmp.data(), L);
    pos += L;
  }
  *out_len = pos;
  return 0;
}

// The decoder's table bank against the reference's definition (decoder.rs:157-191, x3.rs:200-252): for every zero
// run z the 32-bit peek can see, every suffix r whose terminator bit is set and every filler below it, the window t
// must give Q = bits(float(t), toward zero) >> (24 - nbk) whose entry is INV_RICE_CODE[r + level*(z-1)] when that
// index is inside the code's table (inv_len 16 / 26 / 60) and kInvBad otherwise, and the multiply-high step must move
// the position by exactly z + nbk.  The valid entries of the three tables must lie in disjoint shared-memory banks.
// Returns 0 when everything holds, else a failure code.
int sim_inv_table_check() {
  const int inv_len[4] = {0, 16, 26, 60};
  uint32_t banks_used = 0;
  for (int f = 1; f <= 3; f++) {
    const RiceBlockPar bp = rice_block_par((uint32_t)f);
    const int nbk = (int)inv_nbk((uint32_t)f), level = 1 << (nbk - 1);
    if (bp.sh != inv_q_shift((uint32_t)f) || bp.rc != 0u - (158u + (uint32_t)nbk) || bp.tab_off != inv_tab_off((uint32_t)f)) return 10 + f;
    uint32_t banks = 0;
    for (int z = 0; z <= 31; z++)
      for (int r = level; r < (1 << nbk); r++)
        for (int fill = 0; fill < 3; fill++) {
          if (z + nbk > 32) continue;
          const int below = 32 - z - nbk;   // bits of the window after the code
          const uint32_t rest = below == 0 ? 0u : (fill == 0 ? 0u : fill == 1 ? (below == 32 ? ~0u : (1u << below) - 1u) : (0x9e3779b9u >> (32 - below)));
          const uint32_t t = (uint32_t)(((uint64_t)r << below) | rest);
          const uint32_t fb = f32_rz_bits(t);
          const uint32_t Q = fb >> bp.sh;
          if (bp.tab_off + Q >= (uint32_t)kInvTabEntries) return 30 + f;     // must stay inside the bank
          const int i = r + level * (z - 1);
          const bool valid = i < inv_len[f];
          const int d = (int)inv_tab_entry((int)(bp.tab_off + Q));
          if (valid ? d != unfold((uint32_t)i) : d != kInvBad) return 40 + f;
          uint32_t left = 32u;
          left = mad_hi_u32(fb, 512u, left + bp.rc);
          if (left != 32u - (uint32_t)(z + nbk)) return 50 + f;
          if (valid && z <= 3) banks |= 1u << (((bp.tab_off + Q) >> 2) & 31u);   // the frequent codes
        }
    if (inv_tab_entry((int)bp.tab_off) != kInvBad) return 60 + f;          // an all-zero peek: float 0, Q = 0
    if (banks & banks_used) return 70 + f;
    banks_used |= banks;
  }
  for (int j = 0; j < kInvTabEntries; j++) {                              // nothing but deltas and kInvBad
    const int d = inv_tab_entry(j);
    if (d != kInvBad && (d < -30 || d > 30)) return 80;
  }
  if (rice_block_par(0).one != 1u) return 90;
  return 0;
}

// crc16_fold over `len` payload bytes placed at offset `off` (even) of a 16-byte aligned buffer
uint32_t sim_crc_fold(const uint8_t *data, uint32_t len, uint32_t off) {
  std::vector<uint8_t> raw(len + off + 64 + 16, 0xa5);
  uint8_t *base = raw.data() + ((16 - ((uintptr_t)raw.data() & 15)) & 15);
  memcpy(base + off, data, len);
  CrcMemorySource src;
  return crc16_fold(src, base + off, len);
}

uint32_t sim_crc(const uint8_t *data, uint32_t len) {  // len even
  std::vector<uint32_t> w((len + 19) / 4 + 4, 0);
  memcpy(w.data(), data, len);
  return crc_chunked(w.data(), len);
}

// decode one frame payload located at stream+pos (header at pos) the way decode_frames_kernel does;
// returns the kDec* status.  `mode`: 0 = kernel policy (fast if eligible, exact on retry), 1 = exact only.
int sim_decode_frame(const uint8_t *stream, size_t stream_len, size_t pos, uint32_t samples, uint32_t payload_len,
                     const uint32_t *params, int16_t *out, int mode, int *used_fast) {
  CodecParams P;
  P.block_len = params[0];
  P.spf = params[0] * params[1];
  for (int k = 0; k < 3; k++) { P.codes[k] = params[2 + k]; P.thresholds[k] = params[5 + k]; }
  const uint8_t *pl = stream + pos + 20;
  const bool dflt = P.block_len == 20 && P.codes[0] == 0 && P.codes[1] == 1 && P.codes[2] == 3;
  int r = kDecRetryExact;
  *used_fast = 0;
  if (samples == 0 || payload_len < 2) return kDecErrPanic;
  if (mode == 0 && dflt && frame_fast_eligible(samples, payload_len, (uintptr_t)pl, (uintptr_t)out)) {
    alignas(16) uint32_t stage[kStageWords];
    PlainBitReader rd;
    rd.init(pl, stream + stream_len);
    static inv_entry_t inv[kInvTabEntries];
    for (int j = 0; j < kInvTabEntries; j++) inv[j] = inv_tab_entry(j);
    RiceBlockPar par[4];
    for (uint32_t f = 0; f < 4; f++) par[f] = rice_block_par(f);
    r = decode_frame_fast(rd, payload_len, out, samples, stage, 1u, inv, par);
    if (r == kDecOk) *used_fast = 1;
  } else if (mode == 0) {   // every other frame: the generic fast path
    PlainBitReader rd;
    rd.init(pl, stream + stream_len);
    r = decode_frame_generic(rd, payload_len, out, samples, P);
    if (r == kDecOk) *used_fast = 2;
  }
  if (r == kDecRetryExact) r = decode_frame_exact(pl, payload_len, out, samples, P);
  return r;
}

}  // extern "C"
