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

using namespace x3;

namespace {
constexpr int NT = 512;
uint16_t g_T[kCrcBankEntries2];
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
  const bool fast = params_are_default(P) && params[1] <= 512 && !force_generic;
  size_t pos = 0;
  for (size_t s0 = 0; s0 < n; s0 += P.spf) {
    const uint32_t fn = (uint32_t)((n - s0) < P.spf ? (n - s0) : P.spf);
    std::vector<uint8_t> tmp(64 + 2 * (size_t)fn * 9);
    const size_t L = fast ? sim_encode_frame_fast(pcm + s0, fn, P, s0 + fn >= n, tmp.data(), stats)
                          : sim_encode_frame(pcm + s0, fn, P, false, s0 + fn >= n, tmp.data(), stats);
    if (pos + L > cap) return -15;
    memcpy(out + pos, tmp.data(), L);
    pos += L;
  }
  *out_len = pos;
  return 0;
}

// The decoder's table bank against the reference's definition (decoder.rs:157-191, x3.rs:200-252): for every zero
// run z the 32-bit peek can see and every suffix r whose terminator bit is set, q = z*2^nbk + r must map to
// INV_RICE_CODE[r + level*(z-1)] when that index is inside the code's table (inv_len 16 / 26 / 60) and must be
// >= inv_q_end otherwise; q must be monotone in the index.  Returns 0 when everything holds, else a failure code.
int sim_inv_table_check() {
  const int inv_len[4] = {0, 16, 26, 60};
  for (int f = 1; f <= 3; f++) {
    const RiceBlockPar bp = rice_block_par((uint32_t)f);
    const int nbk = (int)bp.nbk, level = 1 << (nbk - 1);
    if (bp.sh != 32u - bp.nbk || bp.q_end != inv_q_end((uint32_t)f)) return 10 + f;
    if (bp.tab_off != (uint32_t)((f - 1) * kInvTabLen + kInvPad)) return 20 + f;
    int last_i = -1;
    for (int z = 0; z <= 31; z++)
      for (int r = level; r < (1 << nbk); r++) {
        const int q = (z << nbk) + r, i = r + level * (z - 1);
        if (q + kInvPad >= kInvTabLen) return 30 + f;          // must stay inside the table
        if (i <= last_i) return 40 + f;                        // q order == index order
        last_i = i;
        const bool valid = i >= 0 && i < inv_len[f];
        if (valid != ((uint32_t)q < bp.q_end)) return 50 + f;
        if (valid && (int)inv_tab_entry(f, q + kInvPad) != unfold((uint32_t)i)) return 60 + f;
      }
    for (int q = -kInvPad; q < 0; q++)                          // the pad an all-zero peek lands in
      if (inv_tab_entry(f, q + kInvPad) != 0) return 70 + f;
  }
  if (rice_block_par(0).q_end != 1u) return 80;                 // the opaque constant 1 of X3_CUM_ADD
  return 0;
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
    for (int j = 0; j < kInvTabEntries; j++) inv[j] = inv_tab_entry(1 + j / kInvTabLen, j % kInvTabLen);
    RiceBlockPar par[4];
    for (uint32_t f = 0; f < 4; f++) par[f] = rice_block_par(f);
    r = decode_frame_fast(rd, payload_len, out, samples, stage, 1u, inv, par);
    if (r == kDecOk) *used_fast = 1;
  }
  if (r == kDecRetryExact) r = decode_frame_exact(pl, payload_len, out, samples, P);
  return r;
}

}  // extern "C"
