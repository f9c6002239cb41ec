// x3_encode.cu -- frame encoder kernels for sm_100a.
//
// Two kernels with the same phases (persistent CTAs, one frame per CTA at a time, frames handed out by an atomic
// ticket): encode_frames_generic_kernel for any Parameters the API accepts, and encode_frames_fast_kernel
// (further down, the one that is tuned) for Parameters::default() with at most 512 blocks per frame.
//   1. the frame's PCM is staged in shared memory (generic: 16-byte cp.async; fast: one TMA bulk copy on an mbarrier);
//   2. one thread per block: first difference, zig-zag fold, max -> mode (encoder.rs:304-314), bit length;
//   3. CTA-wide exclusive scan of the bit lengths -> bit offset of every block inside the frame;
//   4. each thread packs its block MSB-first straight into its final position in a shared-memory byte
//      image (x3_enc_core.cuh);
//   5. CRC-16 of the payload: 16-byte chunks in parallel, combined with multiply-by-x^n tables (shuffle tree +
//      Horner), header built (encoder.rs:122-162);
//   6. the frame's byte offset in the output stream: a prefix over the frame sizes (generic: decoupled look-back
//      per CTA; fast: a scanner CTA that turns published sizes into prefixes);
//   7. the image is copied to global memory with coalesced stores.
// HBM traffic is exactly the algorithmic figure: every PCM byte read once, every output byte written once
// (plus 8 bytes of status per frame).
#include <cuda_runtime.h>

#include "x3_enc_core.cuh"
#include "x3_enc_strip.cuh"
#include "x3_kernels.h"
#include "x3_lookback.cuh"

namespace x3 {

namespace {

constexpr int NT = kEncThreads;
constexpr int NW = NT / 32;

// stage frame f's samples into shared memory (asynchronously where 16-byte alignment allows)
template <int NTHREADS>
__device__ __forceinline__ void issue_frame_load(const EncodeArgs &a, uint32_t f, int16_t *s_in) {
  const unsigned long long s0 = (unsigned long long)f * a.P.spf;
  unsigned long long rem = a.n_samples - s0;
  const uint32_t n = rem < a.P.spf ? (uint32_t)rem : a.P.spf;
  const int16_t *src = a.pcm + s0;
  uint32_t done = 0;
  if ((((uintptr_t)src) & 15u) == 0) {
    const uint32_t chunks = n >> 3;  // 8 samples = 16 bytes
    for (uint32_t c = threadIdx.x; c < chunks; c += NTHREADS) cp_async16(s_in + c * 8, src + c * 8);
    done = chunks << 3;
  }
  for (uint32_t i = done + threadIdx.x; i < n; i += NTHREADS) s_in[i] = __ldg(src + i);
  cp_async_commit();
}

// Generic kernel: any Parameters the API accepts (and default Parameters with more than 512 blocks per frame).
// Same phases as the fast kernel below, but blocks re-read the staged samples and words are merged through the
// Hs/Ts exchange arrays (x3_enc_core.cuh, BitSink).
constexpr bool FAST = false;
__global__ void __launch_bounds__(NT, 1) encode_frames_generic_kernel(const EncodeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- shared memory carve-up (all offsets 16-byte aligned) ----
  const uint32_t in_bytes = (2u * (a.P.spf + 8u) + 15u) & ~15u;
  const uint32_t img_bytes = 32u + 4u * (a.out_words_cap + 8u);
  const uint32_t nblk_cap = a.max_blocks + 2u;
  unsigned char *p = smem_raw;
  int16_t *s_in = reinterpret_cast<int16_t *>(p);                 p += in_bytes;
  unsigned char *s_img = p;                                        p += (img_bytes + 15u) & ~15u;
  uint16_t *s_crcT = reinterpret_cast<uint16_t *>(p);             p += kCrcTableEntries * 2;
  uint32_t *s_offs = reinterpret_cast<uint32_t *>(p);             p += ((nblk_cap * 4u) + 15u) & ~15u;
  uint32_t *s_Hs = reinterpret_cast<uint32_t *>(p);               p += ((nblk_cap * 4u) + 15u) & ~15u;
  uint32_t *s_Ts = reinterpret_cast<uint32_t *>(p);               p += ((nblk_cap * 4u) + 15u) & ~15u;
  uint16_t *s_chunk = reinterpret_cast<uint16_t *>(p);            p += (((a.out_words_cap / 4u + 2u) * 2u) + 15u) & ~15u;
  uint32_t *s_misc = reinterpret_cast<uint32_t *>(p);
  // s_misc: [0..NW) warp totals, [32] next ticket, [33] payload crc, [34..36) out offset (u64), [40..46) stats adj
  uint32_t *s_words = reinterpret_cast<uint32_t *>(s_img + 32);   // payload words
  uint32_t *s_hdr = reinterpret_cast<uint32_t *>(s_img + 12);     // 5 header words

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t BL = a.P.block_len;

  for (int i = tid; i < kCrcTableEntries; i += NT) s_crcT[i] = a.crc_tables[i];
  if (tid < 6) s_misc[40 + tid] = 0;
  if (tid == 0) s_misc[32] = atomicAdd(a.ticket, 1u);
  __syncthreads();
  uint32_t f = s_misc[32];
  if (f < a.n_frames) issue_frame_load<NT>(a, f, s_in);

  uint32_t full_block_count = 0;  // lane m < 6 of every warp counts full blocks coded in mode m

  while (f < a.n_frames) {
    const unsigned long long s0 = (unsigned long long)f * a.P.spf;
    const unsigned long long remn = a.n_samples - s0;
    const uint32_t n = remn < a.P.spf ? (uint32_t)remn : a.P.spf;
    const uint32_t nblk = n > 1 ? (n - 2u) / BL + 1u : 1u;  // ceil((n-1)/BL), at least one (possibly empty) block
    const uint32_t rounds = (nblk + NT - 1) / NT;

    cp_async_wait_all();
    __syncthreads();  // (A) input staged; previous frame's image fully copied out

    uint32_t bit_base = 0;
    uint32_t f_next = a.n_frames;
    for (uint32_t r = 0; r < rounds; r++) {
      const uint32_t b = r * NT + tid;
      const bool active = b < nblk;
      const uint32_t start = 1u + b * BL;
      uint32_t len = 0;
      if (active && n > start) len = (n - start) < BL ? (n - start) : BL;

      // ---- measure ----
      BlockMode mode;
      mode.kind = kRice; mode.k = 0; mode.hdr = 0; mode.stat = 0;
      uint32_t nbits = 0;
      if (active) {
        if (len > 0) {
          mode = block_measure_generic(s_in, start, len, a.P, nbits);
        }
        if (b == 0) nbits += 16;  // <Audio State>: first sample as 16 raw bits, encoder.rs:189
      }

      // ---- statistics (stats[ftype] += block.len(), encoder.rs:199) ----
      {
        const bool full = active && len == BL;
#pragma unroll
        for (int m = 0; m < 6; m++) {
          unsigned bal = __ballot_sync(0xffffffffu, full && mode.stat == (uint32_t)m);
          if (lane == m) full_block_count += __popc(bal);
        }
        if (active && len != BL && len > 0) atomicAdd(&s_misc[40 + mode.stat], len);
      }

      // ---- exclusive scan of nbits over the CTA ----
      uint32_t incl = nbits;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_misc[wid] = incl;
      if (tid == 0 && r == rounds - 1) s_misc[32] = atomicAdd(a.ticket, 1u);  // next frame for this CTA
      __syncthreads();  // (B) warp totals visible; every thread has finished reading s_in for this round
      uint32_t wt = lane < NW ? s_misc[lane] : 0u;
      uint32_t wincl = wt;
#pragma unroll
      for (int d = 1; d < NW; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, wincl, d);
        if (lane >= d) wincl += t;
      }
      const uint32_t round_total = __shfl_sync(0xffffffffu, wincl, NW - 1);
      const uint32_t warp_base = __shfl_sync(0xffffffffu, wincl - wt, wid);
      const uint32_t bit_off = bit_base + warp_base + (incl - nbits);
      bit_base += round_total;
      if (active) s_offs[b] = bit_off;

      if (FAST && r == rounds - 1) {
        // All reads of s_in by the fast path are done (the folded differences live in registers), so the
        // next frame can stream in while this one is packed.  Blocks on the generic path re-read s_in, but
        // they only occur in the stream's final frame, after which there is nothing to prefetch.
        f_next = s_misc[32];
        if (f_next < a.n_frames) issue_frame_load<NT>(a, f_next, s_in);
      }

      // ---- pack ----
      if (active) {
        BitSink sink;
        sink.init(bit_off, s_words, &s_Hs[b]);
        if (b == 0) {
          sink.put((uint32_t)(uint16_t)s_in[0], 16);
          sink.flush();
        }
        if (len > 0) block_pack_generic(s_in, start, len, mode, sink);
        bool has_tail;
        uint32_t timg = sink.finish(has_tail);
        s_Ts[b] = timg;
      }
      if (r + 1 < rounds) __syncthreads();  // (C) s_misc warp totals are rewritten by the next round
    }
    const uint32_t total_bits = bit_base;
    const uint32_t payload_len = payload_bytes(total_bits);
    const uint32_t frame_bytes = (uint32_t)kFrameHeaderLen + payload_len;
    if (!FAST) f_next = s_misc[32];

    if (tid == 0) {
      s_offs[nblk] = total_bits;
      s_offs[nblk + 1] = 0xffffffffu;
      s_Hs[nblk] = 0u;
      s_Hs[nblk + 1] = 0u;
      // publish this frame's size so later frames can look back over it
      st_status(a.status + f, (f == 0 ? kFlagPrefix : kFlagAgg) | (unsigned long long)frame_bytes);
    }
    __syncthreads();  // (D) heads, tails, offsets visible

    // ---- merge: the owner of every partially filled word ORs in the heads that follow it ----
    for (uint32_t b = tid; b < nblk; b += NT) {
      const uint32_t end = s_offs[b + 1];
      if (end & 31u) {
        const uint32_t lw = end >> 5;
        const uint32_t o = s_offs[b];
        if ((o >> 5) < lw || (o & 31u) == 0u) {
          uint32_t v = s_Ts[b] | s_Hs[b + 1];
          for (uint32_t j = b + 2; (s_offs[j] >> 5) == lw; j++) v |= s_Hs[j];  // short blocks: several heads per word
          s_words[lw] = v;
        }
      }
    }
    __syncthreads();  // (E) payload image complete

    // ---- payload CRC, phase A: 16-byte chunks, each from a zero state (chunk 0 from 0xffff) ----
    const uint32_t m_full = payload_len >> 4;
    for (uint32_t c = tid; c < m_full; c += NT) {
      const uint4 q = reinterpret_cast<const uint4 *>(s_words)[c];
      uint32_t s = c == 0 ? 0xffffu : 0u;
      s = crc16_word(s_crcT, s, bswap32(q.x));
      s = crc16_word(s_crcT, s, bswap32(q.y));
      s = crc16_word(s_crcT, s, bswap32(q.z));
      s = crc16_word(s_crcT, s, bswap32(q.w));
      s_chunk[c] = (uint16_t)s;
    }
    __syncthreads();  // (F)

    if (wid == 0) {
      // ---- phase B: combine.  With e = m_full-1-c the distance of chunk c from the end, the state after
      // all full chunks is  sum_e chunk[e] * x^(128 e).  Lane l takes e = l, l+32, ... (Horner in x^4096),
      // then a shuffle tree multiplies by x^(128*2^k). ----
      uint32_t h = 0;
      if (m_full > (uint32_t)lane) {
        for (int i = (int)((m_full - 1u - lane) >> 5); i >= 0; i--) {
          const uint32_t c = m_full - 1u - (32u * (uint32_t)i + lane);
          h = crc16_mulc(s_crcT, 4, h) ^ (uint32_t)s_chunk[c];
        }
      }
#pragma unroll
      for (int k = 0; k < 5; k++) {
        uint32_t o = __shfl_down_sync(0xffffffffu, h, 1 << k);
        h ^= crc16_mulc(s_crcT, 6 + 2 * k, o);
      }
      if (lane == 0) {
        uint32_t s = m_full ? h : 0xffffu;
        const uint32_t rem = payload_len & 15u;  // even
        uint32_t wi = m_full * 4u;
        for (uint32_t done = 0; done + 4u <= rem; done += 4u) s = crc16_word(s_crcT, s, bswap32(s_words[wi++]));
        if (rem & 2u) s = crc16_half(s_crcT, s, bswap32(s_words[wi]) >> 16);
        // ---- frame header, encoder.rs:122-162 (id = 1 for audio frames, encoder.rs:210) ----
        const uint32_t hc = header_crc(s_crcT, 1u, n, payload_len);
        s_hdr[0] = bswap32((kFrameKey << 16) | 0x0101u);
        s_hdr[1] = bswap32(((n & 0xffffu) << 16) | (payload_len & 0xffffu));
        s_hdr[2] = 0u;
        s_hdr[3] = 0u;
        s_hdr[4] = bswap32((hc << 16) | (s & 0xffffu));
      }
    } else if (wid == 1) {
      // ---- decoupled look-back over frame sizes -> this frame's byte offset ----
      const unsigned long long excl = lookback_exclusive(a.status, f, frame_bytes);
      if (lane == 0) {
        s_misc[34] = (uint32_t)excl;
        s_misc[35] = (uint32_t)(excl >> 32);
        if (f == a.n_frames - 1) a.result[0] = excl + frame_bytes;  // total stream length
      }
    }
    __syncthreads();  // (G) header + offset ready

    // ---- copy the frame image to its place in the output stream ----
    {
      const unsigned long long off = (unsigned long long)s_misc[34] | ((unsigned long long)s_misc[35] << 32);
      if (off + frame_bytes <= a.out_cap) {
        unsigned char *dst = a.out + off;
        const uint32_t *src32 = s_hdr;
        const uint32_t L = frame_bytes;
        const uintptr_t al = (uintptr_t)dst & 3u;
        if (al == 0) {
          uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
          const uint32_t nw = L >> 2;
          for (uint32_t i = tid; i < nw; i += NT) d32[i] = src32[i];
          if ((L & 2u) && tid == 0)
            *reinterpret_cast<uint16_t *>(dst + (nw << 2)) = (uint16_t)(src32[nw] & 0xffffu);
        } else if (al == 2) {
          if (tid == 0) *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(src32[0] & 0xffffu);
          uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + 2);
          const uint32_t nw = (L - 2u) >> 2;
          for (uint32_t i = tid; i < nw; i += NT) d32[i] = __byte_perm(src32[i], src32[i + 1], 0x5432);
          if (((L - 2u) & 2u) && tid == 0)
            *reinterpret_cast<uint16_t *>(dst + 2 + (nw << 2)) = (uint16_t)(src32[nw] >> 16);
        } else {
          const unsigned char *sb = reinterpret_cast<const unsigned char *>(s_hdr);
          for (uint32_t i = tid; i < L; i += NT) dst[i] = sb[i];
        }
      } else if (tid == 0) {
        atomicMax(a.result + 1, 1ull);  // ByteWriterInsufficientMemory, bytewriter.rs:88-90
      }
    }

    f = f_next;
    if (!FAST && f < a.n_frames) {
      __syncthreads();  // generic path re-reads s_in while packing, so only now may it be overwritten
      issue_frame_load<NT>(a, f, s_in);
    }
  }

  // ---- flush statistics ----
  if (lane < 6 && full_block_count) atomicAdd(a.result + 2 + lane, (unsigned long long)full_block_count * BL);
  __syncthreads();
  if (tid < 6 && s_misc[40 + tid]) atomicAdd(a.result + 2 + tid, (unsigned long long)s_misc[40 + tid]);
}


// ------------------------------------------------------------------------------------------------
// Fast kernel: Parameters::default() and at most 512 blocks per frame.
//
// 16 worker warps (one thread per block) and 1 control warp that runs ASYNCHRONOUSLY: the two sides meet only
// through named barriers used as producer/consumer signals (bar.arrive / bar.sync) and one mbarrier per ring slot,
// and the frame image lives in a ring of NB = 3 buffers, so a frame's byte offset (which needs every earlier frame's
// size) has two frame times to arrive before anybody waits for it.  CTA 0 is the scanner (scanner_role).
//
//   workers, frame i:  wait for the staged samples (mbarrier of the TMA copy) -> measure + CTA scan -> publish the
//                      size, [size_ready(i)] -> pack into image[i % NB] (plain stores; a block's last partial word
//                      stays in a register) -> wait [off_ready(i-2)], copy frame i-2's payload out -> barrier ->
//                      OR the partial words in -> barrier -> start the TMA copy of the next frame -> per-warp CRC
//                      slices -> [crc_ready(i)].
//   control:           on [size_ready(i)]: finish frame i-1 -- poll its prefix (there by now), wait [crc_ready(i-1)],
//                      fold the slice CRCs, build and write the 20-byte header (encoder.rs:122-162), [off_ready(i-1)].
// ------------------------------------------------------------------------------------------------
constexpr int NTF = kEncFastThreads;      // 544
constexpr int NWW = 16;                   // worker warps
constexpr int kMaxSlices = 64;
constexpr uint32_t kNoFrame = 0xffffffffu;
constexpr uint32_t kStdSpf = 10000, kStdOwc = 5096;  // Parameters::default(): 20 x 500 samples, 16 + 500*326 bits
constexpr uint32_t NB = 3;                // frame images in flight per CTA (slack for the look-back: NB-1 frames)

// Frame stager of the fast kernel: ONE bulk async copy (TMA, cp.async.bulk) of the frame's PCM into shared memory,
// issued by a single thread and completed on an mbarrier that every worker waits on.  The API guarantees a 16-byte
// aligned PCM base and frame size for this kernel.  (Per-thread 16-byte cp.async cost ~30 instructions per thread
// and frame in address arithmetic.)  Samples beyond the last whole 16 bytes -- only in the stream's last frame --
// are copied by stage_tail_fast with plain loads.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "X3_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra X3_MBAR_DONE;\n"
      "bra X3_MBAR_WAIT;\n"
      "X3_MBAR_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
// called by one thread
__device__ __forceinline__ void stage_frame_bulk(const int16_t *pcm, unsigned long long s0, uint32_t n, int16_t *s_in, uint32_t mbar) {
  const uint32_t bytes = (n >> 3) << 4;
  if (bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(s_in)),
                 "l"(pcm + s0), "r"(bytes), "r"(mbar)
                 : "memory");
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(mbar) : "memory");
  }
}
// called by all 512 workers; a barrier must follow before s_in is read
__device__ __forceinline__ void stage_tail_fast(const int16_t *pcm, unsigned long long s0, uint32_t n, int16_t *s_in, int tid) {
  for (uint32_t i = (n & ~7u) + tid; i < n; i += 512u) s_in[i] = __ldg(pcm + s0 + i);
}

#ifdef X3_ENC_TIMING
#define X3_T(k) { const long long now__ = clock64(); tacc[k] += now__ - tlast; tlast = now__; }
#else
#define X3_T(k)
#endif

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }
// the barrier id is a register operand: ptxas then reserves all 16 barriers of the CTA, which still allows
// 4 CTAs per SM (2 are resident); a switch over immediate ids cost ~30 instructions per frame and thread
__device__ __forceinline__ void bar_sync_all(uint32_t id) { asm volatile("bar.sync %0, 544;\n" ::"r"(id) : "memory"); }
// bar.arrive orders this thread's prior shared-memory writes for the threads that bar.sync on the same barrier
// (the PTX producer/consumer idiom); an explicit MEMBAR here cost ~600 cycles per frame
__device__ __forceinline__ void bar_arrive_all(uint32_t id) { asm volatile("bar.arrive %0, 544;\n" ::"r"(id) : "memory"); }
constexpr int kBarSize = 2, kBarCrc = 5;  // + buffer index (0..NB-1)
// "Offset and header of the frame in ring slot q are ready" is an mbarrier per slot (one arrival by the control warp
// per use of the slot, the workers wait on the phase parity of that use): unlike a named barrier it does not make
// the sixteen worker warps meet, so they run from the CRC slices into the copy-out and the next frame's measure
// phase on their own.
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(mbar) : "memory");
}

// payload image (16-byte aligned in shared memory) -> dst (2-byte aligned global address), 512 worker threads.
// Body in 16-byte stores; the shared-memory side is read at a 2- or 4-byte skew and realigned with PRMT.
__device__ __forceinline__ void copy_payload_out(unsigned char *dst, const uint32_t *s_words, uint32_t L, int tid) {
  if (((uintptr_t)dst & 1u) != 0) {  // odd output base: bytewise
    const unsigned char *sb = reinterpret_cast<const unsigned char *>(s_words);
    for (uint32_t i = tid; i < L; i += 512) dst[i] = sb[i];
    return;
  }
  uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);  // bytes until dst is 16-byte aligned (even)
  if (head > L) head = L;
  const uint32_t nvec = (L - head) >> 4;
  const uint32_t tail0 = head + (nvec << 4);
  const uint16_t *s16 = reinterpret_cast<const uint16_t *>(s_words);
  // head and tail halfwords (at most 7 + 7)
  if ((uint32_t)tid < (head >> 1)) reinterpret_cast<uint16_t *>(dst)[tid] = s16[tid];
  if ((uint32_t)tid >= 32u && (uint32_t)tid - 32u < ((L - tail0) >> 1))
    reinterpret_cast<uint16_t *>(dst + tail0)[tid - 32] = s16[(tail0 >> 1) + (tid - 32)];
  uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
  const uint32_t w0 = head >> 2;          // first source word
  if ((head & 3u) == 0u) {
    for (uint32_t i = tid; i < nvec; i += 512) {
      const uint32_t *q = s_words + w0 + 4u * i;
      uint4 v;
      v.x = q[0]; v.y = q[1]; v.z = q[2]; v.w = q[3];
      d4[i] = v;
    }
  } else {                                // source starts in the middle of a word
    for (uint32_t i = tid; i < nvec; i += 512) {
      const uint32_t *q = s_words + w0 + 4u * i;
      const uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
      uint4 v;
      v.x = __byte_perm(a0, a1, 0x5432); v.y = __byte_perm(a1, a2, 0x5432);
      v.z = __byte_perm(a2, a3, 0x5432); v.w = __byte_perm(a3, a4, 0x5432);
      d4[i] = v;
    }
  }
}

// Scanner role (CTA 0, one warp): turns the frame sizes the workers publish (kFlagAgg | bytes) into inclusive
// prefixes (kFlagPrefix | bytes through this frame), in order, up to 256 frames per round trip.  Every frame's
// control warp then polls only its own status word.  (With a per-frame decoupled look-back every one of the
// ~300 resident CTAs walks the same few hundred entries again and again; one scanner does that work once.)
__device__ __forceinline__ void scanner_role(const EncodeArgs &a) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  constexpr int G = 8;
  unsigned long long frontier = 0, running = 0;
  const unsigned long long nf = a.n_frames;
  while (frontier < nf) {
    unsigned long long v[G];
    const unsigned long long base_idx = frontier + (unsigned long long)(G * lane);
#pragma unroll
    for (int g = 0; g < G; g++) v[g] = base_idx + g < nf ? ld_status(a.status + base_idx + g) : 0ull;
    int ready = 0;  // leading published entries of this lane
    unsigned long long incl[G], run = 0;
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (ready == g && (v[g] >> 62) != 0ull) ready = g + 1;
      run += v[g] & kValueMask;
      incl[g] = run;
    }
    const unsigned notfull = __ballot_sync(0xffffffffu, ready < G);
    const int L = notfull ? __ffs((int)notfull) - 1 : 32;   // first lane that is not completely ready
    if (lane > L) ready = 0;
    const int total_ready = __reduce_add_sync(0xffffffffu, ready);
    if (total_ready == 0) { __nanosleep(100); continue; }
    unsigned long long t = ready ? incl[ready - 1 < G ? ready - 1 : G - 1] : 0ull;
    {
      // pick incl[ready-1] without dynamic register indexing
      unsigned long long pick = 0;
#pragma unroll
      for (int g = 0; g < G; g++) if (ready == g + 1) pick = incl[g];
      t = pick;
    }
    unsigned long long scan = t;  // inclusive scan of lane totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, scan, d);
      if (lane >= d) scan += o;
    }
    const unsigned long long lane_base = running + scan - t;
#pragma unroll
    for (int g = 0; g < G; g++)
      if (g < ready) st_status(a.status + base_idx + g, kFlagPrefix | ((lane_base + incl[g]) & kValueMask));
    running += __shfl_sync(0xffffffffu, scan, 31);
    frontier += (unsigned long long)total_ready;
  }
  if (lane == 0) a.result[0] = running;  // total stream length
}

// SPF / OWC != 0: samples per frame and image words are compile-time constants (Parameters::default(), 500 blocks
// per frame), which turns all shared-memory address arithmetic into immediates; 0 = take them from the arguments.
template <uint32_t SPF, uint32_t OWC>
__global__ void __launch_bounds__(NTF, 2) encode_frames_fast_kernel(const __grid_constant__ EncodeArgs a) {
  if (a.choice && *a.choice != (unsigned)kEncKernelFast) return;   // the probe picked the strip kernel
  if (blockIdx.x == 0) {
    scanner_role(a);
    return;
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t spf = SPF ? SPF : a.P.spf;
  const uint32_t owc = OWC ? OWC : a.out_words_cap;
  const uint32_t in_bytes = (2u * (spf + 8u) + 15u) & ~15u;
  const uint32_t img_bytes = (4u * (owc + 8u) + 15u) & ~15u;
  unsigned char *p = smem_raw;
  int16_t *s_in = reinterpret_cast<int16_t *>(p);                 p += in_bytes;
  uint32_t *s_img = reinterpret_cast<uint32_t *>(p);              p += NB * img_bytes;
  const uint32_t img_words = img_bytes >> 2;
  uint16_t *s_crcT = reinterpret_cast<uint16_t *>(p);             p += kCrcBankEntries2 * 2;
  const uint16_t *s_crcT2 = s_crcT + kCrcTableEntries;            // byte-swapped bank (worker CRC)
  uint32_t *s_V = reinterpret_cast<uint32_t *>(p);                p += NB * kMaxSlices * 4;
  uint32_t *s_misc = reinterpret_cast<uint32_t *>(p);
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(s_misc + 80);  // 8 bytes: the frame stager's mbarrier
  const uint32_t mbar_off = mbar + 8u;                                    // NB x 8 bytes: "offset ready" per ring slot
  // s_misc: [0..16) warp totals, [32] next ticket, [40..46) stats of short blocks,
  //         per parity q at [48+8q ..): +0 frame, +1 samples, +2 payload_len, +4,+5 byte offset, +6 fits

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool worker = wid < NWW;
  constexpr uint32_t BL = 20;

  for (int i = tid; i < kCrcBankEntries2; i += NTF) s_crcT[i] = a.crc_tables[i];
  if (tid < 6) s_misc[40 + tid] = 0;
  if (tid == 0) {
    mbar_init(mbar, 1);
    for (uint32_t q = 0; q < NB; q++) mbar_init(mbar_off + 8u * q, 1);
  }
  __syncthreads();
  // Frames are handed out by an atomic ticket, LATE: a CTA draws its next frame only when it has finished packing
  // the current one, so that a ticketed frame publishes its size within about one measure phase.  (Drawing the
  // ticket a whole frame early -- to prefetch -- leaves unmeasured predecessors in everybody's look-back window
  // for a frame time; dealing frames round-robin makes every frame wait for the slowest CTA of its round.)

  if (!worker) {
    // ================================ control warp ================================
#ifdef X3_ENC_TIMING
    long long tacc[12] = {0}, tlast = clock64();
    unsigned long long cframes = 0;
#endif
    // The control warp finishes frame i-1 (offset, CRC fold, header) when frame i's size signal arrives: by then
    // frame i-1's size has been public for a whole frame time, so its prefix is normally there at the first poll,
    // and the warp is never the serial bottleneck (poll + fold used to take as long as the workers take per frame).
    bool have_prev = false;
    uint32_t prev_par = 0;
    for (uint32_t cur = 0;; cur = (cur + 1u == NB ? 0u : cur + 1u)) {
      bar_sync_all(kBarSize + cur);
      X3_T(0)
      const bool stop = s_misc[48 + 8 * cur] == kNoFrame;
      const uint32_t par = prev_par;
      const bool do_prev = have_prev;
      have_prev = true;
      prev_par = cur;
      if (!do_prev) {
        if (stop) break;
        continue;
      }
      uint32_t *info = s_misc + 48 + 8 * par;
      const uint32_t f = info[0];
      const uint32_t n = info[1], payload_len = info[2];
      const uint32_t frame_bytes = (uint32_t)kFrameHeaderLen + payload_len;
      X3_T(1)
      // this frame's byte offset: wait for the scanner to turn the published size into a prefix
      unsigned long long pv = 0;
      if (lane == 0) {
        while (((pv = ld_status(a.status + f)) >> 62) != 2ull) __nanosleep(64);
      }
      pv = __shfl_sync(0xffffffffu, pv, 0);
      const unsigned long long excl = (pv & kValueMask) - frame_bytes;
      const bool fits = excl + frame_bytes <= a.out_cap;
      if (lane == 0 && !fits) atomicMax(a.result + 1, 1ull);         // ByteWriterInsufficientMemory, bytewriter.rs:88
      X3_T(2)
      bar_sync_all(kBarCrc + par);
      X3_T(3)
      // payload CRC = sum_j V_j * x^(4096 j), then the tail bytes, then the header
      const uint32_t *s_words = s_img + par * img_words;
      const uint32_t *V = s_V + par * kMaxSlices;
      uint32_t hw = 0;
      if (lane == 0) {
        const uint32_t m = payload_len >> 4;
        const uint32_t nslices = (m + 31u) >> 5;
        uint32_t s = 0;
        for (int j = (int)nslices - 1; j >= 0; j--) s = crc16_mulc(s_crcT, 4, s) ^ V[j];
        if (m == 0) s = 0xffffu;
        const uint32_t rem = payload_len & 15u;  // even
        uint32_t wi = m * 4u;
        for (uint32_t done = 0; done + 4u <= rem; done += 4u) s = crc16_word(s_crcT, s, bswap32(s_words[wi++]));
        if (rem & 2u) s = crc16_half(s_crcT, s, bswap32(s_words[wi]) >> 16);
        hw = (header_crc(s_crcT, 1u, n, payload_len) << 16) | (s & 0xffffu);
        info[4] = (uint32_t)excl;
        info[5] = (uint32_t)(excl >> 32);
        info[6] = fits ? 1u : 0u;
      }
      hw = __shfl_sync(0xffffffffu, hw, 0);
      if (fits && lane < 10) {
        // header halfwords, big-endian values (id = 1 for audio frames, encoder.rs:210; time = 0, :148-150)
        uint32_t v = 0;
        if (lane == 0) v = kFrameKey;
        else if (lane == 1) v = 0x0101u;
        else if (lane == 2) v = n & 0xffffu;
        else if (lane == 3) v = payload_len & 0xffffu;
        else if (lane == 8) v = hw >> 16;
        else if (lane == 9) v = hw & 0xffffu;
        reinterpret_cast<uint16_t *>(a.out + excl)[lane] = (uint16_t)(((v & 0xff) << 8) | (v >> 8));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(mbar_off + 8u * par);
      X3_T(4)
#ifdef X3_ENC_TIMING
      cframes++;
#endif
      if (stop) break;
    }
#ifdef X3_ENC_TIMING
    if (lane == 0) {
      for (int k = 0; k < 5; k++) atomicAdd(a.timing + 12 + k, (unsigned long long)tacc[k]);
      atomicAdd(a.timing + 17, cframes);
    }
#endif
    return;
  }

  // ================================== workers ==================================
  // samples of frame f: every frame is full except possibly the last (32-bit compare instead of 64-bit arithmetic)
  const uint32_t last_f = a.n_frames - 1u;
  const uint32_t last_n = (uint32_t)(a.n_samples - (unsigned long long)last_f * spf);
  if (tid == 0) s_misc[32] = atomicAdd(a.ticket, 1u);
  bar_workers();
  uint32_t f = s_misc[32];
  if (f < a.n_frames) {
    const unsigned long long s00 = (unsigned long long)f * spf;
    const uint32_t n0 = f == last_f ? last_n : spf;
    if (tid == 0) stage_frame_bulk(a.pcm, s00, n0, s_in, mbar);
    if (n0 & 7u) stage_tail_fast(a.pcm, s00, n0, s_in, tid);
  }
  uint32_t it = 0, par = 0, crc_phase = 0;
  unsigned long long stat_acc = 0;
#ifdef X3_ENC_TIMING
  long long tacc[12] = {0}, tlast = clock64();
#endif

  while (f != kNoFrame && f < a.n_frames) {
    uint32_t *s_words = s_img + par * img_words;
    const uint32_t n = f == last_f ? last_n : spf;
    const uint32_t nblk = n > 1 ? (n - 2u) / BL + 1u : 1u;  // <= 512

    if (n & 7u) bar_workers();  // the last frame's tail samples were stored by other threads
    mbar_wait(mbar, it & 1u);   // (A) samples staged (bulk copy complete)
    X3_T(0)

    // ---- measure ----
    const uint32_t b = tid;
    const bool active = b < nblk;
    const uint32_t start = 1u + b * BL;
    uint32_t len = 0;
    if (active && n > start) len = (n - start) < BL ? (n - start) : BL;
    FastBlock fb;
    BlockMode mode;
    mode.kind = kRice; mode.k = 0; mode.hdr = 0; mode.stat = 0;
    uint32_t nbits = 0;
    bool use_fast = false;
    if (active) {
      if (len >= BL - 1) {
        use_fast = true;
        mode = block_measure_fast(s_in, start, len, fb, nbits, a.neg_one);
      } else if (len > 0) {
        mode = block_measure_generic(s_in, start, len, a.P, nbits);  // short last block of the stream's last frame
      }
      if (b == 0) nbits += 16;  // <Audio State>, encoder.rs:189
    }
    // statistics (stats[ftype] += block.len(), encoder.rs:199): every thread counts its own full blocks per mode;
    // the rare short block adds its length directly
    if (active && len == BL) stat_acc += 1ull << (10u * mode.stat);   // six 10-bit block counters per thread
    if (active && len != BL && len > 0) atomicAdd(&s_misc[40 + mode.stat], len);
    uint32_t incl = nbits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_misc[wid] = incl;
    X3_T(1)
    bar_workers();  // (B) warp totals visible; all reads of s_in by the fast path are done

    X3_T(2)
    const uint32_t wt = lane < NWW ? s_misc[lane] : 0u;
    uint32_t wincl = wt;
#pragma unroll
    for (int d = 1; d < NWW; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, d);
      if (lane >= d) wincl += t;
    }
    const uint32_t total_bits = __shfl_sync(0xffffffffu, wincl, NWW - 1);
    const uint32_t payload_len = payload_bytes(total_bits);
    if (tid == 0) {
      uint32_t *info = s_misc + 48 + 8 * par;
      info[0] = f; info[1] = n; info[2] = payload_len;
      // publish the frame size right away (the control warp may still be busy with the previous frame)
      st_status(a.status + f, kFlagAgg | (unsigned long long)(kFrameHeaderLen + payload_len));
    }
    bar_arrive_all(kBarSize + par);  // -> control: frame size known
    X3_T(3)

    // ---- prefetch the next frame, pack this one ----
    const uint32_t warp_base = __shfl_sync(0xffffffffu, wincl - wt, wid);
    const uint32_t bit_off = warp_base + (incl - nbits);
    FastSink sink;
    uint32_t tail = 0;
    sink.cnt = 0;
    if (active) {
      sink.init(bit_off, s_words);
      if (b == 0) {
        sink.put((uint32_t)(uint16_t)(use_fast ? fb.pred : (int32_t)s_in[0]), 16);
        sink.flush();
      }
      if (use_fast) block_pack_fast(fb, len, mode, sink);
      else if (len > 0) block_pack_generic(s_in, start, len, mode, sink);
      tail = sink.finish_tail();   // this block's last, partial word: ORed in after barrier (D)
    }
    if (tid == 0) {
      s_misc[32] = atomicAdd(a.ticket, 1u);  // this CTA's next frame
      // the frame's last word is completed by nobody: zero it for the tails that are ORed into it (and for the padding)
      if (total_bits & 31u) s_words[total_bits >> 5] = 0u;
    }
    // ---- the oldest frame in the ring goes out now, before barrier (D): a warp that is done packing copies its share
    // instead of waiting for the slowest one.  (Its offset has had NB-1 frame times to arrive; the slot is written
    // again only by the next frame's pack, after that frame's barrier (B).) ----
    if (it >= NB - 1) {
      const uint32_t q = par + 1u == NB ? 0u : par + 1u;  // the buffer the next frame will reuse
      mbar_wait(mbar_off + 8u * q, ((it - (NB - 1u)) / NB) & 1u);   // frame it-(NB-1) of this CTA was the slot's use number (it-(NB-1))/NB
      X3_T(9)
      const uint32_t *info = s_misc + 48 + 8 * q;
      const uint32_t L = info[2];  // payload length (stays valid until this buffer's next measure phase)
      if (info[6]) {
        const unsigned long long off = (unsigned long long)info[4] | ((unsigned long long)info[5] << 32);
        copy_payload_out(a.out + off + kFrameHeaderLen, s_img + q * img_words, L, tid);
      }
      X3_T(10)
    }
    X3_T(4)
    bar_workers();  // (D) every plain store done
    X3_T(5)
    uint32_t f_next = s_misc[32];
    if (f_next >= a.n_frames) f_next = kNoFrame;
    if (sink.cnt) sm_or(sink.dst, tail);
    X3_T(6)
    bar_workers();  // (E) payload image complete
    X3_T(7)
    // stage the next frame now: s_in has been free since barrier (B), and the copy lands while the CRC and
    // the copy-out run
    if (f_next != kNoFrame) {
      const unsigned long long s1 = (unsigned long long)f_next * spf;
      const uint32_t n1 = f_next == last_f ? last_n : spf;
      if (tid == 0) stage_frame_bulk(a.pcm, s1, n1, s_in, mbar);
      if (n1 & 7u) stage_tail_fast(a.pcm, s1, n1, s_in, tid);
    }

    // ---- CRC of 16-byte chunks, combined per slice of 32 chunks by a shuffle tree.  Slice j covers the
    // chunks at distance 32j .. 32j+31 from the end; V_j = sum_l x^(128 l) * crc(chunk at distance 32j+l). ----
    {
      const uint32_t m = payload_len >> 4;
      const uint32_t nslices = (m + 31u) >> 5;
      uint32_t *V = s_V + par * kMaxSlices;
      for (uint32_t j = wid; j < nslices; j += NWW) {
        const uint32_t e = 32u * j + lane;
        uint32_t h = 0;
        if (e < m) {
          const uint32_t c = m - 1u - e;
          const uint4 q = reinterpret_cast<const uint4 *>(s_words)[c];
          h = c == 0 ? 0xffffu : 0u;  // chunk 0 carries the CRC's initial value (byte-swapped state form)
          h = crc16_word_sw(s_crcT2, h, q.x);
          h = crc16_word_sw(s_crcT2, h, q.y);
          h = crc16_word_sw(s_crcT2, h, q.z);
          h = crc16_word_sw(s_crcT2, h, q.w);
        }
        h ^= crc16_mulc_sw<6>(s_crcT2, __shfl_down_sync(0xffffffffu, h, 1));
        h ^= crc16_mulc_sw<8>(s_crcT2, __shfl_down_sync(0xffffffffu, h, 2));
        h ^= crc16_mulc_sw<10>(s_crcT2, __shfl_down_sync(0xffffffffu, h, 4));
        h ^= crc16_mulc_sw<12>(s_crcT2, __shfl_down_sync(0xffffffffu, h, 8));
        h ^= crc16_mulc_sw<14>(s_crcT2, __shfl_down_sync(0xffffffffu, h, 16));
        if (lane == 0) V[j] = bswap16(h);
      }
    }
    bar_arrive_all(kBarCrc + par);  // -> control: slice CRCs and image complete
    X3_T(8)

    f = f_next;
    it++;
    par = par + 1u == NB ? 0u : par + 1u;
    if ((it & 1023u) == 1023u) {  // the 10-bit fields are about to fill up
#pragma unroll
      for (int m = 0; m < 6; m++) {
        const uint32_t c = (uint32_t)(stat_acc >> (10 * m)) & 1023u;
        if (c) atomicAdd(&s_misc[40 + m], c * BL);
      }
      stat_acc = 0;
    }
  }
  // ---- drain: tell the control warp to stop (it finishes the last frame on this signal), then copy out the
  // frames still in the ring, oldest first ----
  {
    if (tid == 0) s_misc[48 + 8 * par] = kNoFrame;
    bar_arrive_all(kBarSize + par);
    const uint32_t pending = it < NB - 1 ? it : NB - 1;
    uint32_t q = (par + NB - pending) % NB;
    for (uint32_t d = 0; d < pending; d++) {
      mbar_wait(mbar_off + 8u * q, ((it - pending + d) / NB) & 1u);
      const uint32_t *info = s_misc + 48 + 8 * q;
      const uint32_t L = info[2];  // payload length (stays valid until this buffer's next measure phase)
      if (info[6]) {
        const unsigned long long off = (unsigned long long)info[4] | ((unsigned long long)info[5] << 32);
        copy_payload_out(a.out + off + kFrameHeaderLen, s_img + q * img_words, L, tid);
      }
      q = q + 1u == NB ? 0u : q + 1u;
    }
  }
#ifdef X3_ENC_TIMING
  if (tid == 0) {
    for (int k = 0; k < 11; k++) atomicAdd(a.timing + k, (unsigned long long)tacc[k]);
    atomicAdd(a.timing + 11, (unsigned long long)it);
  }
#endif
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const uint32_t c = (uint32_t)(stat_acc >> (10 * m)) & 1023u;
    if (c) atomicAdd(&s_misc[40 + m], c * BL);
  }
  bar_workers();
  if (tid < 6 && s_misc[40 + tid]) atomicAdd(a.result + 2 + tid, (unsigned long long)s_misc[40 + tid]);
}


// ------------------------------------------------------------------------------------------------
// Strip kernel: Parameters::default() and at most 512 blocks per frame (x3_enc_strip.cuh has the per-thread logic).
//
// 128 threads (4 warps) per CTA, one frame at a time, 5 CTAs per SM; one thread owns four consecutive blocks.  CTA 0 is
// the scanner (scanner_role) that turns published frame sizes into stream offsets.  Per frame:
//   stage    -- the frame's PCM lands in 176-byte rows (one per thread); every warp stages the rows of its own threads
//               (issued while the previous frame was finishing), so nothing but the warp itself waits for them;
//   pack     -- every thread codes its strip in place in its own row (single pass, nothing shared);
//   scan     -- CTA scan of the strips' bit counts (barrier 1); thread 0 publishes the frame size -- and nobody waits for
//               the frame's offset;
//   relocate -- strips are shifted into one of TWO windows, the finished payload image (barrier 2);
//   crc      -- 32-byte chunks of the window in parallel (slicing-by-4 tables in shared memory, byte-swapped state),
//               32 chunks folded per warp by a shuffle tree; warp 0 alone waits for the four warps' slices (named
//               barrier) and finishes: slices by Horner, tail, header CRC;
//   out      -- the PREVIOUS frame's window goes to the stream now: its offset (which needs every earlier frame's size)
//               has had a whole frame time to arrive.  16-byte stores, realigned with PRMT (offsets are only even).
// A payload of 8 .. 16 KiB (BFP / literal heavy frames) takes both windows as one buffer, after the pending frame has
// gone out; only a payload larger than that (literal almost throughout) takes slow_frame(): merging relocation in
// rounds, written out at once, which needs the frame's own offset.
// ------------------------------------------------------------------------------------------------
constexpr int NTS = kEncStripThreads;                 // 128
constexpr uint32_t kWinBytes = 8192;                  // multiple of 1024 (32 chunks of 32 bytes)
constexpr uint32_t kWinWords = kWinBytes / 4;
constexpr uint32_t kWinSlackWords = 8;
constexpr uint32_t kWinStride = kWinWords + kWinSlackWords;
constexpr uint32_t kWinChunks = kWinBytes / 32;
constexpr uint32_t kMaxSlicesStrip = 32;              // 0x7fe0 / 1024 rounded up
constexpr uint32_t kRowsBytes = kStripMaxRows * kRowWords * 4;

// rows of this warp's 32 threads <- global (5120 contiguous bytes of the frame); n = samples of the frame.
// Coalesced 16-byte cp.async: chunk c of the warp's region goes to row c / 10, behind the row's 16 bytes of padding.
// (One 160-byte bulk async copy per row, issued by lane 0 onto a per-warp mbarrier, was measured: 1.56 -> 1.93 ms.)
__device__ __forceinline__ void stage_rows(const int16_t *frame, uint32_t n, uint32_t *s_rows, uint32_t *s_next, int wid,
                                           int lane) {
  const uint32_t w0 = (uint32_t)wid * 32u * kStripSamples;              // first sample of the warp's rows
  const unsigned char *src = reinterpret_cast<const unsigned char *>(frame + w0);
  unsigned char *dst = reinterpret_cast<unsigned char *>(s_rows + (uint32_t)wid * 32u * kRowWords + kRowPadWords);
  const uint32_t avail = n > w0 ? n - w0 : 0u;                          // samples of the frame from w0 on
  const uint32_t chunks = avail >> 3;                                   // whole 16-byte chunks
#pragma unroll
  for (uint32_t i = 0; i < 10u; i++) {
    const uint32_t c = 32u * i + (uint32_t)lane;                        // chunk of the warp's region; row c / 10
    if (c < chunks) cp_async16(dst + 16u * (c + ((c * 205u) >> 11)), src + 16u * c);
  }
  // the 81st sample of the warp's last strip: first chunk of the next warp's region
  if (lane == 0 && chunks > 320u) cp_async16(s_next + 4 * wid, src + 5120u);
  // samples after the last whole chunk (stream's last frame only)
  if (avail < 32u * kStripSamples + 8u) {
    for (uint32_t i = (chunks << 3) + (uint32_t)lane; i < avail && i < 32u * kStripSamples + 8u; i += 32u) {
      const int16_t v = __ldg(frame + w0 + i);
      if (i < 32u * kStripSamples) {
        const uint32_t r = i / kStripSamples, k = i % kStripSamples;
        reinterpret_cast<int16_t *>(s_rows + ((uint32_t)wid * 32u + r) * kRowWords + kRowPadWords)[k] = v;
      } else {
        reinterpret_cast<int16_t *>(s_next + 4 * wid)[i - 32u * kStripSamples] = v;
      }
    }
  }
  cp_async_commit();
}

// CRC of the whole 32-byte chunks [c_lo, c_hi) of a payload that sit in `win` from chunk c_lo on; nch = whole chunks of
// the payload.  Slice j = the 32 chunks at distance 32j .. 32j+31 from the last whole chunk;
// V_j ^= sum_l x^(256 l) * crc(chunk at distance 32j + l).  Called by all four warps.
// A chunk is summed as two independent 16-byte halves (half the dependent table look-ups in a row; the first half is
// then multiplied by x^128); lane l multiplies its chunk's sum by x^(256 l) = x^(256 (l & 3)) * x^(1024 (l >> 2)) with
// its OWN two nibble tables (NA, NB), and the 32 products are XORed by one warp reduction: no shuffle tree.
__device__ __forceinline__ void crc_slices(const uint32_t *win, uint32_t c_lo, uint32_t c_hi, uint32_t nch, uint32_t *s_V,
                                           const uint16_t *s_T2, const uint16_t *s_N, const uint16_t *NA, const uint16_t *NB,
                                           int wid, int lane) {
  if (c_hi <= c_lo) return;
  const uint32_t j_lo = (nch - c_hi) >> 5, j_hi = (nch - 1u - c_lo) >> 5;
  // slices are dealt from warp 3 down: a fifth slice goes to the warp that is last in the next frame's scan chain
  for (uint32_t j = j_lo + (3u - (uint32_t)wid); j <= j_hi; j += 4u) {
    const uint32_t e = 32u * j + (uint32_t)lane;
    uint32_t h = 0;
    if (e < nch) {
      const uint32_t c = nch - 1u - e;
      if (c >= c_lo && c < c_hi) {
        const uint4 *q = reinterpret_cast<const uint4 *>(win) + 2u * (c - c_lo);
        const uint4 q0 = q[0], q1 = q[1];
        uint32_t h0 = c == 0 ? 0xffffu : 0u, h1 = 0u;   // the CRC's initial value (byte-swapped state form)
        h0 = crc16_word_sw(s_T2, h0, q0.x);
        h1 = crc16_word_sw(s_T2, h1, q1.x);
        h0 = crc16_word_sw(s_T2, h0, q0.y);
        h1 = crc16_word_sw(s_T2, h1, q1.y);
        h0 = crc16_word_sw(s_T2, h0, q0.z);
        h1 = crc16_word_sw(s_T2, h1, q1.z);
        h0 = crc16_word_sw(s_T2, h0, q0.w);
        h1 = crc16_word_sw(s_T2, h1, q1.w);
        h = crc16_mul_nib(s_N + 64 * kCrcMulX128, h0) ^ h1;
        h = crc16_mul_nib(NB, crc16_mul_nib(NA, h));
      }
    }
    h = __reduce_xor_sync(0xffffffffu, h);
    if (lane == 0) s_V[j] ^= h & 0xffffu;
  }
}

// One lane: payload CRC from the slice sums (Horner in x^8192) and the bytes after the last whole chunk,
// which lie in `win` at byte `tail_at`; returns (header CRC << 16) | payload CRC for the frame header.
__device__ __noinline__ uint32_t crc_finish(const uint32_t *s_V, const uint32_t *win, uint32_t tail_at, uint32_t payload_len,
                                            uint32_t n, const uint16_t *s_T2, const uint16_t *s_N) {
  const uint32_t nch = payload_len >> 5, nsl = (nch + 31u) >> 5;
  uint32_t s = 0;
#pragma unroll 1
  for (int j = (int)nsl - 1; j >= 0; j--) s = crc16_mul_nib(s_N + 64 * kCrcMulX8192, s) ^ s_V[j];
  if (nch == 0) s = 0xffffu;
  uint32_t rem = payload_len & 31u, wi = tail_at >> 2;   // rem is even
#pragma unroll 1
  for (; rem >= 4u; rem -= 4u) s = crc16_word_sw(s_T2, s, win[wi++]);
  if (rem) s = crc16_half_sw(s_T2, s, win[wi] & 0xffffu);
  return (header_crc_sw(s_T2, 1u, n, payload_len) << 16) | bswap16(s);
}

// payload image (16-byte aligned in shared memory) -> dst (2-byte aligned global address), all NTS threads.
// Body in 16-byte stores.  The source of an aligned destination vector starts `head` bytes into the image (even,
// 0..14): two 16-byte shared loads, five consecutive words picked from the eight (W0 = head / 4), realigned by two
// bytes with PRMT when SK (head & 2).
template <int W0, bool SK>
__device__ __forceinline__ void copy_vectors(uint4 *d4, const uint32_t *s_words, uint32_t nvec, int tid) {
#pragma unroll 1
  for (uint32_t i = tid; i < nvec; i += NTS) {
    const uint4 A = *reinterpret_cast<const uint4 *>(s_words + 4u * i);
    uint4 B = A;
    if (W0 != 0 || SK) B = *reinterpret_cast<const uint4 *>(s_words + 4u * i + 4u);
    const uint32_t x0 = W0 == 0 ? A.x : W0 == 1 ? A.y : W0 == 2 ? A.z : A.w;
    const uint32_t x1 = W0 == 0 ? A.y : W0 == 1 ? A.z : W0 == 2 ? A.w : B.x;
    const uint32_t x2 = W0 == 0 ? A.z : W0 == 1 ? A.w : W0 == 2 ? B.x : B.y;
    const uint32_t x3 = W0 == 0 ? A.w : W0 == 1 ? B.x : W0 == 2 ? B.y : B.z;
    const uint32_t x4 = W0 == 0 ? B.x : W0 == 1 ? B.y : W0 == 2 ? B.z : B.w;
    uint4 v;
    if (SK) {
      v.x = __byte_perm(x0, x1, 0x5432); v.y = __byte_perm(x1, x2, 0x5432);
      v.z = __byte_perm(x2, x3, 0x5432); v.w = __byte_perm(x3, x4, 0x5432);
    } else {
      v.x = x0; v.y = x1; v.z = x2; v.w = x3;
    }
    d4[i] = v;
  }
}
__device__ __forceinline__ void copy_window_out(unsigned char *dst, const uint32_t *s_words, uint32_t L, int tid) {
  uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);  // bytes until dst is 16-byte aligned (even)
  const uint32_t sel = head >> 1;
  if (head > L) head = L;
  const uint32_t nvec = (L - head) >> 4;
  const uint32_t tail0 = head + (nvec << 4);
  const uint16_t *s16 = reinterpret_cast<const uint16_t *>(s_words);
  if ((uint32_t)tid < (head >> 1)) reinterpret_cast<uint16_t *>(dst)[tid] = s16[tid];
  if ((uint32_t)tid >= 32u && (uint32_t)tid - 32u < ((L - tail0) >> 1))
    reinterpret_cast<uint16_t *>(dst + tail0)[tid - 32] = s16[(tail0 >> 1) + (tid - 32)];
  uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
  switch (sel) {
    case 0: copy_vectors<0, false>(d4, s_words, nvec, tid); break;
    case 1: copy_vectors<0, true>(d4, s_words, nvec, tid); break;
    case 2: copy_vectors<1, false>(d4, s_words, nvec, tid); break;
    case 3: copy_vectors<1, true>(d4, s_words, nvec, tid); break;
    case 4: copy_vectors<2, false>(d4, s_words, nvec, tid); break;
    case 5: copy_vectors<2, true>(d4, s_words, nvec, tid); break;
    case 6: copy_vectors<3, false>(d4, s_words, nvec, tid); break;
    default: copy_vectors<3, true>(d4, s_words, nvec, tid); break;
  }
}

// header of a frame at out + goff: ten big-endian halfwords written by lanes 0..9 of one warp
// (id = 1 for audio frames, encoder.rs:210; time = 0, :148-150)
__device__ __forceinline__ void write_header(unsigned char *out, unsigned long long goff, uint32_t n, uint32_t payload_len,
                                             uint32_t hw, int lane) {
  if (lane < 10) {
    uint32_t v = 0;
    if (lane == 0) v = kFrameKey;
    else if (lane == 1) v = 0x0101u;
    else if (lane == 2) v = n & 0xffffu;
    else if (lane == 3) v = payload_len & 0xffffu;
    else if (lane == 8) v = hw >> 16;
    else if (lane == 9) v = hw & 0xffffu;
    reinterpret_cast<uint16_t *>(out + goff)[lane] = (uint16_t)(((v & 0xff) << 8) | (v >> 8));
  }
}

// stream offset of frame f: wait for the scanner to turn the published size into a prefix (one thread); the offset
// goes to slot[0..1], "fits the output buffer" to slot[2]
__device__ __noinline__ void wait_offset(const EncodeArgs &a, uint32_t f, uint32_t frame_bytes, uint32_t *slot,
                                         unsigned long long pv = 0) {
  while ((pv >> 62) != 2ull) {
    pv = ld_status(a.status + f);
    if ((pv >> 62) != 2ull) __nanosleep(32);
  }
  const unsigned long long excl = (pv & kValueMask) - frame_bytes;
  const bool fits = excl + frame_bytes <= a.out_cap;
  if (!fits) atomicMax(a.result + 1, 1ull);            // ByteWriterInsufficientMemory, bytewriter.rs:88
  slot[0] = (uint32_t)excl;
  slot[1] = (uint32_t)(excl >> 32);
  slot[2] = fits ? 1u : 0u;
}

struct StripShared {
  uint32_t *rows, *win, *next, *V, *misc;
  const uint16_t *T2, *N;
};
// s_misc: [0..4) warp totals, [4 + p] bits of the frame, [6 + p] this CTA's next frame (p = iteration parity: values
//         read after the window barrier of one iteration may be rewritten early in the next), [8] next frame (slow path),
//         [10..13) / [13..16) offset and fits of the pending / of the slow frame (slow path), [16..22) stats,
//         [24 + q] header CRC | payload CRC of the frame in window q, [32 + 4p ..) offset and fits of the pending frame,
//         [48..58) mbarriers: "window flushed", three "warp total ready", "slices summed"; [64..192) bit counts of the strips
constexpr int kMiscT = 64, kMiscWords = 192;

// The frame waiting in a window: every thread keeps its description in registers (the values are uniform); only the
// two CRCs, which warp 0 produces late, travel through shared memory (misc[24 + q]).
struct PendingFrame {
  uint32_t q;            // window 0 or 1; 2 = nothing pending; 3 = both windows as one buffer (a payload of 8 .. 16 KiB)
  uint32_t f, n, len;    // frame index, samples, payload bytes
};
constexpr uint32_t kBothBytes = 2u * kWinBytes;   // the two windows and the slack between them are contiguous
__device__ __forceinline__ uint32_t *window_of(const StripShared &S, uint32_t q) { return q == 3u ? S.win : S.win + q * kWinStride; }
// the pending frame goes to the stream; its offset is in misc[10..13).  All threads.
__device__ __forceinline__ void flush_pending(const EncodeArgs &a, const StripShared &S, const PendingFrame &p, int tid,
                                              const uint32_t *slot, int hdr_warp = 1) {
  if (slot[2]) {
    const unsigned long long goff = (unsigned long long)slot[0] | ((unsigned long long)slot[1] << 32);
    copy_window_out(a.out + goff + kFrameHeaderLen, window_of(S, p.q), p.len, tid);
    if ((tid >> 5) == hdr_warp) write_header(a.out, goff, p.n, p.len, S.misc[24 + p.q], tid & 31);
  }
}

// A frame that is not regular: merging relocation (strip_relocate) through window 0 in as many rounds as the payload
// needs, each round summed and written to the stream at once -- which needs the frame's own offset, so the CTA waits
// for it here (after the pending frame is out).  Returns this CTA's next frame (the ticket is drawn after the wait:
// a frame that is ticketed but cannot start would hold up every later frame's offset).
__device__ __noinline__ uint32_t slow_frame(const EncodeArgs &a, const StripShared &S, uint32_t f, uint32_t n, uint32_t T,
                                            uint32_t O, uint32_t total_bits, const PendingFrame &pend, uint32_t spf, uint32_t last_f,
                                            uint32_t last_n) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t payload_len = payload_bytes(total_bits), nch = payload_len >> 5;
  const uint32_t *row = S.rows + (uint32_t)tid * kRowWords;
  uint32_t *misc = S.misc;
  const uint16_t *NA = S.N + 64 * (lane & 3), *NB = S.N + 64 * ((lane >> 2) ? 3 + (lane >> 2) : 0);
  if (tid == 32) {
    if (pend.q != 2u) wait_offset(a, pend.f, (uint32_t)kFrameHeaderLen + pend.len, misc + 10);
    wait_offset(a, f, (uint32_t)kFrameHeaderLen + payload_len, misc + 13);
    misc[8] = atomicAdd(a.ticket, 1u);
  }
  __syncthreads();                                   // every warp is here: the previous frame's finisher is done with S.V
  if (tid < (int)kMaxSlicesStrip) S.V[tid] = 0u;
  if (pend.q != 2u) flush_pending(a, S, pend, tid, misc + 10);
  const unsigned long long goff = (unsigned long long)misc[13] | ((unsigned long long)misc[14] << 32);
  const bool fits = misc[15] != 0u;
  const uint32_t f_next = misc[8];
  uint32_t *win = S.win;
  __syncthreads();                                   // the pending frame may have been in window 0
  const uint32_t nrounds = (payload_len + kWinBytes - 1u) / kWinBytes;
  for (uint32_t r = 0; r < nrounds; r++) {
    const int32_t wbit0 = (int32_t)(8u * r * kWinBytes);
    if (tid == 0) {
      const int32_t zt = (int32_t)total_bits - wbit0;
      if (zt >= 0 && (zt >> 5) < (int32_t)kWinWords) {
        win[zt >> 5] = 0u;                           // completed by nobody: tails OR into it
        win[(zt >> 5) + 1] = 0u;
      }
    }
    uint32_t tail;
    int32_t tail_idx;
    strip_relocate(row, T, (int32_t)O - wbit0, win, kWinWords, tail, tail_idx);
    __syncthreads();
    if (tail_idx >= 0) atomicOr(&win[tail_idx], tail);
    if (r == nrounds - 1u && f_next < a.n_frames)
      stage_rows(a.pcm + (unsigned long long)f_next * spf, f_next == last_f ? last_n : spf, S.rows, S.next, wid, lane);
    __syncthreads();
    const uint32_t vb1 = payload_len - r * kWinBytes < kWinBytes ? payload_len - r * kWinBytes : kWinBytes;
    const uint32_t c_lo = r * kWinChunks, c_hi = nch < (r + 1u) * kWinChunks ? nch : (r + 1u) * kWinChunks;
    crc_slices(win, c_lo, c_hi, nch, S.V, S.T2, S.N, NA, NB, wid, lane);
    if (fits) copy_window_out(a.out + goff + kFrameHeaderLen + (size_t)r * kWinBytes, win, vb1, tid);
    __syncthreads();                                 // window reused by the next round; slice sums complete
  }
  if (wid == 0) {
    uint32_t hw = 0;
    if (lane == 0) hw = crc_finish(S.V, win, 32u * nch - (nrounds - 1u) * kWinBytes, payload_len, n, S.T2, S.N);
    hw = __shfl_sync(0xffffffffu, hw, 0);
    if (fits) write_header(a.out, goff, n, payload_len, hw, lane);
  }
  __syncthreads();                                   // the slice sums and window 0 are free again
  return f_next;
}

__global__ void __launch_bounds__(NTS, 5) encode_frames_strip_kernel(const __grid_constant__ EncodeArgs a) {
  if (a.choice && *a.choice != (unsigned)kEncKernelStrip) return;  // the probe picked the block-per-thread kernel
  if (blockIdx.x == 0) {
    scanner_role(a);
    return;
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StripShared S;
  S.rows = reinterpret_cast<uint32_t *>(smem_raw);
  S.win = reinterpret_cast<uint32_t *>(smem_raw + kRowsBytes);                                // 2 windows
  uint16_t *s_T2 = reinterpret_cast<uint16_t *>(S.win + 2u * kWinStride);
  S.T2 = s_T2;
  S.next = reinterpret_cast<uint32_t *>(s_T2 + 1024);                                         // 4 x 16 bytes
  S.V = S.next + 16;                                                                          // 2 x kMaxSlicesStrip
  S.misc = S.V + 2 * kMaxSlicesStrip;
  uint32_t *s_misc = S.misc;
  uint16_t *s_N = reinterpret_cast<uint16_t *>(s_misc + kMiscWords);                                 // kCrcMulEntries
  S.N = s_N;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint16_t *Tg2 = a.crc_tables + kCrcTableEntries;     // byte-swapped bank (global)
  for (int i = tid; i < 1024; i += NTS) s_T2[i] = Tg2[i];
  for (int i = tid; i < kCrcMulEntries; i += NTS) s_N[i] = a.crc_tables[kCrcBankEntries2 + i];
  // lane l of a slice scales its chunk's sum by x^(256 l): constants (l & 3) and 3 + (l >> 2) of the nibble bank
  const uint16_t *NA = s_N + 64 * (lane & 3), *NB = s_N + 64 * ((lane >> 2) ? 3 + (lane >> 2) : 0);
  if (tid < 6) s_misc[16 + tid] = 0;
  const uint32_t mb_flush = (uint32_t)__cvta_generic_to_shared(s_misc + 48);   // one arrival per warp and iteration
  const uint32_t mb_tot = mb_flush + 8u, mb_crc = mb_flush + 32u;   // s_misc[50..56): three "total ready", [56..58): "slices summed"
  if (tid == 0) {
    s_misc[8] = atomicAdd(a.ticket, 1u);
    mbar_init(mb_flush, 4);
    for (uint32_t k = 0; k < 3u; k++) mbar_init(mb_tot + 8u * k, 1);
    mbar_init(mb_crc, 4);
  }
  __syncthreads();
  const uint32_t spf = a.P.spf;
  const uint32_t last_f = a.n_frames - 1u;
  const uint32_t last_n = (uint32_t)(a.n_samples - (unsigned long long)last_f * spf);
  uint32_t f = s_misc[8];
  if (f < a.n_frames) stage_rows(a.pcm + (unsigned long long)f * spf, f == last_f ? last_n : spf, S.rows, S.next, wid, lane);
  unsigned long long stat_acc = 0;
  uint32_t it = 0, par = 0, crc_phase = 0;
  PendingFrame pend;                                   // the frame not yet written out
  pend.q = 2u; pend.f = 0; pend.n = 0; pend.len = 0;
  uint32_t *row = S.rows + (uint32_t)tid * kRowWords;

  while (f < a.n_frames) {
    const uint32_t n = f == last_f ? last_n : spf;
    const uint32_t nblk = n > 1 ? (n - 2u) / 20u + 1u : 1u;
    const uint32_t nstrips = (nblk + kStripBlocks - 1u) / kStripBlocks;
    // the pending frame's look-back word: asked for now, looked at after the scan (an L2 round trip otherwise sits
    // between the scan and the window barrier)
    // roles rotate over the warps (warp w of every CTA runs on sub-partition w: a fixed finisher would make one
    // scheduler the critical path of all frames)
    const int fin = (int)(it & 3u), pol = (int)((it + 1u) & 3u);
    unsigned long long pend_status = 0;
    if (wid == pol && lane == 0 && pend.q != 2u) pend_status = ld_status(a.status + pend.f);
    cp_async_wait_all();
    __syncwarp();                                      // this warp's rows are staged (nobody else touches them)

    // ---- local pack ----
    uint32_t T = 0;
    {
      const uint32_t nxt = lane == 31 ? S.next[4 * wid] : row[kRowWords + kRowPadWords];
      __syncwarp();                                    // every lane has its look-ahead word before any row changes
      const uint32_t b0 = kStripBlocks * (uint32_t)tid;
      if (b0 < nblk) {
        if (b0 + kStripBlocks <= nblk && n >= kStripSamples * ((uint32_t)tid + 1u)) {
          const bool full = n > kStripSamples * ((uint32_t)tid + 1u);
          uint32_t s19 = 0;
          T = strip_pack_fast(row, nxt, full, tid == 0, a.neg_one, stat_acc, s19);
          if (!full) atomicAdd(&s_misc[16 + s19], 19u);
        } else {
          uint32_t ss[6] = {0, 0, 0, 0, 0, 0};
          T = strip_pack_generic(row, nxt, (uint32_t)tid, n, nblk, a.P, ss);
#pragma unroll
          for (int m = 0; m < 6; m++)
            if (ss[m]) atomicAdd(&s_misc[16 + m], ss[m]);
        }
      }
    }

    // ---- scan of the strips' bit counts: within the warp by shuffles, across warps by a CHAIN of mbarriers --
    // warp w signals "my total is there" and waits only for the warps before it, so an early warp relocates while a
    // late one is still packing (a CTA-wide barrier here cost 15 % of the kernel: the four warps sit on four
    // different schedulers and rarely finish together) ----
    uint32_t incl = T;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_misc[wid] = incl;
    s_misc[kMiscT + tid] = T;
    uint32_t *s_V = S.V + (it & 1u) * kMaxSlicesStrip;   // by iteration: the finisher of the previous frame may still read the other one
    if (tid >= 32 && tid < 32 + (int)kMaxSlicesStrip) s_V[tid - 32] = 0u;
    // "warp w's total and bit counts are there": one mbarrier per warp 0..2, one arrival (lane 0, after the warp has
    // synchronised), waited for by the warps behind it
    uint32_t wbase = 0;
#ifdef X3_RACECHECK_BARRIERS
    // sanitizer build (tools/sanitize_run.sh): compute-sanitizer's racecheck does not follow mbarrier arrive -> wait
    // ordering (tools/ubench/mbar_racecheck.cu), so this build puts a CTA barrier wherever the product waits on an
    // mbarrier -- a superset of the product's ordering that racecheck can see; everything else is unchanged
    __syncthreads();
    for (int k = 0; k < wid; k++) wbase += s_misc[k];
#else
    __syncwarp();
    if (wid < 3 && lane == 0) mbar_arrive(mb_tot + 8u * (uint32_t)wid);
    for (int k = 0; k < wid; k++) {
      mbar_wait(mb_tot + 8u * (uint32_t)k, it & 1u);
      wbase += s_misc[k];
    }
#endif
    const uint32_t O = wbase + incl - T;
    uint32_t *win = S.win + par * kWinStride;
    // The last warp knows the frame's size first: it publishes it (nobody waits for the offset here), decides whether
    // the frame is regular, and draws the CTA's next frame -- the ticket is consumed after the window barrier, so the
    // atomic's round trip hides behind the relocation.  (A strip that is not the frame's last always holds four whole
    // blocks, at least 88 bits: "regular" is only a question of size.)
    uint32_t ticket = 0;
    if (tid == NTS - 1) {
      const uint32_t tb = wbase + incl;
      const uint32_t pl = payload_bytes(tb);
      st_status(a.status + f, kFlagAgg | (unsigned long long)((uint32_t)kFrameHeaderLen + pl));
      s_misc[4 + (it & 1u)] = tb;
      if (pl <= kBothBytes) ticket = atomicAdd(a.ticket, 1u);
    }
    // the pending frame's offset (published a frame time ago): one thread asks, everybody knows after (B4)
    uint32_t *off_slot = s_misc + 32 + 4 * (it & 1u);
    if (wid == pol && lane == 0 && pend.q != 2u)
      wait_offset(a, pend.f, (uint32_t)kFrameHeaderLen + pend.len, off_slot, pend_status);
    // The window about to be written is the one the previous iteration copied out at its end: every warp has said
    // "my share is out" on an mbarrier since (a split barrier: the arrival was a whole pack phase ago, so nobody waits).
#ifdef X3_RACECHECK_BARRIERS
    __syncthreads();
#else
    if (it) mbar_wait(mb_flush, (it - 1u) & 1u);
#endif
    // relocation, before the frame is known to fit the window: a strip that would leave it stays where it is (the
    // frame then takes the slow path, which starts over from the rows)
    const bool optimistic = pend.q != 3u;              // a pending frame that fills both windows must go out first
    if (optimistic && (uint32_t)tid < nstrips && ((O + T + 31u) >> 5) <= kWinWords)
      strip_relocate_fast(row, T, O, tid ? s_misc[kMiscT + tid - 1] : 32u, (uint32_t)tid + 1u == nstrips, win);
    if (tid == NTS - 1) s_misc[6 + (it & 1u)] = ticket;
    // (Letting the warps that are done early copy their share of the pending frame out here, before the barrier, was
    // measured: 1.56 -> 1.68 ms.)
    __syncthreads();                                   // (B4) window complete; size, ticket and pending offset visible
    const uint32_t total_bits = s_misc[4 + (it & 1u)];
    const uint32_t payload_len = payload_bytes(total_bits);
    const uint32_t nch = payload_len >> 5;             // whole 32-byte CRC chunks

    if (payload_len <= kWinBytes && optimistic) {
      const uint32_t f_next = s_misc[6 + (it & 1u)];
      if (f_next < a.n_frames)
        stage_rows(a.pcm + (unsigned long long)f_next * spf, f_next == last_f ? last_n : spf, S.rows, S.next, wid, lane);
      crc_slices(win, 0u, nch, nch, s_V, s_T2, s_N, NA, NB, wid, lane);
      // (B5) "this warp's slices are summed": every warp arrives, only the finisher waits
#ifdef X3_RACECHECK_BARRIERS
      __syncthreads();
      if (wid == fin) {
#else
      __syncwarp();
      if (lane == 0) mbar_arrive(mb_crc);
      if (wid == fin) {
        mbar_wait(mb_crc, crc_phase & 1u);
#endif
        if (lane == 0) s_misc[24 + par] = crc_finish(s_V, win, 32u * nch, payload_len, n, s_T2, s_N);
      }
      crc_phase++;
      if (pend.q != 2u) flush_pending(a, S, pend, tid, off_slot, pol);
      pend.q = par; pend.f = f; pend.n = n; pend.len = payload_len;
      par ^= 1u;
      f = f_next;
    } else if (payload_len <= kBothBytes) {
      // A payload of 8 .. 16 KiB (BFP / literal heavy frames), or any frame behind one: the pending frame goes out
      // FIRST (its offset has had a frame time to arrive), then the strips are relocated into the window, or into
      // both windows taken as one buffer.  Two more CTA barriers than the common path, but still no wait for the
      // frame's own offset.
      if (pend.q != 2u) flush_pending(a, S, pend, tid, off_slot, pol);
      __syncthreads();
      const uint32_t q = payload_len <= kWinBytes ? par : 3u;
      uint32_t *wq = window_of(S, q);
      if ((uint32_t)tid < nstrips)
        strip_relocate_fast(row, T, O, tid ? s_misc[kMiscT + tid - 1] : 32u, (uint32_t)tid + 1u == nstrips, wq);
      __syncthreads();
      const uint32_t f_next = s_misc[6 + (it & 1u)];
      if (f_next < a.n_frames)
        stage_rows(a.pcm + (unsigned long long)f_next * spf, f_next == last_f ? last_n : spf, S.rows, S.next, wid, lane);
      crc_slices(wq, 0u, nch, nch, s_V, s_T2, s_N, NA, NB, wid, lane);
#ifdef X3_RACECHECK_BARRIERS
      __syncthreads();
      if (wid == fin) {
#else
      __syncwarp();
      if (lane == 0) mbar_arrive(mb_crc);
      if (wid == fin) {
        mbar_wait(mb_crc, crc_phase & 1u);
#endif
        if (lane == 0) s_misc[24 + q] = crc_finish(s_V, wq, 32u * nch, payload_len, n, s_T2, s_N);
      }
      crc_phase++;
      pend.q = q; pend.f = f; pend.n = n; pend.len = payload_len;
      if (q != 3u) par ^= 1u;
      f = f_next;
    } else {
      f = slow_frame(a, S, f, n, T, O, total_bits, pend, spf, last_f, last_n);
      pend.q = 2u;
    }
#ifndef X3_RACECHECK_BARRIERS
    __syncwarp();
    if (lane == 0) mbar_arrive(mb_flush);              // this warp is done with the windows of this iteration
#endif
    it++;
    if ((it & 127u) == 0u) {                           // the 10-bit counters (4 blocks per frame) are about to fill up
#pragma unroll
      for (int m = 0; m < 6; m++) {
        const uint32_t c = (uint32_t)(stat_acc >> (10 * m)) & 1023u;
        if (c) atomicAdd(&s_misc[16 + m], c * 20u);
      }
      stat_acc = 0;
    }
  }
  // ---- drain: the last frame of this CTA is still in its window ----
  __syncthreads();
  if (pend.q != 2u) {
    if (tid == 32) wait_offset(a, pend.f, (uint32_t)kFrameHeaderLen + pend.len, s_misc + 10);
    __syncthreads();
    flush_pending(a, S, pend, tid, s_misc + 10);
  }
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const uint32_t c = (uint32_t)(stat_acc >> (10 * m)) & 1023u;
    if (c) atomicAdd(&s_misc[16 + m], c * 20u);
  }
  __syncthreads();
  if (tid < 6 && s_misc[16 + tid]) atomicAdd(a.result + 2 + tid, (unsigned long long)s_misc[16 + tid]);
}


// Workspace zeroing + kernel choice (x3_kernels.h).  CTA 0 zeroes the first 128 bytes (results, ticket, choice), codes
// the sampled blocks and writes the choice; the other CTAs zero the rest of the workspace.
__global__ void __launch_bounds__(256) encode_probe_kernel(const int16_t *pcm, unsigned long long n_samples, uint4 *ws,
                                                           unsigned long long ws_vecs, unsigned int *choice) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  if (blockIdx.x != 0) {
    for (unsigned long long i = 8ull + (unsigned long long)(blockIdx.x - 1u) * 256u + threadIdx.x; i < ws_vecs;
         i += (unsigned long long)(gridDim.x - 1u) * 256u)
      ws[i] = z;
    return;
  }
  __shared__ unsigned int s_big, s_all;
  if (threadIdx.x < 8) ws[threadIdx.x] = z;
  if (threadIdx.x == 0) { s_big = 0; s_all = 0; }
  if (gridDim.x == 1)
    for (unsigned long long i = 8ull + threadIdx.x; i < ws_vecs; i += 256u) ws[i] = z;
  __syncthreads();
  const unsigned long long nblocks = n_samples > 1 ? (n_samples - 1) / 20ull : 0ull;   // 20 differences each
  const unsigned long long take = nblocks < 256ull ? nblocks : 256ull;                 // one block per thread
  unsigned int big = 0, all = 0;
  if (threadIdx.x < take) {
    const int16_t *q = pcm + (unsigned long long)threadIdx.x * (nblocks / take) * 20ull;
    int v[21];
#pragma unroll
    for (int i = 0; i <= 20; i++) v[i] = q[i];     // 21 independent loads: one memory latency
    int m = 0;
#pragma unroll
    for (int i = 1; i <= 20; i++) {
      const int d = v[i] - v[i - 1], ad = d < 0 ? -d : d;
      m = ad > m ? ad : m;
    }
    all = 1;
    big = m > 20;   // Parameters::default(): a block whose largest difference exceeds thresholds[2] is BFP / literal
  }
  atomicAdd(&s_big, big);
  atomicAdd(&s_all, all);
  __syncthreads();
  if (threadIdx.x == 0)
    *choice = (s_all && 100u * s_big > kProbeBigPercent * s_all) ? (unsigned)kEncKernelFast : (unsigned)kEncKernelStrip;
}

}  // namespace

size_t encode_smem_bytes(const CodecParams &P, uint32_t max_blocks, uint32_t out_words_cap) {
  const uint32_t in_bytes = (2u * (P.spf + 8u) + 15u) & ~15u;
  const uint32_t img_bytes = (32u + 4u * (out_words_cap + 8u) + 15u) & ~15u;
  const uint32_t nblk_cap = max_blocks + 2u;
  const uint32_t arr = ((nblk_cap * 4u) + 15u) & ~15u;
  const uint32_t chunk = (((out_words_cap / 4u + 2u) * 2u) + 15u) & ~15u;
  return (size_t)in_bytes + img_bytes + kCrcTableEntries * 2 + 3u * arr + chunk + 64u * 4u;
}

size_t encode_fast_smem_bytes(const CodecParams &P, uint32_t out_words_cap) {
  const uint32_t in_bytes = (2u * (P.spf + 8u) + 15u) & ~15u;
  const uint32_t img_bytes = (4u * (out_words_cap + 8u) + 15u) & ~15u;
  return (size_t)in_bytes + NB * img_bytes + kCrcBankEntries2 * 2 + NB * kMaxSlices * 4u + 96u * 4u;
}

size_t encode_strip_smem_bytes() {
  return (size_t)kRowsBytes + 2u * kWinStride * 4u + 1024u * 2u + 16u * 4u + 2u * kMaxSlicesStrip * 4u + (size_t)kMiscWords * 4u + (size_t)kCrcMulEntries * 2u;
}

cudaError_t launch_encode(const EncodeArgs &a, int kind, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e;
  if (kind == kEncKernelStrip) {
    e = cudaFuncSetAttribute(encode_frames_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    encode_frames_strip_kernel<<<grid, NTS, smem, stream>>>(a);
  } else if (kind == kEncKernelFast) {
    const bool std_frame = a.P.spf == kStdSpf && a.out_words_cap == kStdOwc;
    auto kern = std_frame ? encode_frames_fast_kernel<kStdSpf, kStdOwc> : encode_frames_fast_kernel<0, 0>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NTF, smem, stream>>>(a);
  } else {
    e = cudaFuncSetAttribute(encode_frames_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    encode_frames_generic_kernel<<<grid, NT, smem, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_encode_probe(const int16_t *pcm, unsigned long long n_samples, unsigned char *ws, size_t ws_bytes,
                                unsigned int *choice, cudaStream_t stream) {
  const unsigned long long vecs = ws_bytes / 16;   // the workspace size is a multiple of 16
  unsigned grid = 1u + (unsigned)((vecs + 256ull * 16ull - 1ull) / (256ull * 16ull));
  if (grid > 149u) grid = 149u;
  encode_probe_kernel<<<grid, 256, 0, stream>>>(pcm, n_samples, reinterpret_cast<uint4 *>(ws), vecs, choice);
  return cudaGetLastError();
}

int encode_occupancy(int kind, size_t smem) {
  int nb = 0;
  cudaError_t e;
  if (kind == kEncKernelStrip) {
    cudaFuncSetAttribute(encode_frames_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, encode_frames_strip_kernel, NTS, smem);
  } else if (kind == kEncKernelFast) {
    cudaFuncSetAttribute(encode_frames_fast_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, encode_frames_fast_kernel<0, 0>, NTF, smem);
  } else {
    cudaFuncSetAttribute(encode_frames_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, encode_frames_generic_kernel, NT, smem);
  }
  if (e != cudaSuccess || nb < 1) nb = 1;
  return nb;
}

}  // namespace x3
