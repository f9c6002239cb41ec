// x3_decode.cu -- frame index, payload CRC and frame decode kernels for sm_100a.
//
// Decode of a device-resident frame stream is stream-ordered, no host round trip before the end:
//   1. hop_index_kernel     -- one warp per 64 KiB tile finds the tile's first frame header (key 'x3' + header CRC +
//      field checks of decoder.rs:69-118) and follows header -> payload_len -> next header to the tile's end;
//      tile_prefix_kernel + place_frames_kernel turn the tiles' counts into frame ordinals / sample offsets and
//      write the dense frame table; check_chain_kernel then proves the table is exactly the header ->
//      payload_len -> next header walk of decodefile.rs:105-126.  Anything it cannot prove (false candidate,
//      corrupt header, odd payload length) raises a flag and the API falls back to the sequential host walk,
//      which reproduces the reference's error behaviour.  scan_headers_kernel (every even offset tested) indexes
//      streams of frames too small for the hop index's 64 frames per tile.
//   2. crc_frames_kernel    -- one thread per frame, word folding (crc16_fold: three xors per 32 bits), payload
//      streamed through a per-lane cp.async ring: decodefile.rs:93-103.  Runs on a second stream beside step 3.
//   3. decode_frames_kernel -- one thread per frame (frames are independent, blocks inside a frame are
//      not): x3_dec_core.cuh.  Every lane streams its own payload through a private 144-byte
//      shared-memory ring filled by cp.async one block ahead, and writes whole 32-byte sectors (one 256-bit store).
#include <cuda_runtime.h>

#include <cstdlib>

#include "x3_dec_core.cuh"
#include "x3_kernels.h"
#include "x3_lookback.cuh"

namespace x3 {

namespace {

// ------------------------------------------------------------------------------------------------
// 1. frame index
// ------------------------------------------------------------------------------------------------
constexpr int kScanCap = 1024;                // candidates per tile (128 KiB, or 16 KiB on the retry) before falling back
constexpr int kHopCap = 64;                   // frames per tile the hop index keeps (two per lane)
constexpr int kCountShift = 38;               // look-back value = (frames << 38) | samples

struct Cand {
  uint32_t off;       // byte offset inside the tile
  uint32_t samples;
  uint32_t payload_len;
  uint32_t payload_crc;
};

// Stream-ordered API: the stream's length lies on the device (ScanArgs::len_dev); the length the host passed is only an
// upper bound, and so is the number of tiles derived from it.
__device__ __forceinline__ void apply_device_len(ScanArgs &a) {
  if (a.len_dev) {
    const unsigned long long v = *a.len_dev;
    a.stream_len = v < a.stream_len ? v : a.stream_len;
    const unsigned long long t = (a.stream_len + a.tile_bytes - 1ull) / a.tile_bytes;
    a.n_tiles = t < a.n_tiles ? (uint32_t)t : a.n_tiles;
  }
}

// decoder::read_frame_header checks (decoder.rs:69-118) on 20 bytes at p (p is 2-byte aligned)
__device__ bool header_valid(const uint8_t *p, const uint16_t *T, Cand &c) {
  const uint16_t *h = reinterpret_cast<const uint16_t *>(p);
  uint32_t w[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const uint32_t v = h[i];
    w[i] = ((v & 0xff) << 8) | (v >> 8);  // big-endian halfword
  }
  uint32_t s = 0xffffu;
#pragma unroll
  for (int i = 0; i < 8; i++) s = crc16_half(T, s, w[i]);
  if (s != w[8]) return false;                    // FrameHeaderInvalidHeaderCRC
  if (w[0] != kFrameKey) return false;            // FrameHeaderInvalidKey
  if ((w[1] & 0xff) > 1u) return false;           // MoreThanOneChannel
  if (w[3] >= kFrameMaxLength) return false;      // FrameLength
  c.samples = w[2];
  c.payload_len = w[3];
  c.payload_crc = w[9];
  return true;
}

// nonzero iff one of the two halfwords of w is the frame key (bytes 78 33 in memory order); may report a false
// positive (borrow out of a zero low halfword), which examine_piece sorts out
__device__ __forceinline__ uint32_t haskey(uint32_t w) {
  const uint32_t x = w ^ 0x33783378u;
  return (x - 0x00010001u) & ~x & 0x80008000u;
}

// rare path: a 16-byte piece that contains the key somewhere
__device__ __noinline__ void examine_piece(const ScanArgs &a, const uint16_t *s_T, Cand *s_cand, unsigned int *s_count,
                                           unsigned long long t0, unsigned long long o, uint4 v) {
  const uint32_t xs[4] = {v.x ^ 0x33783378u, v.y ^ 0x33783378u, v.z ^ 0x33783378u, v.w ^ 0x33783378u};
  for (int j = 0; j < 4; j++) {
    for (int hsel = 0; hsel < 2; hsel++) {
      if (((xs[j] >> (16 * hsel)) & 0xffffu) != 0u) continue;
      const unsigned long long p = o + 4u * j + 2u * hsel;
      if (p + kFrameHeaderLen > a.stream_len) continue;
      Cand c;
      if (!header_valid(a.stream + p, s_T, c)) continue;
      c.off = (uint32_t)(p - t0);
      const unsigned int idx = atomicAdd(s_count, 1u);
      if (idx < kScanCap) s_cand[idx] = c;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_headers_kernel(const ScanArgs a_in) {
  ScanArgs a = a_in;
  apply_device_len(a);
  __shared__ uint16_t s_T[512];  // T_0, T_1 of the CRC bank are enough for halfword updates
  __shared__ Cand s_cand[kScanCap];
  __shared__ uint32_t s_rank[kScanCap];
  __shared__ unsigned int s_count;
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x;
  for (int i = tid; i < 512; i += kScanThreads) s_T[i] = a.crc_tables[i];

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_tile = atomicAdd(a.ticket, 1u); s_count = 0; }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= a.n_tiles) break;
    const unsigned long long t0 = (unsigned long long)tile * a.tile_bytes;
    const unsigned long long t1 = t0 >= a.stream_len ? t0 : (t0 + a.tile_bytes < a.stream_len ? t0 + a.tile_bytes : a.stream_len);

    // ---- candidates: halfword 'x','3' at an even offset with a valid header behind it ----
    // Whole 16-byte pieces, four independent loads in flight per thread; key hits are rare and handled out of line.
    {
      const unsigned long long t1v = t0 + ((t1 - t0) & ~15ull);  // end of the whole 16-byte pieces of this tile
      const unsigned long long step = (unsigned long long)kScanThreads * 16u;
      unsigned long long o = t0 + (unsigned long long)tid * 16u;
#define X3_KEYTEST(v) (haskey((v).x) | haskey((v).y) | haskey((v).z) | haskey((v).w))
      for (; o + 3u * step < t1v; o += 4u * step) {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(a.stream + o);
        const uint4 v1 = *reinterpret_cast<const uint4 *>(a.stream + o + step);
        const uint4 v2 = *reinterpret_cast<const uint4 *>(a.stream + o + 2u * step);
        const uint4 v3 = *reinterpret_cast<const uint4 *>(a.stream + o + 3u * step);
        if (X3_KEYTEST(v0)) examine_piece(a, s_T, s_cand, &s_count, t0, o, v0);
        if (X3_KEYTEST(v1)) examine_piece(a, s_T, s_cand, &s_count, t0, o + step, v1);
        if (X3_KEYTEST(v2)) examine_piece(a, s_T, s_cand, &s_count, t0, o + 2u * step, v2);
        if (X3_KEYTEST(v3)) examine_piece(a, s_T, s_cand, &s_count, t0, o + 3u * step, v3);
      }
      for (; o < t1v; o += step) {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(a.stream + o);
        if (X3_KEYTEST(v0)) examine_piece(a, s_T, s_cand, &s_count, t0, o, v0);
      }
#undef X3_KEYTEST
      if (tid == 0 && t1v < t1) {  // fewer than 16 bytes left: no header fits, nothing to find
      }
    }
    __syncthreads();
    const uint32_t cnt = s_count < (unsigned)kScanCap ? s_count : (unsigned)kScanCap;
    if (tid == 0 && s_count > (unsigned)kScanCap) atomicOr(a.result + 2, 3ull);   // bit 1: a capacity, not the stream

    // ---- order inside the tile (rank sort; a tile holds about a dozen frames) ----
    unsigned long long tile_samples = 0;
    for (uint32_t j = tid; j < cnt; j += kScanThreads) {
      uint32_t r = 0;
      const uint32_t mine = s_cand[j].off;
      for (uint32_t m = 0; m < cnt; m++) r += s_cand[m].off < mine;
      s_rank[r] = j;
    }
    __syncthreads();

    // ---- hand the tile's frames over, in order, with sample offsets relative to the tile; the tile's place in the
    // stream (frame ordinal and sample offset of its first frame) is worked out afterwards by tile_prefix_kernel and
    // place_frames_kernel.  (A decoupled look-back here made every CTA wait, once per tile, for all the tiles in
    // flight before it: 40 % of the kernel's time.) ----
    if (tid < 32) {
      for (uint32_t j = tid; j < cnt; j += 32) tile_samples += s_cand[j].samples;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) tile_samples += __shfl_xor_sync(0xffffffffu, tile_samples, d);
      if (tid == 0) {
        const unsigned long long off = atomicAdd(a.rec_cursor, (unsigned long long)cnt);
        s_base = off;
        a.tile_status[tile] = ((unsigned long long)cnt << kCountShift) | tile_samples;
        a.tile_recs[tile] = (off << 11) | cnt;   // cnt <= kScanCap = 1024
      }
    }
    __syncthreads();
    const unsigned long long rec_off = s_base;
    for (uint32_t r = tid; r < cnt; r += kScanThreads) {
      const Cand c = s_cand[s_rank[r]];
      unsigned long long before = 0;
      for (uint32_t m = 0; m < r; m++) before += s_cand[s_rank[m]].samples;
      if (rec_off + r < a.max_frames) {
        FrameRec fr;
        fr.pos = t0 + c.off;
        fr.out_off = before;
        fr.samples = c.samples;
        fr.payload_len = c.payload_len;
        fr.payload_crc = c.payload_crc;
        fr.pad = 0;
        a.recs[rec_off + r] = fr;
      } else {
        atomicOr(a.result + 2, 3ull);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 1b. frame index by hopping: one warp per tile finds the tile's first frame header, then follows
// header -> payload_len -> next header (decodefile.rs:105-126) to the end of the tile and validates the headers it
// visited in parallel, one lane each.  It reads a few KB per tile (up to the first header) plus one or two sectors per
// frame instead of the whole stream -- the scan above reads every byte (0.13 ms on a 655 MB stream).  A tile's chain
// is only a guess until check_chain_kernel has proven that the tiles' chains join up from offset 0, exactly as for
// the scan; a false candidate before a tile's first real header (2^-32 per position) or a corrupt header breaks the
// chain and sends the stream to the host walk.  More than kHopCap frames in a tile raise the capacity flag and the
// stream is indexed by the scan kernel with small tiles instead.
// ------------------------------------------------------------------------------------------------
constexpr int kHopThreads = 256, kHopWarps = kHopThreads / 32;

// lowest valid header position in the 16-byte piece at o (o is 16-byte aligned), or ~0
__device__ __noinline__ unsigned long long first_header_in_piece(const ScanArgs &a, const uint16_t *s_T,
                                                                 unsigned long long o, uint4 v) {
  const uint32_t xs[4] = {v.x ^ 0x33783378u, v.y ^ 0x33783378u, v.z ^ 0x33783378u, v.w ^ 0x33783378u};
  for (int j = 0; j < 4; j++) {
    for (int hsel = 0; hsel < 2; hsel++) {
      if (((xs[j] >> (16 * hsel)) & 0xffffu) != 0u) continue;
      const unsigned long long p = o + 4u * j + 2u * hsel;
      if (p + kFrameHeaderLen > a.stream_len) continue;
      Cand c;
      if (header_valid(a.stream + p, s_T, c)) return p;
    }
  }
  return ~0ull;
}

#ifndef X3_HOP_MINBLOCKS
#define X3_HOP_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(kHopThreads, X3_HOP_MINBLOCKS) hop_index_kernel(const ScanArgs a_in) {
  ScanArgs a = a_in;
  apply_device_len(a);
  __shared__ uint16_t s_T[512];
  __shared__ unsigned long long s_pos[kHopWarps][kHopCap];
  const int tid = threadIdx.x;
  for (int i = tid; i < 512; i += kHopThreads) s_T[i] = a.crc_tables[i];
  __syncthreads();
  const uint32_t lane = tid & 31u, wid = tid >> 5;
  const uint32_t warps = gridDim.x * kHopWarps;
  const unsigned long long none = ~0ull;
  for (uint32_t tile = blockIdx.x * kHopWarps + wid; tile < a.n_tiles; tile += warps) {
    const unsigned long long t0 = (unsigned long long)tile * a.tile_bytes;
    const unsigned long long t1 = t0 >= a.stream_len ? t0 : (t0 + a.tile_bytes < a.stream_len ? t0 + a.tile_bytes : a.stream_len);
    const unsigned long long t1v = t0 + ((t1 - t0) & ~15ull);  // a header does not fit in a shorter tail piece

    // ---- the tile's first header: 2 KiB per step, four 16-byte loads in flight per lane ----
    unsigned long long first = none;
    for (unsigned long long o = t0; o < t1v && first == none; o += 2048u) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned long long po = o + (unsigned long long)(k * 32 + (int)lane) * 16u;
        v[k] = po < t1v ? *reinterpret_cast<const uint4 *>(a.stream + po) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t m = __ballot_sync(0xffffffffu, (haskey(v[k].x) | haskey(v[k].y) | haskey(v[k].z) | haskey(v[k].w)) != 0u);
        while (m != 0u && first == none) {
          const int l = __ffs((int)m) - 1;
          unsigned long long found = none;
          if ((int)lane == l) found = first_header_in_piece(a, s_T, o + (unsigned long long)(k * 32 + l) * 16u, v[k]);
          first = __shfl_sync(0xffffffffu, found, l);
          m &= m - 1u;
        }
      }
    }

    // ---- follow the chain to the end of the tile (every lane walks it; the loads are broadcasts) ----
    uint32_t cnt = 0;
    bool over = false;
    for (unsigned long long p = first; p < t1 && p + kFrameHeaderLen <= a.stream_len;) {
      if (cnt == (uint32_t)kHopCap) { over = true; break; }
      if (lane == 0) s_pos[wid][cnt] = p;
      cnt++;
      const uint32_t v = __ldg(reinterpret_cast<const uint16_t *>(a.stream + p + 6u));  // payload_len, big-endian
      p += (unsigned long long)kFrameHeaderLen + (((v & 0xffu) << 8) | (v >> 8));
    }
    __syncwarp();

    // ---- validate, one lane per header; sample offsets relative to the tile ----
    Cand c[kHopCap / 32];
    unsigned long long before[kHopCap / 32];
    unsigned long long tile_samples = 0;
    bool bad = false;
#pragma unroll
    for (int r = 0; r < kHopCap / 32; r++) {
      const uint32_t j = (uint32_t)r * 32u + lane;
      c[r].samples = 0;
      if (j < cnt) {
        const unsigned long long p = s_pos[wid][j];
        if (!header_valid(a.stream + p, s_T, c[r])) { bad = true; c[r].samples = 0; c[r].payload_len = 0; c[r].payload_crc = 0; }
        c[r].off = (uint32_t)(p - t0);
      }
      unsigned long long incl = c[r].samples;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
      }
      before[r] = tile_samples + incl - c[r].samples;
      tile_samples += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.result + 2, 1ull);
    if (over && lane == 0) atomicOr(a.result + 2, 3ull);

    unsigned long long rec_off = 0;
    if (lane == 0) {
      rec_off = atomicAdd(a.rec_cursor, (unsigned long long)cnt);
      a.tile_status[tile] = ((unsigned long long)cnt << kCountShift) | tile_samples;
      a.tile_recs[tile] = (rec_off << 11) | cnt;
    }
    rec_off = __shfl_sync(0xffffffffu, rec_off, 0);
#pragma unroll
    for (int r = 0; r < kHopCap / 32; r++) {
      const uint32_t j = (uint32_t)r * 32u + lane;
      if (j >= cnt) continue;
      if (rec_off + j < a.max_frames) {
        FrameRec fr;
        fr.pos = t0 + c[r].off;
        fr.out_off = before[r];
        fr.samples = c[r].samples;
        fr.payload_len = c[r].payload_len;
        fr.payload_crc = c[r].payload_crc;
        fr.pad = 0;
        a.recs[rec_off + j] = fr;
      } else {
        atomicOr(a.result + 2, 3ull);
      }
    }
    __syncwarp();
  }
}

// Exclusive prefix over the tiles of (frames << 38 | samples), in place; totals to result[0], result[1].  One CTA.
__global__ void __launch_bounds__(1024) tile_prefix_kernel(const ScanArgs a_in) {
  ScanArgs a = a_in;
  apply_device_len(a);
  __shared__ unsigned long long s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const uint32_t per = (a.n_tiles + 1023u) / 1024u;
  const uint32_t t0 = tid * per, t1 = t0 + per < a.n_tiles ? t0 + per : a.n_tiles;
  unsigned long long sum = 0;
  for (uint32_t t = t0; t < t1; t++) sum += a.tile_status[t];
  unsigned long long incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += o;
  }
  if (lane == 31u) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= (uint32_t)d) wi += o;
    }
    s_warp[lane] = wi - w;  // exclusive over warps
    if (lane == 31u) {
      a.result[0] = wi >> kCountShift;
      a.result[1] = wi & ((1ull << kCountShift) - 1ull);
    }
  }
  __syncthreads();
  unsigned long long run = s_warp[wid] + incl - sum;
  for (uint32_t t = t0; t < t1; t++) {
    const unsigned long long own = a.tile_status[t];
    a.tile_status[t] = run;
    run += own;
  }
}

// One warp per tile: the tile's records go to their final place in the frame table.
__global__ void __launch_bounds__(256) place_frames_kernel(const ScanArgs a_in) {
  ScanArgs a = a_in;
  apply_device_len(a);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < a.n_tiles; t += warps) {
    const unsigned long long tr = a.tile_recs[t], base = a.tile_status[t];
    const uint32_t cnt = (uint32_t)(tr & 2047ull);
    const unsigned long long rec_off = tr >> 11;
    const unsigned long long base_count = base >> kCountShift, base_samples = base & ((1ull << kCountShift) - 1ull);
    for (uint32_t j = lane; j < cnt; j += 32u) {
      if (rec_off + j >= a.max_frames) continue;   // flagged by the scan kernel
      FrameRec fr = a.recs[rec_off + j];
      fr.out_off += base_samples;
      if (base_count + j < a.max_frames) a.frames[base_count + j] = fr;
      else atomicOr(a.result + 2, 3ull);
    }
  }
}

// The table is the reference's walk iff it starts at 0, every frame ends where the next begins, and
// nothing but a short tail (<= 20 bytes, decodefile.rs:107-109) or a truncated final frame
// (decodefile.rs:114-116) follows.
__global__ void check_chain_kernel(const ScanArgs a_in) {
  ScanArgs a = a_in;
  apply_device_len(a);
  const unsigned long long n = a.result[0];
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  if (i == 0) {
    if (n == 0) {
      a.result[4] = 0;
      a.result[5] = 0;
      if (a.stream_len > (unsigned long long)kFrameHeaderLen) atomicOr(a.result + 2, 1ull);
    } else if (n <= a.max_frames && a.frames[0].pos != 0) {
      atomicOr(a.result + 2, 1ull);
    }
    if (n > a.max_frames) atomicOr(a.result + 2, 3ull);
  }
  if (n > a.max_frames) return;
  for (unsigned long long k = i; k < n; k += stride) {
    const FrameRec fr = a.frames[k];
    const unsigned long long end = fr.pos + kFrameHeaderLen + fr.payload_len;
    if (k + 1 < n) {
      if (a.frames[k + 1].pos != end) atomicOr(a.result + 2, 1ull);
    } else {
      // final counts for the kernels that follow (result[4] frames, result[5] samples)
      if (end > a.stream_len) {
        a.result[4] = n - 1;  // truncated final frame: the reference stops cleanly before it
        a.result[5] = a.result[1] - fr.samples;
        a.result[6] = fr.pos;  // bytes the walk consumed
      } else {
        a.result[4] = n;
        a.result[5] = a.result[1];
        a.result[6] = end;
        if (a.stream_len - end > (unsigned long long)kFrameHeaderLen) atomicOr(a.result + 2, 1ull);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2. payload CRC, one warp per frame
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_be32_2aligned(const uint8_t *p) {
  // p is 2-byte aligned
  const uint16_t *h = reinterpret_cast<const uint16_t *>(p);
  const uint32_t a = h[0], b = h[1];
  return ((a & 0xff) << 24) | ((a >> 8) << 16) | ((b & 0xff) << 8) | (b >> 8);
}

#ifndef X3_CRC_FOLD
#define X3_CRC_FOLD 1   // 1: one thread per frame, word folding (crc16_fold); 0: one warp per frame, chunk CRCs + combine tree
#endif
#ifndef X3_CRC_THREADS
#define X3_CRC_THREADS (X3_CRC_FOLD ? 64 : 256)
#endif
#if X3_CRC_FOLD
// One thread per frame: three xors per 32-bit word (crc16_fold, x3_dec_core.cuh) instead of the ~26 shifts and xors of
// the word-at-a-time CRC -- a tenth of the instructions of the warp-per-frame kernel below, which matters because
// this kernel shares the SMs with decode_frames_kernel (both live on the ALU pipe).  With so little arithmetic the
// kernel is bound by how many bytes it keeps in flight: every lane streams its payload through a private ring of
// kCrcRingBlocks 64-byte blocks in shared memory, filled by 16-byte cp.async seven blocks ahead of the fold.
constexpr int kCrcRingBlocks = 8;
constexpr int kCrcRowBytes = kCrcRingBlocks * 64 + 16;   // + 16: rows 4 banks apart, 16-byte reads of 8 lanes never collide
struct CrcRingSource {
  const uint8_t *base;
  unsigned char *row;     // this lane's ring
  uint32_t nb, issued;
  __device__ __forceinline__ void issue_one() {   // one block, one commit group (empty past the end: the count stays uniform)
    if (issued < nb) {
      unsigned char *dst = row + 64u * (issued & (kCrcRingBlocks - 1));
      const uint8_t *src = base + 64ull * issued;
#pragma unroll
      for (int q = 0; q < 4; q++) cp_async16(dst + 16 * q, src + 16 * q);
    }
    cp_async_commit();
    issued++;
  }
  __device__ __forceinline__ void begin(const uint8_t *first_block, uint32_t n_blocks) {
    base = first_block;
    nb = n_blocks;
    issued = 0;
#pragma unroll
    for (int i = 0; i < kCrcRingBlocks - 1; i++) issue_one();
  }
  __device__ __forceinline__ void block(uint32_t b, CrcVec v[4]) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kCrcRingBlocks - 2) : "memory");   // block b has landed
    const uint4 *p = reinterpret_cast<const uint4 *>(row + 64u * (b & (kCrcRingBlocks - 1)));
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 x = p[q];
      v[q].w[0] = x.x; v[q].w[1] = x.y; v[q].w[2] = x.z; v[q].w[3] = x.w;
    }
    issue_one();   // block b + 7 goes to the slot of block b - 1, whose words were consumed before this call
  }
};
__global__ void __launch_bounds__(X3_CRC_THREADS, 1024 / X3_CRC_THREADS) crc_frames_kernel(const DecodeArgs a_in) {   // <= 64 registers
  DecodeArgs a = a_in;
  if (a.len_dev) a.stream_len = *a.len_dev < a.stream_len ? *a.len_dev : a.stream_len;
  __shared__ uint16_t s_T[256];  // byte table, for payloads at odd addresses only
  __shared__ __align__(16) unsigned char s_ring[X3_CRC_THREADS * kCrcRowBytes];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_T[i] = a.crc_tables[i];
  __syncthreads();
  const unsigned long long n = *a.n_frames < a.max_frames ? *a.n_frames : a.max_frames;
  const unsigned long long threads = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += threads) {
    const FrameRec fr = a.frames[f];
    int status = kDecOk;
    const unsigned long long pend = fr.pos + kFrameHeaderLen + fr.payload_len;
    if (pend > a.stream_len) {
      status = kDecErrPanic;  // never produced by the index (truncated frames are dropped); defensive
    } else if (fr.payload_len > a.max_payload) {
      status = kDecErrPayloadLen;  // decodefile.rs:118-121
    } else {
      const uint8_t *pl = a.stream + fr.pos + kFrameHeaderLen;
      const uint32_t len = fr.payload_len;
      uint32_t s;
      if (((uintptr_t)pl & 1u) || (len & 1u) || len == 0u) {
        s = crc16_bytes(s_T, pl, len);  // foreign stream: bytewise
      } else {
        CrcRingSource src;
        src.row = s_ring + threadIdx.x * kCrcRowBytes;
        s = crc16_fold(src, pl, len);
        cp_async_wait_all();
      }
      if ((s & 0xffffu) != fr.payload_crc) status = kDecErrPayloadCrc;  // decodefile.rs:97-100
    }
    a.crc_status[f] = status;
    if (status != kDecOk) atomicMin(a.result, f);
  }
}
#else
// (128 or 64 threads per CTA -- more CTAs beside the decode kernel -- were measured: the pair finishes later)
__global__ void __launch_bounds__(X3_CRC_THREADS) crc_frames_kernel(const DecodeArgs a_in) {
  DecodeArgs a = a_in;
  if (a.len_dev) a.stream_len = *a.len_dev < a.stream_len ? *a.len_dev : a.stream_len;
  __shared__ uint16_t s_T[kCrcTableEntries];
  for (int i = threadIdx.x; i < kCrcTableEntries; i += blockDim.x) s_T[i] = a.crc_tables[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned long long n = *a.n_frames < a.max_frames ? *a.n_frames : a.max_frames;
  const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long f = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n; f += warps) {
    const FrameRec fr = a.frames[f];
    int status = kDecOk;
    const unsigned long long pend = fr.pos + kFrameHeaderLen + fr.payload_len;
    if (pend > a.stream_len) {
      status = kDecErrPanic;  // never produced by the index (truncated frames are dropped); defensive
    } else if (fr.payload_len > a.max_payload) {
      status = kDecErrPayloadLen;  // decodefile.rs:118-121
    } else {
      const uint8_t *pl = a.stream + fr.pos + kFrameHeaderLen;
      const uint32_t len = fr.payload_len;
      uint32_t s;
      if (((uintptr_t)pl & 1u) || (len & 1u)) {
        s = 0xffffu;  // odd placement (foreign stream): bytewise
        if (lane == 0) s = crc16_bytes(s_T, pl, len);
        s = __shfl_sync(0xffffffffu, s, 0);
      } else {
        const uint32_t m = len >> 4;  // whole 16-byte chunks
        uint32_t h = 0;
        if (m > (uint32_t)lane) {
          for (int i = (int)((m - 1u - lane) >> 5); i >= 0; i--) {
            const uint32_t c = m - 1u - (32u * (uint32_t)i + lane);
            const uint8_t *q = pl + 16u * c;
            uint32_t cs = c == 0 ? 0xffffu : 0u;
            if (((uintptr_t)q & 3u) == 0) {
              const uint32_t *q4 = reinterpret_cast<const uint32_t *>(q);
#pragma unroll
              for (int w = 0; w < 4; w++) cs = crc16_word_alu(cs, bswap32(q4[w]));
            } else {
#pragma unroll
              for (int w = 0; w < 4; w++) cs = crc16_word_alu(cs, load_be32_2aligned(q + 4 * w));
            }
            h = crc16_mulc(s_T, 4, h) ^ cs;
          }
        }
#pragma unroll
        for (int k = 0; k < 5; k++) {
          const uint32_t o = __shfl_down_sync(0xffffffffu, h, 1 << k);
          h ^= crc16_mulc(s_T, 6 + 2 * k, o);
        }
        s = m ? h : 0xffffu;
        if (lane == 0) {
          const uint8_t *q = pl + 16u * m;
          const uint32_t rem = len & 15u;
          uint32_t done = 0;
          for (; done + 4u <= rem; done += 4u) s = crc16_word(s_T, s, load_be32_2aligned(q + done));
          if (rem & 2u) {
            const uint32_t v = *reinterpret_cast<const uint16_t *>(q + done);
            s = crc16_half(s_T, s, ((v & 0xff) << 8) | (v >> 8));
          }
        }
        s = __shfl_sync(0xffffffffu, s, 0);
      }
      if ((s & 0xffffu) != fr.payload_crc) status = kDecErrPayloadCrc;  // decodefile.rs:97-100
    }
    if (lane == 0) {
      a.crc_status[f] = status;
      if (status != kDecOk) atomicMin(a.result, f);
    }
  }
}

#endif  // X3_CRC_FOLD

// ------------------------------------------------------------------------------------------------
// 3. decode, one thread per frame
// ------------------------------------------------------------------------------------------------
constexpr int kRingWords = 36;  // 8 chunks of 16 B (32 words) + 4 words of padding: 144 B.  (An odd stride of 33 words with
                                // 4-byte cp.async removes the 4-way bank conflict of the ring reads but was measured slower.)

// Per-lane reader: the lane's payload streams through its own shared-memory ring, filled by cp.async at least
// one block ahead of consumption (a block consumes at most 41 bytes).  Four consecutive big-endian words
// A,B,C,D are kept in registers; the 64-bit window at the current bit position is two funnel shifts of them, and
// when the position crosses a word boundary they shift by one and the next D is loaded -- one full step before it
// can be needed, so the shared-memory latency is off the decoder's dependency chain.
struct RingReader {
  const unsigned char *g0;   // 16-byte aligned global address of chunk 0 (contains the payload's first byte)
  uint32_t *ring;
  const unsigned char *end;  // end of the stream buffer
  uint32_t issued;           // chunks issued so far
  uint32_t n_async;          // chunks [0, n_async) lie wholly inside the stream buffer
  uint32_t pos0, pos;        // bit positions relative to chunk 0: payload start, current
  uint32_t s;                // pos & 31
  uint32_t rb;               // shared-window byte address of the ring
  uint32_t one;              // 1 (DecodeArgs::one)
  uint32_t A, B, C, D;       // big-endian words w, w+1, w+2, w+3 with w = pos >> 5

  __device__ __forceinline__ void issue_to(uint32_t want) {
    while (issued < want) {
      uint32_t *slot = ring + (issued & 7u) * 4u;
      const unsigned char *src = g0 + 16ull * issued;
      if (issued < n_async) {
        cp_async16(slot, src);
      } else {
        for (int w = 0; w < 4; w++) {
          uint32_t v = 0;
          for (int b = 0; b < 4; b++)
            if (src + 4 * w + b < end) v |= (uint32_t)src[4 * w + b] << (8 * b);
          slot[w] = v;
        }
      }
      issued++;
    }
    cp_async_commit();
  }
  __device__ __forceinline__ uint32_t word(uint32_t w) const { return bswap32(ring[w & 31u]); }
  __device__ __forceinline__ void start(const uint8_t *payload, const uint8_t *stream_end, uint32_t *ring_) {
    ring = ring_;
    end = stream_end;
    g0 = reinterpret_cast<const unsigned char *>((uintptr_t)payload & ~(uintptr_t)15);
    const uintptr_t span = (uintptr_t)stream_end - (uintptr_t)g0;
    n_async = (uint32_t)(span >> 4 > 0xffffffffull ? 0xffffffffull : span >> 4);
    issued = 0;
    pos0 = pos = 8u * (uint32_t)((uintptr_t)payload & 15u);
    s = pos & 31u;
    rb = (uint32_t)__cvta_generic_to_shared(ring_);
    issue_to(((pos >> 3) + 112u + 15u) >> 4);
    cp_async_wait_all();
    const uint32_t w = pos >> 5;
    A = word(w); B = word(w + 1); C = word(w + 2); D = word(w + 3);
  }
  __device__ __forceinline__ void block_begin() {
    issue_to(((pos >> 3) + 112u + 15u) >> 4);
    cp_async_wait_1();
  }
  __device__ __forceinline__ void window(uint32_t &hi, uint32_t &lo) const {
    hi = funnel_l(B, A, s);
    lo = funnel_l(C, B, s);
  }
  __device__ __forceinline__ void advance(uint32_t n) {  // n <= 32: crosses at most one word boundary
    const uint32_t s2 = s + n;
    pos += n;
    const bool cross = s2 > 31u;
    // word (pos >> 5) + 3 of the ring, i.e. byte ((pos >> 3) + 12) & 124; only used when crossing, the slot is valid
    // either way
#ifndef X3_DEC_ASMLD
#define X3_DEC_ASMLD 4  // 0: selects; 1, 2: ld.shared from a precomputed 32-bit address (measured 2 % slower); 3: predicated load; 4: see below
#endif
#if X3_DEC_ASMLD == 4
    // the window shifts by one word as four predicated IMADs (x * one + 0): the FMA pipe is idle, the ALU pipe -- where
    // the four selects would go -- is this kernel's top limiter
    const uint32_t nd4 = bswap32(ring[((pos >> 5) + 3u) & 31u]);
    asm("{\n"
        ".reg .pred p;\n"
        "setp.gt.u32 p, %4, 31;\n"
        "@p mad.lo.u32 %0, %1, %5, 0;\n"
        "@p mad.lo.u32 %1, %2, %5, 0;\n"
        "@p mad.lo.u32 %2, %3, %5, 0;\n"
        "@p mad.lo.u32 %3, %6, %5, 0;\n"
        "}\n"
        : "+r"(A), "+r"(B), "+r"(C), "+r"(D)
        : "r"(s2), "r"(one), "r"(nd4));
#elif X3_DEC_ASMLD == 3
    // load only in the lanes that cross a word boundary: fewer lanes per shared-memory access, fewer bank conflicts
    // (the ring positions of the 32 lanes are unrelated, so every active lane is a potential conflict)
    A = cross ? B : A;
    B = cross ? C : B;
    C = cross ? D : C;
    if (cross) D = bswap32(ring[((pos >> 5) + 3u) & 31u]);
#else
    uint32_t nd;
#if X3_DEC_ASMLD == 1
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(nd) : "r"(rb + (((pos >> 3) + 12u) & 124u)) : "memory");
#elif X3_DEC_ASMLD == 2
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(nd) : "r"(rb + (((pos >> 3) + 12u) & 124u)));
#else
    nd = ring[((pos >> 5) + 3u) & 31u];
#endif
    nd = bswap32(nd);
    A = cross ? B : A;
    B = cross ? C : B;
    C = cross ? D : C;
    D = cross ? nd : D;
#endif
    s = s2 & 31u;
  }
  __device__ __forceinline__ uint32_t bits_used() const { return pos - pos0; }
};

__global__ void __launch_bounds__(kDecThreads, X3_DEC_MINBLOCKS) decode_frames_kernel(const DecodeArgs a_in) {
  DecodeArgs a = a_in;
  if (a.len_dev) a.stream_len = *a.len_dev < a.stream_len ? *a.len_dev : a.stream_len;   // stream-ordered API: the length lies on the device
  __shared__ __align__(16) uint32_t s_ring[kDecThreads * kRingWords];
  __shared__ __align__(16) uint32_t s_stage[kDecThreads * kStageWords];  // [word][thread]
  __shared__ __align__(16) inv_entry_t s_inv[kInvTabEntries];
  __shared__ RiceBlockPar s_par[4];
  const int tid = threadIdx.x;
  for (int j = tid; j < kInvTabEntries; j += kDecThreads) s_inv[j] = inv_tab_entry(j);
  if (tid < 4) s_par[tid] = rice_block_par((uint32_t)tid);
  __syncthreads();
  const unsigned long long n = *a.n_frames < a.max_frames ? *a.n_frames : a.max_frames;
  const bool dflt = a.P.block_len == 20 && a.P.codes[0] == 0 && a.P.codes[1] == 1 && a.P.codes[2] == 3;
  const uint8_t *stream_end = a.stream + a.stream_len;
  for (unsigned long long base = (unsigned long long)blockIdx.x * kDecThreads; base < n;
       base += (unsigned long long)gridDim.x * kDecThreads) {
    const unsigned long long i = base + tid;
    if (i >= n) continue;
    const FrameRec fr = a.frames[i];
    // The payload CRC is checked by crc_frames_kernel, which runs beside this kernel on another stream (it fills
    // the SMs this kernel leaves idle while its last frames finish); the host combines the two verdicts.  Frames
    // that kernel refuses to read (decodefile.rs:118-121 and truncation) are not decoded either.
    int status = kDecOk;
    if (fr.pos + kFrameHeaderLen + fr.payload_len <= a.stream_len && fr.payload_len <= a.max_payload) {
      if (fr.samples == 0u || fr.payload_len < 2u) {
        status = kDecErrPanic;
      } else if (fr.out_off + fr.samples > a.pcm_cap) {
        status = kDecErrNoSpace;
      } else {
        const uint8_t *pl = a.stream + fr.pos + kFrameHeaderLen;
        int16_t *out = a.pcm + fr.out_off;
        int r = kDecRetryExact;
        {
          // the tuned path for Parameters::default() frames of whole 80-sample groups; every other frame (other
          // Parameters, a stream's short last frame, an output that is not sector aligned) takes the generic fast path
          RingReader rd;
          rd.one = a.one;
          rd.start(pl, stream_end, s_ring + tid * kRingWords);
          if (dflt && frame_fast_eligible(fr.samples, fr.payload_len, (uintptr_t)pl, (uintptr_t)out))
            r = decode_frame_fast(rd, fr.payload_len, out, fr.samples, s_stage + tid, (uint32_t)kDecThreads, s_inv, s_par);
          else
            r = decode_frame_generic(rd, fr.payload_len, out, fr.samples, a.P);
          cp_async_wait_all();
        }
        if (r == kDecRetryExact) r = decode_frame_exact(pl, fr.payload_len, out, fr.samples, a.P);
        status = r;
      }
    }
    a.frame_status[i] = status;
    if (status != kDecOk) atomicMin(a.result, i);
  }
}


// Stream-ordered API: the caller's 8 words (x3_device_result).  One thread.
__global__ void decode_finalize_kernel(const ScanArgs sa, const DecodeArgs da, unsigned long long *out) {
  const unsigned long long flags = sa.result[2], n = sa.result[4] < da.max_frames ? sa.result[4] : da.max_frames;
  const unsigned long long first_bad = da.result[0];
  unsigned long long samples = sa.result[5], frames = n;
  long long code = 0;
  if (first_bad < n) {
    const int cs = da.crc_status[first_bad], fs = da.frame_status[first_bad];
    code = cs != kDecOk ? cs : fs;   // the reference checks the payload CRC before it decodes (decodefile.rs:93-103)
    samples = da.frames[first_bad].out_off;
    frames = first_bad;
  }
  out[0] = samples;
  out[1] = flags;
  out[2] = frames;
  out[3] = first_bad < n ? first_bad : ~0ull;
  out[4] = (unsigned long long)code;
  out[5] = sa.result[6];   // bytes the walk consumed
  out[6] = 0;
  out[7] = 0;
}

}  // namespace

cudaError_t launch_scan(const ScanArgs &a, bool hop, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (hop) {
    unsigned g = (a.n_tiles + kHopWarps - 1u) / kHopWarps;
    if (g > (unsigned)sms * 8u) g = (unsigned)sms * 8u;
    if (g < 1u) g = 1u;
    hop_index_kernel<<<g, kHopThreads, 0, stream>>>(a);
  } else {
    int grid = sms * 6;
    if ((uint32_t)grid > a.n_tiles) grid = (int)a.n_tiles;
    if (grid < 1) grid = 1;
    scan_headers_kernel<<<grid, kScanThreads, 0, stream>>>(a);
  }
  tile_prefix_kernel<<<1, 1024, 0, stream>>>(a);
  unsigned pg = (a.n_tiles + 7u) / 8u;
  if (pg > (unsigned)sms * 8u) pg = (unsigned)sms * 8u;
  if (pg < 1u) pg = 1u;
  place_frames_kernel<<<pg, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_chain_check(const ScanArgs &a, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  check_chain_kernel<<<sms * 4, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_decode_finalize(const ScanArgs &sa, const DecodeArgs &da, unsigned long long *out8, cudaStream_t stream) {
  decode_finalize_kernel<<<1, 1, 0, stream>>>(sa, da, out8);
  return cudaGetLastError();
}

cudaError_t launch_crc(const DecodeArgs &a, unsigned long long n_frames_hint, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (n_frames_hint < 1) n_frames_hint = 1;
  unsigned long long g = (n_frames_hint * (X3_CRC_FOLD ? 1ull : 32ull) + X3_CRC_THREADS - 1ull) / X3_CRC_THREADS;  // one thread / one warp per frame
  unsigned long long cap = (unsigned long long)sms * (2048ull / X3_CRC_THREADS);
  {
    static const long per_sm = [] { const char *e = getenv("X3_CRC_CTAS_PER_SM"); return e ? atol(e) : 0L; }();   // tuning runs
    if (per_sm > 0) cap = (unsigned long long)sms * (unsigned long long)per_sm;
  }
  if (g > cap) g = cap;
  crc_frames_kernel<<<(unsigned)g, X3_CRC_THREADS, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeArgs &a, unsigned long long n_frames_hint, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (n_frames_hint < 1) n_frames_hint = 1;
  {
    unsigned long long g = (n_frames_hint + kDecThreads - 1) / kDecThreads;
    const unsigned long long cap = (unsigned long long)sms * 16ull;
    if (g > cap) g = cap;
    decode_frames_kernel<<<(unsigned)g, kDecThreads, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace x3
