// x3_kernels.h -- host-visible launch interface of the CUDA kernels (internal to libx3b200.so).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "x3_common.cuh"

namespace x3 {

constexpr int kEncThreads = 512;       // generic kernel
constexpr int kEncFastThreads = 544;   // fast kernel: 16 worker warps + 1 control warp
constexpr int kEncStripThreads = 128;  // strip kernel: 4 warps, one thread per four blocks
enum : int { kEncKernelGeneric = 0, kEncKernelFast = 1, kEncKernelStrip = 2 };
#ifndef X3_DEC_THREADS
#define X3_DEC_THREADS 128
#endif
#ifndef X3_DEC_MINBLOCKS
#define X3_DEC_MINBLOCKS 7
#endif
constexpr int kDecThreads = X3_DEC_THREADS;
constexpr int kScanThreads = 256;

struct EncodeArgs {
  const int16_t *pcm;
  unsigned long long n_samples;
  uint8_t *out;
  unsigned long long out_cap;
  CodecParams P;
  uint32_t n_frames;
  uint32_t max_blocks;      // blocks in a full frame
  uint32_t out_words_cap;   // 32-bit words of payload image a frame can need
  unsigned long long *status;  // [n_frames] look-back words, zeroed before launch
  unsigned int *ticket;        // zeroed before launch
  unsigned long long *result;  // [0] total bytes, [1] overflow flag, [2..8) stats; zeroed before launch
  unsigned long long *timing;  // 32 words, only written when built with -DX3_ENC_TIMING
  const uint16_t *crc_tables;  // kCrcBankEntries3 (normal bank, byte-swapped bank, nibble multiply tables)
  int32_t neg_one;             // -1, opaque to the compiler: lets the kernel compute ~x as x*(-1)-1 on the FMA pipe
  const unsigned int *choice;  // null, or the kernel kind encode_probe_kernel picked: a kernel of another kind returns at once
};

size_t encode_smem_bytes(const CodecParams &P, uint32_t max_blocks, uint32_t out_words_cap);
size_t encode_fast_smem_bytes(const CodecParams &P, uint32_t out_words_cap);
size_t encode_strip_smem_bytes();
int encode_occupancy(int kind, size_t smem);
cudaError_t launch_encode(const EncodeArgs &a, int kind, int grid, size_t smem, cudaStream_t stream);
// Zeroes the encoder workspace [ws, ws + ws_bytes) and picks the kernel for a Parameters::default() input: it codes
// 256 blocks spread evenly over the input and writes kEncKernelFast to *choice when more than
// kProbeBigPercent of them are BFP / literal blocks (frames of such input overflow the strip kernel's 8 KiB windows
// and take its slow paths; the block-per-thread kernel keeps a whole frame image in shared memory), else
// kEncKernelStrip.
constexpr unsigned kProbeBigPercent = 20;
cudaError_t launch_encode_probe(const int16_t *pcm, unsigned long long n_samples, unsigned char *ws, size_t ws_bytes,
                                unsigned int *choice, cudaStream_t stream);

// One frame of a stream, as found by the frame index (device scan or host walk).
struct FrameRec {
  unsigned long long pos;      // byte offset of the frame header in the stream
  unsigned long long out_off;  // sample offset of the frame's first sample in the PCM output
  uint32_t samples;            // header.samples
  uint32_t payload_len;        // header.payload_len
  uint32_t payload_crc;        // header.payload_crc
  uint32_t pad;
};

struct DecodeArgs {
  const uint8_t *stream;
  unsigned long long stream_len;
  int16_t *pcm;
  unsigned long long pcm_cap;
  CodecParams P;
  const FrameRec *frames;
  const unsigned long long *n_frames;  // device scalar (the index kernel produces it)
  unsigned long long max_frames;       // capacity of `frames` / status
  int *frame_status;                   // [max_frames] verdict of decode_frames_kernel
  int *crc_status;                     // [max_frames] verdict of crc_frames_kernel (runs concurrently with the decode;
                                       // a frame's status is crc_status if that is an error, else frame_status)
  unsigned long long *result;          // [0] first bad frame of either kernel (init ~0), [1] unused
  uint32_t one;                        // 1, opaque to the compiler: x * one + 0 is a register move on the FMA pipe
  uint32_t max_payload;                // X3_READ_BUFFER_SIZE for a stream (decodefile.rs:118-121); no such limit for
                                       // decoder::decode_frame on its own
  const uint16_t *crc_tables;
  const unsigned long long *len_dev;   // null, or where the stream's length lies on the device (stream-ordered API):
                                       // read by the kernels in place of stream_len, which is then only an upper bound
};

cudaError_t launch_crc(const DecodeArgs &a, unsigned long long n_frames_hint, cudaStream_t stream);
cudaError_t launch_decode(const DecodeArgs &a, unsigned long long n_frames_hint, cudaStream_t stream);

struct ScanArgs {
  const uint8_t *stream;
  unsigned long long stream_len;
  FrameRec *frames;
  unsigned long long max_frames;
  unsigned long long *tile_status;  // [n_tiles] frames << 38 | samples of the tile, then its exclusive prefix
  unsigned long long *tile_recs;    // [n_tiles] (first record << 11) | records of the tile
  FrameRec *recs;                   // [max_frames] the tiles' records in arrival order (sample offsets tile-relative)
  unsigned long long *rec_cursor;   // zeroed
  unsigned int *ticket;             // zeroed
  unsigned long long *result;       // [0] n_frames, [1] total samples, [2] flags (bit 0: the table is not proven, needs
                                    // the host walk or a retry; bit 1: because a capacity -- candidates per tile, table size -- was exceeded)
  const uint16_t *crc_tables;
  uint32_t n_tiles;
  uint32_t tile_bytes;              // kScanTileBytes, or kScanTileBytesSmall on the retry (multiple of 16)
  const unsigned long long *len_dev;   // see DecodeArgs
};
constexpr uint32_t kScanTileBytes = 128 * 1024;
constexpr uint32_t kScanTileBytesSmall = 16 * 1024;   // >= 22 bytes per frame: at most 745 frames per tile, under the cap
constexpr uint32_t kHopTileBytes = 64 * 1024;         // hop index: one warp per tile, at most 64 frames in a tile
cudaError_t launch_scan(const ScanArgs &a, bool hop, cudaStream_t stream);  // hop: follow the headers instead of reading every byte
cudaError_t launch_chain_check(const ScanArgs &a, cudaStream_t stream);
// Stream-ordered API: fold the index's and the kernels' verdicts into the caller's 8-word result (x3_device_result).
cudaError_t launch_decode_finalize(const ScanArgs &sa, const DecodeArgs &da, unsigned long long *out8, cudaStream_t stream);

cudaError_t launch_synth(int kind, uint32_t seed, uint32_t fs, unsigned long long n0, unsigned long long count,
                         int16_t *out, cudaStream_t stream);

}  // namespace x3
