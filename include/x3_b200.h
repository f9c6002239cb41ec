/*
 * x3_b200.h -- C ABI of the B200-native X3 codec hot path (libx3b200.so).
 *
 * This is the drop-in boundary for the frame encode / frame decode path of psiphi75/x3-rust.
 * The reference has no FFI of its own (it is a pure-Rust crate); each entry point below names the
 * reference Rust interface it replaces (file:line under /root/reference/src) and INTEGRATION.md shows
 * the `extern "C"` block + build.rs a maintainer adds to bind it.
 *
 * Conventions: plain pointers and sizes, no exceptions, return code 0 = OK, negative = error
 * (one-to-one with X3Error variants, error.rs:27-62, plus CUDA/usage errors below -100).
 * No CPU fallback exists: every entry point that computes needs a CUDA device (sm_100a) and
 * returns X3_ERR_CUDA when none is usable.
 *
 * Threading: entry points are re-entrant; concurrent calls may share a device. `*_device` calls are
 * stream-ordered on `cuda_stream` (a cudaStream_t, NULL = default stream) and synchronise that stream
 * once before returning, because they report lengths / error codes to the host; the `*_device_async`
 * variants only enqueue work and leave their results on the device.
 */
#ifndef X3_B200_H
#define X3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X3_B200_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------------------------- */
enum {
  X3_OK = 0,
  X3_ERR_INVALID_ENCODING_THRESH = -1,          /* X3Error::InvalidEncodingThresh, x3.rs:107-112 */
  X3_ERR_OUT_OF_BOUNDS_INVERSE = -2,            /* X3Error::OutOfBoundsInverse, decoder.rs:161,187 */
  X3_ERR_MORE_THAN_ONE_CHANNEL = -3,            /* X3Error::MoreThanOneChannel, encoder.rs:55, decoder.rs:92 */
  X3_ERR_ARCHIVE_XML_INVALID = -4,              /* X3Error::ArchiveHeaderXMLInvalid */
  X3_ERR_ARCHIVE_XML_RICE_CODE = -5,            /* X3Error::ArchiveHeaderXMLRiceCode */
  X3_ERR_ARCHIVE_INVALID_KEY = -6,              /* X3Error::ArchiveHeaderXMLInvalidKey */
  X3_ERR_FRAME_LENGTH = -7,                     /* X3Error::FrameLength, decoder.rs:101 */
  X3_ERR_FRAME_HEADER_INVALID_KEY = -8,         /* X3Error::FrameHeaderInvalidKey, decoder.rs:82 */
  X3_ERR_FRAME_HEADER_INVALID_PAYLOAD_LEN = -9, /* X3Error::FrameHeaderInvalidPayloadLen, decodefile.rs:118 */
  X3_ERR_FRAME_HEADER_INVALID_HEADER_CRC = -10, /* X3Error::FrameHeaderInvalidHeaderCRC, decoder.rs:76 */
  X3_ERR_FRAME_HEADER_INVALID_PAYLOAD_CRC = -11,/* X3Error::FrameHeaderInvalidPayloadCRC, decodefile.rs:99 */
  X3_ERR_FRAME_DECODE_INVALID_FTYPE = -12,      /* X3Error::FrameDecodeInvalidFType */
  X3_ERR_FRAME_DECODE_INVALID_BPF = -13,        /* X3Error::FrameDecodeInvalidBPF, decoder.rs:215 */
  X3_ERR_FRAME_DECODE_UNEXPECTED_END = -14,     /* X3Error::FrameDecodeUnexpectedEnd, decoder.rs:71 */
  X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY = -15,  /* X3Error::ByteWriterInsufficientMemory, bytewriter.rs:72,88 */
  X3_ERR_IO = -16,                              /* X3Error::Io (read past the end of the stream) */
  /* not X3Error variants */
  X3_ERR_INVALID_ARGUMENT = -101,               /* NULL pointer, zero block_len, code id > 3 ... */
  X3_ERR_UNSUPPORTED_PARAMS = -102,             /* valid for the reference but outside the GPU path's limits */
  X3_ERR_CUDA = -103,                           /* CUDA runtime failure; x3_last_cuda_error() has the text */
  X3_ERR_REFERENCE_PANIC = -104                 /* input on which the reference panics (e.g. a zero-sample frame) */
};

/* x3::Parameters (x3.rs:81-87) without the derived rice_codes pointers. */
typedef struct {
  uint32_t block_len;
  uint32_t blocks_per_frame;
  uint32_t codes[3];
  uint32_t thresholds[3];
} x3_params;

/* `stats: [usize; 6]` of encoder.rs:63 -- samples coded as Rice-0..3, BFP, pass-through. */
typedef struct {
  uint64_t samples_by_mode[6];
} x3_stats;

/* decoder::FrameHeader (x3.rs:147-162). */
typedef struct {
  uint8_t source_id;
  uint8_t channels;
  uint16_t samples;
  uint32_t payload_len;
  uint16_t payload_crc;
} x3_frame_header;

/* Result of a stream decode: what decodefile.rs:105-136 + :202-209 would have produced. */
typedef struct {
  uint64_t samples;          /* PCM samples written (all frames before the first bad one) */
  uint64_t frames;           /* frames decoded */
  uint64_t frame_errors;     /* X3aReader.frame_errors (decodefile.rs:131): 0 or 1 */
  uint64_t first_bad_frame;  /* index of the frame that stopped the decode, UINT64_MAX if none */
  int32_t first_bad_code;    /* the X3_ERR_* of that frame, 0 if none */
  int32_t used_host_walk;    /* 1 if frame discovery fell back to the sequential host header walk */
} x3_decode_result;

/* ---- parameters (no GPU needed) ---------------------------------------------------------- */
int x3_abi_version(void);
/* Parameters::default(), x3.rs:124-134 */
int x3_params_default(x3_params *p);
/* Parameters::new(), x3.rs:99-122 (InvalidEncodingThresh); also INVALID_ARGUMENT / UNSUPPORTED_PARAMS
 * for values the GPU path cannot take (see DESIGN.md "limits"). */
int x3_params_validate(const x3_params *p);
/* Upper bound on the bytes encoder::encode writes for n samples; the reference offers none
 * (README.md:46-47 suggests n*2, which is too small for incompressible input). */
size_t x3_encode_bound(size_t n_samples, const x3_params *p);
/* The same for encoder::encode_frame (encoder.rs:175): ONE frame holding all n_samples (<= 65535), whatever
 * blocks_per_frame says.  0 for parameters x3_params_validate rejects. */
size_t x3_encode_frame_bound(size_t n_samples, const x3_params *p);
const char *x3_strerror(int code);
const char *x3_last_cuda_error(void);

/* ---- header helpers (host, no GPU) ------------------------------------------------------- */
/* encoder::write_frame_header, encoder.rs:122-162 */
int x3_write_frame_header(size_t num_samples, uint8_t id, size_t payload_len, uint16_t payload_crc,
                          uint8_t header[20]);
/* decoder::read_frame_header, decoder.rs:69-118 */
int x3_read_frame_header(const uint8_t *bytes, size_t len, x3_frame_header *h);
/* crc::crc16, crc.rs:49-58 */
uint16_t x3_crc16(const uint8_t *data, size_t len);

/* ---- encode ------------------------------------------------------------------------------ */
/* encoder::encode (encoder.rs:51-111) for one contiguous channel: splits `pcm` into frames of
 * block_len*blocks_per_frame samples and writes header+payload for every frame to out[0..*out_len).
 * The first frame starts at out[0] (an even stream position, encoder.rs:182).
 * Host pointers: input is copied to the device in pipelined chunks and frames are copied back. */
int x3_encode_host(const int16_t *pcm, size_t n_samples, const x3_params *p, uint8_t *out,
                   size_t out_cap, size_t *out_len, x3_stats *stats);
/* Device pointers: zero-copy, stream-ordered; d_pcm and d_out are device memory. */
int x3_encode_device(const int16_t *d_pcm, size_t n_samples, const x3_params *p, uint8_t *d_out,
                     size_t out_cap, size_t *out_len, x3_stats *stats, void *cuda_stream);
/* encoder::encode_frame (encoder.rs:175-214): exactly one frame from all n_samples (<= 65535). */
int x3_encode_frame_host(const int16_t *pcm, size_t n_samples, const x3_params *p, uint8_t *out,
                         size_t out_cap, size_t *out_len, x3_stats *stats);

/* ---- decode ------------------------------------------------------------------------------ */
/* The frame loop of X3aReader::decode_next_frame (decodefile.rs:105-136) driven as x3a_to_wav does
 * (decodefile.rs:202-209) over a bare frame stream (no archive header): header check, payload CRC
 * check, decoder::decode_frame (decoder.rs:36-58) for every frame, stopping at the first bad frame.
 * Returns 0 or the error the reference would propagate; `res` (optional) has the details. */
int x3_decode_host(const uint8_t *frames, size_t len, const x3_params *p, int16_t *pcm,
                   size_t pcm_cap, size_t *n_out, x3_decode_result *res);
int x3_decode_device(const uint8_t *d_frames, size_t len, const x3_params *p, int16_t *d_pcm,
                     size_t pcm_cap, size_t *n_out, x3_decode_result *res, void *cuda_stream);

/* ---- stream-ordered variants (ABI v3): nothing is read back, nothing synchronises ------------ */
/* For callers that keep the data on the device and chain calls on one CUDA stream (encode -> decode, or many
 * batches back to back): the calls only enqueue work.  Results stay on the device in an x3_device_result the
 * caller owns (device memory, 8-byte aligned); it is valid once `cuda_stream` has reached the end of the call.
 * The decode takes the stream's length from device memory (`d_len`, e.g. &encode_result->value) -- `len_cap` is an
 * upper bound used to size the workspace -- so an encode and the decode of its output need no host round trip.
 * These entry points do not fall back: where x3_decode_device would retry with a larger frame table or walk the
 * stream on the host (flags != 0), they report the flag and the caller repeats the call with x3_decode_device.
 *   encode: value = bytes written, flags bit 0 = output capacity exceeded (nothing usable), detail[0..6) =
 *           x3_stats.samples_by_mode.
 *   decode: value = samples written (all frames before the first bad one), flags bit 0 = the frame table could not
 *           be proven to be the reference's walk, bit 1 = a capacity (frames per tile / table size) was exceeded,
 *           detail[0] = frames decoded, [1] = first bad frame (or ~0), [2] = its kDec* status as a signed value
 *           (-11 payload CRC, -9 payload length, -2 / -13 in-frame decode errors, -15 output capacity), [3] = bytes
 *           of the stream the walk consumed. */
typedef struct x3_device_result {
  uint64_t value;
  uint64_t flags;
  uint64_t detail[6];
} x3_device_result;
int x3_encode_device_async(const int16_t *d_pcm, size_t n_samples, const x3_params *p, uint8_t *d_out,
                           size_t out_cap, x3_device_result *d_res, void *cuda_stream);
int x3_decode_device_async(const uint8_t *d_frames, size_t len_cap, const uint64_t *d_len, const x3_params *p,
                           int16_t *d_pcm, size_t pcm_cap, x3_device_result *d_res, void *cuda_stream);

/* decoder::decode_frame (decoder.rs:36-58): one payload (no header), `samples` from its header.  Like the
 * reference's function it has no 24 KiB payload limit (that one belongs to X3aReader, decodefile.rs:118-121):
 * any payload below Frame::MAX_LENGTH is taken. */
int x3_decode_frame_host(const uint8_t *payload, size_t payload_len, const x3_params *p,
                         int16_t *pcm, size_t pcm_cap, size_t samples, size_t *n_out);

/* ---- frame-range sharding across GPUs ----------------------------- */
/* Frames are independent (own first sample, own CRCs, even length: encoder.rs:175-214), so a recording is cut at
 * frame boundaries, every rank encodes / decodes its own range with the *_device calls above, and the only thing
 * ranks exchange is the compressed size of their shard (one NCCL / MPI all-gather of a uint64, done by the host
 * program).  These helpers are the arithmetic around that exchange; none of them needs a GPU except the copy.
 *
 * Sample range [*s0, *s1) of rank's shard of an n_samples recording: frames [rank*F/world, (rank+1)*F/world). */
int x3_shard_range(uint64_t n_samples, const x3_params *p, uint32_t rank, uint32_t world, uint64_t *s0, uint64_t *s1);
/* Whole files dealt by cumulative frame count (a file keeps its own archive header, so files are never split):
 * rank_of_file[i] = the rank that takes file i. */
int x3_deal_files(const uint64_t *frames_per_file, size_t n_files, uint32_t world, uint32_t *rank_of_file);
/* Byte offset of rank's shard in the concatenated stream, from the all-gathered shard sizes. */
int x3_shard_base(const uint64_t *shard_bytes, uint32_t world, uint32_t rank, uint64_t *base);
/* Optional gather: copy a shard's frame bytes to d_stream + base, where d_stream may live on another GPU of the box
 * (peer copy over NVLink; unified addressing picks the route).  Stream-ordered on cuda_stream, no synchronisation. */
int x3_place_shard_device(uint8_t *d_stream, uint64_t base, const uint8_t *d_shard, size_t shard_bytes, void *cuda_stream);

/* ---- synthetic signals (bench / tests) ---------------------------- */
/* Fill d_out[0..count) on the device with samples n0..n0+count-1 of generator `kind`
 * (1 = S1, 2 = S2/S5, 4 = S4).  Integer-only; bit-identical to oracle/x3o_synth. */
int x3_synth_device(int kind, uint32_t seed, uint32_t fs, uint64_t n0, uint64_t count,
                    int16_t *d_out, void *cuda_stream);

/* ---- introspection ----------------------------------------------------------------------- */
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t x3_kernel_launch_count(void);
/* Device time (ms, CUDA events) of the kernels of the most recent *_device / *_host call on this
 * thread: [0] encode_frames / decode_frames kernel, [1] frame index (scan + chain check, decode only),
 * [2] whole device section, [3] payload CRC kernel (decode only). */
int x3_last_kernel_ms(float ms[4]);
/* Which encode kernel the most recent x3_encode_device call on this thread ran when the choice was made on the
 * device (Parameters::default() inputs of 64 frames and more): 1 = block-per-thread kernel (inputs dominated by BFP /
 * literal blocks), 2 = strip kernel; negative when no choice was made (other parameters, small or forced calls). */
int x3_last_encode_kernel(void);

#ifdef __cplusplus
}
#endif
#endif
