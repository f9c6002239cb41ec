/*
 * x3_oracle.h -- CPU restatement of the X3 codec hot path of psiphi75/x3-rust.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (x3-rust_b200/, include/)
 * links, imports or executes this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it, as the checker.
 *
 * Parity pinning: the reference cannot be compiled here (no cargo/rustc, no vendored
 * crates).  This restatement is pinned against every golden vector the reference's own
 * unit tests hold for this path (encoder.rs:341-620, decoder.rs:256-355,
 * bitpacker.rs:196-289, bitreader.rs:194-304, crc.rs:77-106), transcribed under
 * tests/golden/ and checked by tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line (under /root/reference/src) it follows.
 */
#ifndef X3_ORACLE_H
#define X3_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: 0 OK, negative = X3Error variant (error.rs:27-62) */
enum {
  X3O_OK = 0,
  X3O_ERR_INVALID_ENCODING_THRESH = -1,   /* x3.rs:107-112 */
  X3O_ERR_OUT_OF_BOUNDS_INVERSE = -2,     /* decoder.rs:161,187 */
  X3O_ERR_MORE_THAN_ONE_CHANNEL = -3,     /* decoder.rs:92 */
  X3O_ERR_ARCHIVE_XML_INVALID = -4,
  X3O_ERR_ARCHIVE_XML_RICE_CODE = -5,
  X3O_ERR_ARCHIVE_INVALID_KEY = -6,
  X3O_ERR_FRAME_LENGTH = -7,              /* decoder.rs:101 */
  X3O_ERR_FRAME_HEADER_INVALID_KEY = -8,  /* decoder.rs:82 */
  X3O_ERR_FRAME_HEADER_INVALID_PAYLOAD_LEN = -9, /* decodefile.rs:118 */
  X3O_ERR_FRAME_HEADER_INVALID_HEADER_CRC = -10, /* decoder.rs:76 */
  X3O_ERR_FRAME_HEADER_INVALID_PAYLOAD_CRC = -11, /* decodefile.rs:99 */
  X3O_ERR_FRAME_DECODE_INVALID_FTYPE = -12,
  X3O_ERR_FRAME_DECODE_INVALID_BPF = -13, /* decoder.rs:215 */
  X3O_ERR_FRAME_DECODE_UNEXPECTED_END = -14, /* decoder.rs:71 */
  X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY = -15, /* bytewriter.rs:72,88 */
  X3O_ERR_IO = -16,                       /* std::io::Error (read_exact past EOF) */
  X3O_ERR_REFERENCE_PANIC = -100          /* the reference would panic (index out of bounds etc.) */
};

typedef struct {
  uint32_t block_len;
  uint32_t blocks_per_frame;
  uint32_t codes[3];
  uint32_t thresholds[3];
} x3o_params;

typedef struct {
  uint8_t source_id;
  uint16_t samples;
  uint8_t channels;
  uint32_t payload_len;
  uint16_t payload_crc;
} x3o_frame_header;

/* x3.rs:124-134 */
void x3o_params_default(x3o_params *p);
/* x3.rs:99-122 (Parameters::new); also rejects codes > 3 (RiceCodes::get would panic) */
int x3o_params_validate(const x3o_params *p);

/* crc.rs:44-58 */
uint16_t x3o_update_crc16(uint16_t crc, uint8_t data);
uint16_t x3o_crc16(const uint8_t *data, size_t len);

/* x3.rs:207-252: table accessors (tables are built from the closed form and cross-checked
 * against the transcribed reference tables in tests/golden/rice_tables.json). */
int x3o_rice_table(uint32_t code_id, uint32_t *nsubs, uint32_t *offset, uint32_t *n,
                   uint32_t *code /*[56]*/, uint32_t *num_bits /*[56]*/, uint32_t *inv_len);
int16_t x3o_inv_rice(uint32_t i); /* x3.rs:200-204, i < 60 */

/* bitpacker.rs: write a sequence of (value,num_bits) into buf (zero-initialised by caller is NOT
 * required; bytes are assigned, like SliceByteWriter) and return bytes written incl. final drop flush.
 * Used to replay bitpacker.rs:196-289. */
int x3o_bitpack(const uint64_t *values, const uint32_t *num_bits, size_t n,
                uint8_t *buf, size_t cap, size_t *out_len);

/* bitreader.rs: scripted replay.  ops[i] = n>0: read_nbits(n); n==0: count_zero_bits().
 * results[i] gets the value; lead[i], rem[i] get leading_word / rem_bit after the op. */
int x3o_bitread(const uint8_t *buf, size_t len, const uint32_t *ops, size_t n_ops,
                uint32_t *results, uint32_t *lead, uint32_t *rem);

/* encoder.rs:122-162 */
void x3o_write_frame_header(size_t num_samples, uint8_t id, size_t payload_len,
                            uint16_t payload_crc, uint8_t header[20]);

/* encoder.rs:289-315 (x3_encode_block) applied to wav[1..] with diffs from wav (n = block len + 1),
 * starting after `lead_zero_bits` packed zeros, followed by word_align; replays encoder.rs:494-620. */
int x3o_encode_block_test(const int16_t *wav, size_t n, const x3o_params *p, uint32_t lead_zero_bits,
                          uint8_t *buf, size_t cap, size_t *out_len);

/* encoder.rs:175-214: encode one frame at writer position *pos (SliceByteWriter over buf[0..cap)). */
int x3o_encode_frame(const int16_t *wav, size_t n, const x3o_params *p, uint8_t *buf, size_t cap,
                     size_t *pos, uint64_t stats[6]);

/* encoder.rs:51-111: split into frames and encode all (single channel), starting at *pos. */
int x3o_encode(const int16_t *wav, size_t n, const x3o_params *p, uint8_t *buf, size_t cap,
               size_t *pos, uint64_t stats[6]);

/* worst-case bytes encode() can write for n samples (not in the reference; sizing helper) */
size_t x3o_encode_bound(size_t n, const x3o_params *p);

/* decoder.rs:69-118 */
int x3o_read_frame_header(const uint8_t *bytes, size_t len, x3o_frame_header *h);

/* decoder.rs:132-145 on a BitReader over buf, after skipping skip_bits; replays decoder.rs:256-355 */
int x3o_decode_block_test(const uint8_t *buf, size_t len, uint32_t skip_bits, int16_t last_wav,
                          const x3o_params *p, int16_t *wav, size_t block_len);

/* decoder.rs:36-58 */
int x3o_decode_frame(const uint8_t *payload, size_t payload_len, int16_t *wav, size_t wav_cap,
                     const x3o_params *p, size_t samples, size_t *n_out);

/* Frame loop of decodefile.rs:105-136 + x3a_to_wav loop (:202-209) over an in-memory frame stream.
 * `remaining0` is the reader's initial remaing_bytes (for a bare frame stream pass len; for a file
 * see x3o_x3a_decode).  Returns 0 or the error the reference would propagate; *n_out = samples
 * written before stopping; *frames_ok = frames decoded; *frame_errors as decodefile.rs:131. */
int x3o_decode_stream(const uint8_t *bytes, size_t len, size_t remaining0, const x3o_params *p,
                      int16_t *wav, size_t wav_cap, size_t *n_out, size_t *frames_ok,
                      size_t *frame_errors);

/* encodefile.rs:82-138: archive id + header frame + XML (+pad). Returns bytes written. */
int x3o_archive_header(uint32_t sample_rate, const x3o_params *p, uint8_t *buf, size_t cap,
                       size_t *out_len);
/* decodefile.rs:142-176 + 232-303 (parse). *header_size = 20 + payload_len (NOT incl. the 8-byte id,
 * reproducing decodefile.rs:61-65,166). */
int x3o_archive_parse(const uint8_t *bytes, size_t len, uint32_t *sample_rate, x3o_params *p,
                      size_t *header_size, size_t *frames_offset);

/* whole-file: wav samples -> .x3a bytes (encodefile.rs:48-78) and back (decodefile.rs:189-212) */
int x3o_x3a_encode(const int16_t *wav, size_t n, uint32_t sample_rate, uint8_t *buf, size_t cap,
                   size_t *out_len, uint64_t stats[6]);
int x3o_x3a_decode(const uint8_t *bytes, size_t len, int16_t *wav, size_t wav_cap, size_t *n_out,
                   uint32_t *sample_rate, size_t *frames_ok, size_t *frame_errors);

/* multi-threaded frame-parallel variants for the CPU baseline (not in the reference, which is
 * single-threaded; frames are independent so the bytes are identical). */
int x3o_encode_mt(const int16_t *wav, size_t n, const x3o_params *p, uint8_t *buf, size_t cap,
                  size_t *out_len, uint64_t stats[6], int threads);
int x3o_decode_stream_mt(const uint8_t *bytes, size_t len, const x3o_params *p, int16_t *wav,
                         size_t wav_cap, size_t *n_out, int threads);

/* Synthetic signals of SURVEY.md section 8(d).  kind: 1=S1, 2=S2 (also S5 with fs=96000), 4=S4.
 * Fills out[0..count) with samples n0 .. n0+count-1. */
int x3o_synth(int kind, uint32_t seed, uint32_t fs, uint64_t n0, uint64_t count, int16_t *out);

#ifdef __cplusplus
}
#endif
#endif
