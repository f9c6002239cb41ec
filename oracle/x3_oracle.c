/*
 * x3_oracle.c -- CPU restatement of the X3 codec hot path of psiphi75/x3-rust.
 *
 * TEST INFRASTRUCTURE ONLY (see x3_oracle.h).  It is written to be obviously equal to the
 * reference: byte-at-a-time bit packer with a one-byte scratch, table-driven Rice coder,
 * word-structured bit reader with the reference's refill quirks.  Citations are file:line under
 * /root/reference/src.  Parity is pinned by the reference's own unit-test vectors
 * (tests/golden/, tests/test_oracle_golden.py); the reference itself cannot be built here.
 */
#include "x3_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "x3_sin1024.h"

#define FRAME_HEADER_LEN 20        /* x3.rs:166 */
#define FRAME_KEY 30771            /* x3.rs:169  "x3" */
#define FRAME_MAX_LENGTH 0x7fe0    /* x3.rs:145 */
#define MAX_BLOCK_LENGTH 60        /* x3.rs:90 */
#define X3_READ_BUFFER_SIZE (1024 * 24)                 /* decodefile.rs:44 */
#define X3_WRITE_BUFFER_SIZE (X3_READ_BUFFER_SIZE * 8)  /* decodefile.rs:45 */

/* ------------------------------------------------------------------------------------------ */
/* x3.rs:81-134 Parameters                                                                     */
/* ------------------------------------------------------------------------------------------ */

void x3o_params_default(x3o_params *p) {
  p->block_len = 20;        /* x3.rs:93 */
  p->blocks_per_frame = 500; /* x3.rs:96 */
  p->codes[0] = 0; p->codes[1] = 1; p->codes[2] = 3;            /* x3.rs:94 */
  p->thresholds[0] = 3; p->thresholds[1] = 8; p->thresholds[2] = 20; /* x3.rs:95 */
}

/* ------------------------------------------------------------------------------------------ */
/* x3.rs:186-261 Rice code tables.  Built from the closed form                                 */
/*   fold u = d<0 ? -2d-1 : 2d ; codeword = (1<<k) | (u & ((1<<k)-1)) in (u>>k)+k+1 bits       */
/* over the reference's table domains; tests cross-check them against the transcribed tables.  */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint32_t nsubs, offset, n, inv_len;
  uint32_t code[56], num_bits[56];
} rice_code;

static rice_code RICE[4];
static int16_t INV_RICE_CODE[60];
static int tables_ready = 0;
static uint16_t CRC_TABLE[256];

static void build_tables(void) {
  if (tables_ready) return;
  static const uint32_t offs[4] = {6, 11, 20, 28};      /* x3.rs:210,218,226,240 */
  static const uint32_t lens[4] = {14, 22, 40, 56};      /* table lengths x3.rs:211-250 */
  static const uint32_t invl[4] = {16, 26, 44, 60};      /* x3.rs:214,222,236,250 */
  for (int k = 0; k < 4; k++) {
    RICE[k].nsubs = (uint32_t)k;
    RICE[k].offset = offs[k];
    RICE[k].n = lens[k];
    RICE[k].inv_len = invl[k];
    for (uint32_t ii = 0; ii < lens[k]; ii++) {
      int d = (int)ii - (int)offs[k];
      uint32_t u = d < 0 ? (uint32_t)(-2 * d - 1) : (uint32_t)(2 * d);
      RICE[k].code[ii] = (1u << k) | (u & ((1u << k) - 1u));
      RICE[k].num_bits[ii] = (u >> k) + (uint32_t)k + 1u;
    }
  }
  for (int i = 0; i < 60; i++) /* x3.rs:200-204 */
    INV_RICE_CODE[i] = (int16_t)((i & 1) ? -(i + 1) / 2 : i / 2);
  /* crc.rs:22-42: CRC-16/CCITT-FALSE table, poly 0x1021 */
  for (int b = 0; b < 256; b++) {
    uint16_t c = (uint16_t)(b << 8);
    for (int j = 0; j < 8; j++) c = (uint16_t)((c & 0x8000) ? ((c << 1) ^ 0x1021) : (c << 1));
    CRC_TABLE[b] = c;
  }
  __sync_synchronize();
  tables_ready = 1;
}

int x3o_rice_table(uint32_t code_id, uint32_t *nsubs, uint32_t *offset, uint32_t *n,
                   uint32_t *code, uint32_t *num_bits, uint32_t *inv_len) {
  build_tables();
  if (code_id > 3) return X3O_ERR_REFERENCE_PANIC;
  const rice_code *rc = &RICE[code_id];
  *nsubs = rc->nsubs; *offset = rc->offset; *n = rc->n; *inv_len = rc->inv_len;
  memcpy(code, rc->code, rc->n * sizeof(uint32_t));
  memcpy(num_bits, rc->num_bits, rc->n * sizeof(uint32_t));
  return X3O_OK;
}

int16_t x3o_inv_rice(uint32_t i) { build_tables(); return INV_RICE_CODE[i]; }

/* x3.rs:99-122 Parameters::new */
int x3o_params_validate(const x3o_params *p) {
  build_tables();
  for (int k = 0; k < 3; k++)
    if (p->codes[k] > 3) return X3O_ERR_REFERENCE_PANIC; /* RiceCodes::get indexes CODE[..] x3.rs:256 */
  for (int k = 0; k < 2; k++) /* x3.rs:107 `for k in 0..2` */
    if (p->thresholds[k] > RICE[p->codes[k]].offset) return X3O_ERR_INVALID_ENCODING_THRESH;
  return X3O_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* crc.rs:44-58                                                                                */
/* ------------------------------------------------------------------------------------------ */

uint16_t x3o_update_crc16(uint16_t crc, uint8_t data) {
  build_tables();
  uint8_t lookup = (uint8_t)(data ^ (uint8_t)(crc >> 8)); /* crc.rs:45 */
  return (uint16_t)((uint16_t)(crc << 8) ^ CRC_TABLE[lookup]); /* crc.rs:46 */
}

uint16_t x3o_crc16(const uint8_t *data, size_t len) {
  uint16_t crc = 0xffff; /* crc.rs:50 */
  for (size_t i = 0; i < len; i++) crc = x3o_update_crc16(crc, data[i]);
  return crc;
}

/* ------------------------------------------------------------------------------------------ */
/* bytewriter.rs:27-100 SliceByteWriter                                                        */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint8_t *slice;
  size_t len;
  size_t p_byte;
  size_t stream_length;
} slice_writer;

static int sw_write_all(slice_writer *w, const uint8_t *v, size_t n) {
  if (n > w->len - w->p_byte) return X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY; /* bytewriter.rs:88 */
  memcpy(w->slice + w->p_byte, v, n);
  w->p_byte += n;
  if (w->p_byte > w->stream_length) w->stream_length = w->p_byte;
  return X3O_OK;
}

static int sw_seek_abs(slice_writer *w, size_t abs_pos) { /* bytewriter.rs:58-80 */
  if (abs_pos > w->len) return X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY;
  w->p_byte = abs_pos;
  if (w->p_byte > w->stream_length) w->stream_length = w->p_byte;
  return X3O_OK;
}

static int sw_align2(slice_writer *w) { /* bytewriter.rs:43-52 with N = 2 */
  size_t residual = w->p_byte % 2;
  if (residual == 0) return X3O_OK;
  uint8_t z = 0;
  return sw_write_all(w, &z, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* bitpacker.rs:46-177 BitPacker                                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  slice_writer *writer;
  uint8_t scratch_byte;
  size_t p_bit;
  size_t byte_len;
  uint16_t crc;
} bit_packer;

static void bp_new(bit_packer *bp, slice_writer *w) { /* bitpacker.rs:64-72 */
  bp->writer = w; bp->scratch_byte = 0; bp->p_bit = 0; bp->byte_len = 0; bp->crc = 0xffff;
}

static int bp_flush(bit_packer *bp) { /* bitpacker.rs:79-86 */
  bp->crc = x3o_update_crc16(bp->crc, bp->scratch_byte);
  bp->byte_len += 1;
  int r = sw_write_all(bp->writer, &bp->scratch_byte, 1);
  if (r) return r;
  bp->scratch_byte = 0;
  bp->p_bit = 0;
  return X3O_OK;
}

static int bp_write_bits(bit_packer *bp, uint64_t value, size_t num_bits) { /* bitpacker.rs:142-163 */
  size_t rem_bit = 8 - bp->p_bit;
  uint64_t mask = (num_bits >= 64) ? ~0ull : ((1ull << num_bits) - 1ull);
  value &= mask;
  int r;
  if (num_bits == rem_bit) {
    bp->scratch_byte |= (uint8_t)value;
    if ((r = bp_flush(bp))) return r;
  } else if (num_bits < rem_bit) {
    size_t shift_l = rem_bit - num_bits;
    bp->scratch_byte |= (uint8_t)(value << shift_l);
    bp->p_bit += num_bits;
  } else {
    size_t shift_r = num_bits - rem_bit;
    bp->scratch_byte |= (uint8_t)(value >> shift_r);
    if ((r = bp_flush(bp))) return r;
    if ((r = bp_write_bits(bp, value, shift_r))) return r;
  }
  return X3O_OK;
}

static int bp_word_align(bit_packer *bp) { /* bitpacker.rs:124-132 */
  int r;
  if (bp->p_bit != 0)
    if ((r = bp_flush(bp))) return r;
  while (0 != (bp->writer->p_byte % 2))
    if ((r = bp_flush(bp))) return r;
  return X3O_OK;
}

static int bp_drop(bit_packer *bp) { /* bitpacker.rs:55-61 Drop */
  if (bp->p_bit != 0) return bp_flush(bp);
  return X3O_OK;
}

int x3o_bitpack(const uint64_t *values, const uint32_t *num_bits, size_t n, uint8_t *buf,
                size_t cap, size_t *out_len) {
  slice_writer w = {buf, cap, 0, 0};
  bit_packer bp;
  bp_new(&bp, &w);
  int r;
  for (size_t i = 0; i < n; i++)
    if ((r = bp_write_bits(&bp, values[i], num_bits[i]))) return r;
  if ((r = bp_drop(&bp))) return r;
  *out_len = w.p_byte;
  return X3O_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* encoder.rs                                                                                  */
/* ------------------------------------------------------------------------------------------ */

static uint32_t count_bits(uint32_t n) { /* encoder.rs:229-231 */
  return n == 0 ? 0u : 32u - (uint32_t)__builtin_clz(n);
}

void x3o_write_frame_header(size_t num_samples, uint8_t id, size_t payload_len,
                            uint16_t payload_crc, uint8_t header[20]) { /* encoder.rs:122-162 */
  memset(header, 0, FRAME_HEADER_LEN);
  size_t p = 0;
  header[p] = (uint8_t)(FRAME_KEY >> 8); header[p + 1] = (uint8_t)(FRAME_KEY & 0xff); p += 2;
  header[p] = id; p += 1;                /* <Source Id> */
  header[p] = id; p += 1;                /* <Num Channels> written as id, encoder.rs:135 */
  uint16_t ns = (uint16_t)num_samples;   /* `as u16` encoder.rs:141 */
  header[p] = (uint8_t)(ns >> 8); header[p + 1] = (uint8_t)(ns & 0xff); p += 2;
  uint16_t pl = (uint16_t)payload_len;   /* encoder.rs:145 */
  header[p] = (uint8_t)(pl >> 8); header[p + 1] = (uint8_t)(pl & 0xff); p += 2;
  p += 8;                                /* <Time> zero, encoder.rs:148-150 */
  uint16_t hc = x3o_crc16(header, 16);   /* encoder.rs:153 */
  header[p] = (uint8_t)(hc >> 8); header[p + 1] = (uint8_t)(hc & 0xff); p += 2;
  header[p] = (uint8_t)(payload_crc >> 8); header[p + 1] = (uint8_t)(payload_crc & 0xff);
}

static int encode_rice_block(const int32_t *wav_diff, size_t n, bit_packer *bp, const x3o_params *p,
                             int32_t max_abs, size_t *stat_idx) { /* encoder.rs:233-267 */
  size_t ftype = 0;
  for (int t = 0; t < 3; t++)
    if (max_abs > (int32_t)p->thresholds[t]) ftype += 1;
  int r;
  if ((r = bp_write_bits(bp, ftype + 1, 2))) return r;
  const rice_code *rc = &RICE[p->codes[ftype]];
  for (size_t i = 0; i < n; i++) {
    int64_t ii = (int64_t)wav_diff[i] + (int64_t)rc->offset;
    if (ii < 0 || (uint64_t)ii >= rc->n) return X3O_ERR_REFERENCE_PANIC; /* slice index panic */
    uint32_t code = rc->code[ii];
    size_t rc_num_bits = rc->num_bits[ii];
    size_t num_zeros = rc_num_bits - count_bits(code);
    if ((r = bp_write_bits(bp, 0, num_zeros))) return r;          /* write_packed_zeros */
    if ((r = bp_write_bits(bp, code, rc_num_bits - num_zeros))) return r;
  }
  *stat_idx = rc->nsubs;
  return X3O_OK;
}

static int encode_bfp_block(const int32_t *wav_diff, size_t n, bit_packer *bp, size_t num_bits,
                            size_t *stat_idx) { /* encoder.rs:269-276 */
  int r;
  if ((r = bp_write_bits(bp, num_bits, 6))) return r;
  for (size_t i = 0; i < n; i++)
    if ((r = bp_write_bits(bp, (uint64_t)(int64_t)wav_diff[i], num_bits + 1))) return r;
  *stat_idx = 4;
  return X3O_OK;
}

static int encode_literal(const int16_t *wav, size_t n, bit_packer *bp, size_t *stat_idx) { /* :278-285 */
  int r;
  if ((r = bp_write_bits(bp, 15, 6))) return r;
  for (size_t i = 0; i < n; i++)
    if ((r = bp_write_bits(bp, (uint64_t)(int64_t)wav[i], 16))) return r;
  *stat_idx = 5;
  return X3O_OK;
}

/* encoder.rs:289-315.  `wav` = the block's samples, `prev` = the sample before the block (the diff
 * iterator of encoder.rs:192 yields wav[i]-wav[i-1] across block boundaries within the frame). */
static int x3_encode_block(const int16_t *wav, size_t n, int16_t prev, bit_packer *bp,
                           const x3o_params *p, size_t *stat_idx) {
  if (n > MAX_BLOCK_LENGTH) return X3O_ERR_REFERENCE_PANIC; /* wav_diff[i] on [0i32;60] */
  int32_t wav_diff[MAX_BLOCK_LENGTH];
  int32_t max_abs = 0;
  int32_t last = prev;
  for (size_t i = 0; i < n; i++) {
    int32_t wd = (int32_t)wav[i] - last; /* encoder.rs:224 */
    last = wav[i];
    wav_diff[i] = wd;
    int32_t a = wd < 0 ? -wd : wd;
    if (a > max_abs) max_abs = a;
  }
  if (max_abs <= (int32_t)p->thresholds[2]) {
    return encode_rice_block(wav_diff, n, bp, p, max_abs, stat_idx);
  } else {
    size_t num_bits = count_bits((uint32_t)max_abs);
    if (num_bits >= 15) return encode_literal(wav, n, bp, stat_idx);
    return encode_bfp_block(wav_diff, n, bp, num_bits, stat_idx);
  }
}

int x3o_encode_block_test(const int16_t *wav, size_t n, const x3o_params *p, uint32_t lead_zero_bits,
                          uint8_t *buf, size_t cap, size_t *out_len) {
  build_tables();
  slice_writer w = {buf, cap, 0, 0};
  bit_packer bp;
  bp_new(&bp, &w);
  int r;
  if (lead_zero_bits && (r = bp_write_bits(&bp, 0, lead_zero_bits))) return r;
  size_t st;
  if ((r = x3_encode_block(wav + 1, n - 1, wav[0], &bp, p, &st))) return r;
  if ((r = bp_word_align(&bp))) return r;
  *out_len = bp.byte_len;
  return X3O_OK;
}

static int encode_frame_w(const int16_t *wav, size_t n, slice_writer *w, const x3o_params *p,
                          uint64_t stats[6]) { /* encoder.rs:175-214 */
  int r;
  if ((r = sw_align2(w))) return r;                              /* :182 */
  size_t frame_header_pos = w->p_byte;                           /* :183 */
  if ((r = sw_seek_abs(w, w->p_byte + FRAME_HEADER_LEN))) return r; /* :184 */
  bit_packer bp;
  bp_new(&bp, w);
  if ((r = bp_write_bits(&bp, (uint64_t)(int64_t)wav[0], 16))) return r; /* :189 */
  if (p->block_len == 0) return X3O_ERR_REFERENCE_PANIC;         /* chunks(0) panics */
  for (size_t s = 1; s < n; s += p->block_len) {                 /* :194 wav[1..].chunks(block_len) */
    size_t bl = n - s < p->block_len ? n - s : p->block_len;
    size_t st = 0;
    if ((r = x3_encode_block(wav + s, bl, wav[s - 1], &bp, p, &st))) return r;
    stats[st] += bl;                                             /* :199 */
  }
  if ((r = bp_word_align(&bp))) return r;                        /* :203 */
  size_t payload_len = bp.byte_len;
  uint16_t payload_crc = bp.crc;
  size_t return_position = w->p_byte;                            /* :208 */
  if ((r = sw_seek_abs(w, frame_header_pos))) return r;
  uint8_t hdr[FRAME_HEADER_LEN];
  x3o_write_frame_header(n, 1, payload_len, payload_crc, hdr);   /* :210 */
  if ((r = sw_write_all(w, hdr, FRAME_HEADER_LEN))) return r;
  return sw_seek_abs(w, return_position);
}

int x3o_encode_frame(const int16_t *wav, size_t n, const x3o_params *p, uint8_t *buf, size_t cap,
                     size_t *pos, uint64_t stats[6]) {
  build_tables();
  if (n == 0) return X3O_ERR_REFERENCE_PANIC; /* wav[0] */
  slice_writer w = {buf, cap, *pos, *pos};
  int r = encode_frame_w(wav, n, &w, p, stats);
  *pos = w.p_byte;
  return r;
}

int x3o_encode(const int16_t *wav, size_t n, const x3o_params *p, uint8_t *buf, size_t cap,
               size_t *pos, uint64_t stats[6]) { /* encoder.rs:51-111 */
  build_tables();
  size_t samples_per_frame = (size_t)p->block_len * p->blocks_per_frame; /* :61 */
  slice_writer w = {buf, cap, *pos, *pos};
  size_t off = 0;
  for (;;) {                                                    /* :67-73 */
    size_t take = n - off < samples_per_frame ? n - off : samples_per_frame;
    if (take == 0) break;
    int r = encode_frame_w(wav + off, take, &w, p, stats);
    if (r) { *pos = w.p_byte; return r; }
    off += take;
  }
  *pos = w.p_byte;
  return X3O_OK;
}

size_t x3o_encode_bound(size_t n, const x3o_params *p) {
  size_t spf = (size_t)p->block_len * p->blocks_per_frame;
  if (spf == 0) return 0;
  size_t frames = (n + spf - 1) / spf;
  size_t blocks = (n + p->block_len - 1) / p->block_len + frames;
  /* per frame: 20 header + 2 first sample + 1 align pad; per block 6 bits; per sample <= 17 bits
   * (BFP nb=14 -> 15 bits, literal 16, rice <= 2*t2>>k+k+1 bits is covered by validate for defaults) */
  return frames * 24 + blocks + (n * 17 + 7) / 8 + 16;
}

/* ------------------------------------------------------------------------------------------ */
/* bitreader.rs:29-176                                                                         */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  const uint8_t *array;
  size_t len;
  size_t idx;
  uint32_t leading_word;
  size_t rem_bit;
} bit_reader;

static void read_word(const uint8_t *a, size_t len, size_t idx, uint32_t *word, size_t *nbytes) {
  /* bitreader.rs:29-48 */
  if (len - idx >= 4) {
    *word = ((uint32_t)a[idx] << 24) | ((uint32_t)a[idx + 1] << 16) | ((uint32_t)a[idx + 2] << 8) | a[idx + 3];
    *nbytes = 4;
  } else {
    size_t remaining_idx = len - idx;
    uint32_t w = 0;
    if (remaining_idx >= 1) w |= (uint32_t)a[idx] << 24;
    if (remaining_idx >= 2) w |= (uint32_t)a[idx + 1] << 16;
    if (remaining_idx == 3) w |= (uint32_t)a[idx + 2] << 8;
    *word = w;
    *nbytes = remaining_idx;
  }
}

static void br_new(bit_reader *br, const uint8_t *a, size_t len) { /* bitreader.rs:65-74 */
  br->array = a; br->len = len;
  uint32_t w; size_t nb;
  read_word(a, len, 0, &w, &nb);
  br->idx = nb; br->leading_word = w; br->rem_bit = nb * 8;
}

static int br_peek_next(const bit_reader *br, uint32_t *word, size_t *nbytes) { /* :169-175 */
  if (br->idx >= br->len) return 0;
  read_word(br->array, br->len, br->idx, word, nbytes);
  return 1;
}

static void br_get_next(bit_reader *br) { /* bitreader.rs:149-164 */
  uint32_t w; size_t nb;
  if (br_peek_next(br, &w, &nb)) {
    br->leading_word = w; br->idx += nb; br->rem_bit = nb * 8;
  } else {
    br->leading_word = 0; br->rem_bit = 0;
  }
}

static uint32_t shl32(uint32_t v, size_t n) { return n >= 32 ? 0u : v << n; } /* release-mode wrapping is
  never relied on by valid streams; n==32 only arises from count_zero_bits()==32, where the reference
  (release) computes v << (32 & 31) = v; that path is only reachable when leading_word == 0 so both give 0. */

static void br_inc_bits(bit_reader *br, size_t n) { /* bitreader.rs:77-92 */
  if (n < br->rem_bit) {
    br->leading_word = shl32(br->leading_word, n);
    br->rem_bit -= n;
  } else if (n > br->rem_bit) {
    size_t rem = n - br->rem_bit;
    br_get_next(br);
    br->rem_bit = 32 - rem;
    br->leading_word = shl32(br->leading_word, rem);
  } else {
    br_get_next(br);
  }
}

static uint32_t br_read_nbits(bit_reader *br, size_t n) { /* bitreader.rs:105-118 */
  if (n <= br->rem_bit) {
    uint32_t result = br->leading_word >> (32 - n);
    br_inc_bits(br, n);
    return result;
  } else {
    size_t rem = n - br->rem_bit;
    uint32_t result = br->leading_word >> (32 - n);
    br_inc_bits(br, br->rem_bit);
    result |= br->leading_word >> (32 - rem);
    br_inc_bits(br, rem);
    return result;
  }
}

static size_t br_count_zero_bits(bit_reader *br) { /* bitreader.rs:127-139 */
  size_t count = br->leading_word == 0 ? 32 : (size_t)__builtin_clz(br->leading_word);
  if (count > br->rem_bit) {
    uint32_t w; size_t nb;
    if (br_peek_next(br, &w, &nb))
      count = br->rem_bit + (w == 0 ? 32 : (size_t)__builtin_clz(w));
    else
      count = br->rem_bit;
  }
  br_inc_bits(br, count);
  return count;
}

int x3o_bitread(const uint8_t *buf, size_t len, const uint32_t *ops, size_t n_ops,
                uint32_t *results, uint32_t *lead, uint32_t *rem) {
  bit_reader br;
  br_new(&br, buf, len);
  for (size_t i = 0; i < n_ops; i++) {
    if (ops[i] == 0) results[i] = (uint32_t)br_count_zero_bits(&br);
    else if (ops[i] == 0xffffffffu) results[i] = 0; /* no-op: report initial state */
    else results[i] = br_read_nbits(&br, ops[i]);
    lead[i] = br.leading_word;
    rem[i] = (uint32_t)br.rem_bit;
  }
  return X3O_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* decoder.rs                                                                                  */
/* ------------------------------------------------------------------------------------------ */

int x3o_read_frame_header(const uint8_t *bytes, size_t len, x3o_frame_header *h) { /* decoder.rs:69-118 */
  if (len < FRAME_HEADER_LEN) return X3O_ERR_FRAME_DECODE_UNEXPECTED_END;
  uint16_t header_crc = x3o_crc16(bytes, 16);
  uint16_t expected = (uint16_t)((bytes[16] << 8) | bytes[17]);
  if (expected != header_crc) return X3O_ERR_FRAME_HEADER_INVALID_HEADER_CRC;
  uint16_t key = (uint16_t)((bytes[0] << 8) | bytes[1]);
  if (key != FRAME_KEY) return X3O_ERR_FRAME_HEADER_INVALID_KEY;
  h->source_id = bytes[2];
  h->channels = bytes[3];
  if (h->channels > 1) return X3O_ERR_MORE_THAN_ONE_CHANNEL;
  h->samples = (uint16_t)((bytes[4] << 8) | bytes[5]);
  h->payload_len = (uint32_t)((bytes[6] << 8) | bytes[7]);
  if (h->payload_len >= FRAME_MAX_LENGTH) return X3O_ERR_FRAME_LENGTH;
  h->payload_crc = (uint16_t)((bytes[18] << 8) | bytes[19]);
  return X3O_OK;
}

static int decode_ricecode_block_r1(bit_reader *br, int16_t *wav, size_t n, int16_t *last_wav,
                                    const x3o_params *p, size_t ftype) { /* decoder.rs:147-170 */
  const rice_code *code = &RICE[p->codes[ftype - 1]];
  int16_t lw = *last_wav;
  for (size_t b = 0; b < n; b++) {
    size_t i = br_count_zero_bits(br);
    br_read_nbits(br, 1);
    if (i >= code->inv_len) return X3O_ERR_OUT_OF_BOUNDS_INVERSE;
    lw = (int16_t)(lw + INV_RICE_CODE[i]);
    wav[b] = lw;
  }
  *last_wav = lw;
  return X3O_OK;
}

static int decode_ricecode_block_r2r3(bit_reader *br, int16_t *wav, size_t n, int16_t *last_wav,
                                      const x3o_params *p, size_t ftype) { /* decoder.rs:172-196 */
  const rice_code *code = &RICE[p->codes[ftype - 1]];
  size_t nb = ftype == 2 ? 2 : 4;
  int16_t level = (int16_t)(1 << code->nsubs);
  int16_t lw = *last_wav;
  for (size_t b = 0; b < n; b++) {
    int16_t nn = (int16_t)br_count_zero_bits(br);
    int16_t r = (int16_t)br_read_nbits(br, nb);
    int16_t iv = (int16_t)(r + (int16_t)(level * (int16_t)(nn - 1)));
    size_t i = (size_t)(int64_t)iv; /* `as usize` sign-extends a negative i16 to a huge index */
    if (i >= code->inv_len) return X3O_ERR_OUT_OF_BOUNDS_INVERSE;
    lw = (int16_t)(lw + INV_RICE_CODE[i]);
    wav[b] = lw;
  }
  *last_wav = lw;
  return X3O_OK;
}

static int16_t unsigned_to_i16(uint16_t a16, size_t num_bits) { /* decoder.rs:198-207 */
  int32_t a = (int32_t)a16;
  int32_t neg_thresh = 1 << (num_bits - 1);
  int32_t neg = 1 << num_bits;
  if (a > neg_thresh) a -= neg;
  return (int16_t)a;
}

static int decode_bpf_block(bit_reader *br, int16_t *wav, size_t n, int16_t *last_wav) { /* :209-235 */
  size_t num_bits = (size_t)br_read_nbits(br, 4) + 1;
  if (num_bits <= 5) return X3O_ERR_FRAME_DECODE_INVALID_BPF;
  if (num_bits == 16) {
    for (size_t i = 0; i < n; i++) wav[i] = (int16_t)br_read_nbits(br, 16);
  } else {
    int16_t value = *last_wav;
    for (size_t i = 0; i < n; i++) {
      uint16_t diff = (uint16_t)br_read_nbits(br, num_bits);
      value = (int16_t)(value + unsigned_to_i16(diff, num_bits));
      wav[i] = value;
    }
  }
  *last_wav = wav[n - 1];
  return X3O_OK;
}

static int decode_block(bit_reader *br, int16_t *wav, size_t n, int16_t *last_wav,
                        const x3o_params *p) { /* decoder.rs:132-145 */
  size_t ftype = br_read_nbits(br, 2);
  switch (ftype) {
    case 0: return decode_bpf_block(br, wav, n, last_wav);
    case 1: return decode_ricecode_block_r1(br, wav, n, last_wav, p, ftype);
    case 2: case 3: return decode_ricecode_block_r2r3(br, wav, n, last_wav, p, ftype);
    default: return X3O_ERR_FRAME_DECODE_INVALID_FTYPE;
  }
}

int x3o_decode_block_test(const uint8_t *buf, size_t len, uint32_t skip_bits, int16_t last_wav,
                          const x3o_params *p, int16_t *wav, size_t block_len) {
  build_tables();
  bit_reader br;
  br_new(&br, buf, len);
  if (skip_bits) br_read_nbits(&br, skip_bits);
  int16_t lw = last_wav;
  return decode_block(&br, wav, block_len, &lw, p);
}

int x3o_decode_frame(const uint8_t *payload, size_t payload_len, int16_t *wav, size_t wav_cap,
                     const x3o_params *p, size_t samples, size_t *n_out) { /* decoder.rs:36-58 */
  build_tables();
  *n_out = 0;
  if (payload_len < 2 || samples == 0 || samples > wav_cap || p->block_len == 0)
    return X3O_ERR_REFERENCE_PANIC; /* slice index / usize underflow panics */
  int16_t last_wav = (int16_t)((payload[0] << 8) | payload[1]);
  size_t p_wav = 0;
  wav[p_wav++] = last_wav;
  bit_reader br;
  br_new(&br, payload + 2, payload_len - 2);
  size_t remaining = samples - 1;
  while (remaining > 0) {
    size_t block_len = remaining < p->block_len ? remaining : p->block_len;
    int r = decode_block(&br, wav + p_wav, block_len, &last_wav, p);
    if (r) return r;
    remaining -= block_len;
    p_wav += block_len;
  }
  *n_out = p_wav;
  return X3O_OK;
}

int x3o_decode_stream(const uint8_t *bytes, size_t len, size_t remaining0, const x3o_params *p,
                      int16_t *wav, size_t wav_cap, size_t *n_out, size_t *frames_ok,
                      size_t *frame_errors) {
  /* decodefile.rs:105-136 driven by the loop of :202-209.  `cursor` is the BufReader position. */
  build_tables();
  size_t cursor = 0, remaining = remaining0, produced = 0, frames = 0;
  *frame_errors = 0;
  static __thread int16_t frame_buf[X3_WRITE_BUFFER_SIZE];
  int rc = X3O_OK;
  for (;;) {
    if (remaining <= FRAME_HEADER_LEN) break;                        /* :107-109 */
    /* read_bytes(20): clamps to remaining (not needed here), read_exact -> Io error at EOF */
    if (cursor + FRAME_HEADER_LEN > len) { rc = X3O_ERR_IO; break; }
    remaining -= FRAME_HEADER_LEN;
    x3o_frame_header h;
    rc = x3o_read_frame_header(bytes + cursor, FRAME_HEADER_LEN, &h); /* :112 */
    cursor += FRAME_HEADER_LEN;
    if (rc) break;
    size_t samples = h.samples;
    if (remaining < h.payload_len) break;                            /* :114-116 Ok(None) */
    if (h.payload_len > X3_READ_BUFFER_SIZE) { rc = X3O_ERR_FRAME_HEADER_INVALID_PAYLOAD_LEN; break; }
    if (cursor + h.payload_len > len) { rc = X3O_ERR_IO; break; }    /* read_exact */
    remaining -= h.payload_len;
    const uint8_t *payload = bytes + cursor;
    cursor += h.payload_len;
    if (x3o_crc16(payload, h.payload_len) != h.payload_crc) {        /* :97-100 */
      rc = X3O_ERR_FRAME_HEADER_INVALID_PAYLOAD_CRC; break;
    }
    size_t got = 0;
    int dr = x3o_decode_frame(payload, h.payload_len, frame_buf, X3_WRITE_BUFFER_SIZE, p, samples, &got);
    if (dr == X3O_ERR_REFERENCE_PANIC) { rc = dr; break; }
    if (dr) { *frame_errors += 1; break; }                           /* :130-134 Ok(None) */
    if (produced + got > wav_cap) { rc = X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY; break; }
    memcpy(wav + produced, frame_buf, got * sizeof(int16_t));        /* write_samples :205 */
    produced += got;
    frames += 1;
  }
  *n_out = produced;
  *frames_ok = frames;
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* encodefile.rs:82-138 / decodefile.rs:142-303: archive header                                */
/* ------------------------------------------------------------------------------------------ */

static const uint8_t ARCHIVE_ID[8] = {0x58, 0x33, 0x41, 0x52, 0x43, 0x48, 0x49, 0x56}; /* x3.rs:139 */

int x3o_archive_header(uint32_t sample_rate, const x3o_params *p, uint8_t *buf, size_t cap,
                       size_t *out_len) {
  char xml[1024];
  int n = snprintf(xml, sizeof xml,
                   "<X3ARCH PROG=\"x3new.m\" VERSION=\"2.0\" />"
                   "<CFG ID=\"0\" FTYPE=\"XML\" />"
                   "<CFG ID=\"1\" FTYPE=\"WAV\">"
                   "<FS UNIT=\"Hz\">%u</FS>"
                   "<SUFFIX>wav</SUFFIX>"
                   "<CODEC TYPE=\"X3\" VERS=\"2\">"
                   "<BLKLEN>%u</BLKLEN>"
                   "<CODES N=\"4\">RICE%u,RICE%u,RICE%u,BFP</CODES>"
                   "<FILTER>DIFF</FILTER>"
                   "<NBITS>16</NBITS>"
                   "<T N=\"3\">%u,%u,%u</T>"
                   "</CODEC>"
                   "</CFG>",
                   sample_rate, p->block_len, p->codes[0], p->codes[1], p->codes[2],
                   p->thresholds[0], p->thresholds[1], p->thresholds[2]); /* encodefile.rs:93-117 */
  size_t payload_len = (size_t)n;
  uint16_t payload_crc = x3o_crc16((const uint8_t *)xml, payload_len);
  size_t padded = payload_len + (payload_len % 2);
  if (8 + FRAME_HEADER_LEN + padded > cap) return X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY;
  memcpy(buf, ARCHIVE_ID, 8);                                       /* :87 */
  memcpy(buf + 8 + FRAME_HEADER_LEN, xml, payload_len);             /* :123 */
  if (payload_len % 2 == 1) {                                       /* :124-129 */
    buf[8 + FRAME_HEADER_LEN + payload_len] = 0;
    payload_len += 1;
    payload_crc = x3o_update_crc16(payload_crc, 0);
  }
  x3o_write_frame_header(0, 0, payload_len, payload_crc, buf + 8);  /* :135 */
  *out_len = 8 + FRAME_HEADER_LEN + payload_len;
  return X3O_OK;
}

/* minimal scanner for <TAG ...>text</TAG> as used by decodefile.rs:232-303 (first occurrence) */
static int xml_text(const char *xml, size_t len, const char *tag, char *out, size_t out_cap) {
  size_t tl = strlen(tag);
  for (size_t i = 0; i + tl + 1 < len; i++) {
    if (xml[i] == '<' && !memcmp(xml + i + 1, tag, tl) && (xml[i + 1 + tl] == '>' || xml[i + 1 + tl] == ' ')) {
      size_t j = i + 1 + tl;
      while (j < len && xml[j] != '>') j++;
      if (j >= len) return -1;
      if (xml[j - 1] == '/') continue; /* empty element */
      size_t s = j + 1, e = s;
      while (e < len && xml[e] != '<') e++;
      while (s < e && (xml[s] == ' ' || xml[s] == '\n' || xml[s] == '\t' || xml[s] == '\r')) s++; /* trim_text */
      while (e > s && (xml[e - 1] == ' ' || xml[e - 1] == '\n' || xml[e - 1] == '\t' || xml[e - 1] == '\r')) e--;
      size_t n = e - s;
      if (n + 1 > out_cap) return -1;
      memcpy(out, xml + s, n);
      out[n] = 0;
      return 0;
    }
  }
  return -1;
}

int x3o_archive_parse(const uint8_t *bytes, size_t len, uint32_t *sample_rate, x3o_params *p,
                      size_t *header_size, size_t *frames_offset) {
  if (len < 8) return X3O_ERR_IO;
  if (memcmp(bytes, ARCHIVE_ID, 8)) return X3O_ERR_ARCHIVE_INVALID_KEY;  /* decodefile.rs:147-149 */
  if (len < 8 + FRAME_HEADER_LEN) return X3O_ERR_IO;
  x3o_frame_header h;
  int r = x3o_read_frame_header(bytes + 8, FRAME_HEADER_LEN, &h);         /* :154-157 */
  if (r) return r;
  if (len < 8 + FRAME_HEADER_LEN + (size_t)h.payload_len) return X3O_ERR_IO;
  const char *xml = (const char *)bytes + 8 + FRAME_HEADER_LEN;
  char fs[64], bl[64], codes[128], th[128];
  if (xml_text(xml, h.payload_len, "FS", fs, sizeof fs) || xml_text(xml, h.payload_len, "BLKLEN", bl, sizeof bl) ||
      xml_text(xml, h.payload_len, "CODES", codes, sizeof codes) || xml_text(xml, h.payload_len, "T", th, sizeof th))
    return X3O_ERR_REFERENCE_PANIC; /* fs[0] etc. index panic when a tag is missing */
  *sample_rate = (uint32_t)strtoul(fs, NULL, 10);
  p->block_len = (uint32_t)strtoul(bl, NULL, 10);
  p->blocks_per_frame = 500; /* decodefile.rs:297 DEFAULT_BLOCKS_PER_FRAME */
  uint32_t ids[8]; size_t nid = 0;
  {
    char *save = NULL;
    for (char *w = strtok_r(codes, ",", &save); w; w = strtok_r(NULL, ",", &save)) { /* :276-285 */
      if (!strcmp(w, "RICE0")) { if (nid < 8) ids[nid++] = 0; }
      else if (!strcmp(w, "RICE1")) { if (nid < 8) ids[nid++] = 1; }
      else if (!strcmp(w, "RICE2")) { if (nid < 8) ids[nid++] = 2; }
      else if (!strcmp(w, "RICE3")) { if (nid < 8) ids[nid++] = 3; }
      else if (!strcmp(w, "BFP")) {}
      else return X3O_ERR_ARCHIVE_XML_RICE_CODE;
    }
  }
  uint32_t ths[8]; size_t nth = 0;
  {
    char *save = NULL;
    for (char *w = strtok_r(th, ",", &save); w; w = strtok_r(NULL, ",", &save))
      if (nth < 8) ths[nth++] = (uint32_t)strtoul(w, NULL, 10);
  }
  if (nid < 3 || nth < 3) return X3O_ERR_REFERENCE_PANIC;
  for (int i = 0; i < 3; i++) { p->codes[i] = ids[i]; p->thresholds[i] = ths[i]; }
  r = x3o_params_validate(p);
  if (r) return r;
  *header_size = FRAME_HEADER_LEN + h.payload_len; /* decodefile.rs:166: id length NOT included */
  *frames_offset = 8 + FRAME_HEADER_LEN + h.payload_len;
  return X3O_OK;
}

int x3o_x3a_encode(const int16_t *wav, size_t n, uint32_t sample_rate, uint8_t *buf, size_t cap,
                   size_t *out_len, uint64_t stats[6]) { /* encodefile.rs:48-78 */
  x3o_params p;
  x3o_params_default(&p);                                            /* :57 */
  size_t pos = 0;
  int r = x3o_archive_header(sample_rate, &p, buf, cap, &pos);       /* :72 */
  if (r) return r;
  r = x3o_encode(wav, n, &p, buf, cap, &pos, stats);                 /* :74 */
  *out_len = pos;
  return r;
}

int x3o_x3a_decode(const uint8_t *bytes, size_t len, int16_t *wav, size_t wav_cap, size_t *n_out,
                   uint32_t *sample_rate, size_t *frames_ok, size_t *frame_errors) {
  x3o_params p;
  size_t header_size = 0, off = 0;
  *n_out = 0; *frames_ok = 0; *frame_errors = 0;
  int r = x3o_archive_parse(bytes, len, sample_rate, &p, &header_size, &off);
  if (r) return r;
  size_t remaining0 = len - header_size; /* decodefile.rs:61-65: file length minus (20 + xml) only */
  return x3o_decode_stream(bytes + off, len - off, remaining0, &p, wav, wav_cap, n_out, frames_ok,
                           frame_errors);
}

/* ------------------------------------------------------------------------------------------ */
/* frame-parallel multi-threaded variants (CPU baseline on all host cores)                     */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  const int16_t *wav; size_t n; const x3o_params *p;
  size_t f0, f1, spf;
  uint8_t *tmp; size_t tmp_cap; size_t tmp_len;
  uint64_t stats[6];
  int rc;
} enc_job;

static void *enc_worker(void *arg) {
  enc_job *j = (enc_job *)arg;
  size_t pos = 0;
  size_t s0 = j->f0 * j->spf, s1 = j->f1 * j->spf;
  if (s1 > j->n) s1 = j->n;
  j->rc = x3o_encode(j->wav + s0, s1 - s0, j->p, j->tmp, j->tmp_cap, &pos, j->stats);
  j->tmp_len = pos;
  return NULL;
}

int x3o_encode_mt(const int16_t *wav, size_t n, const x3o_params *p, uint8_t *buf, size_t cap,
                  size_t *out_len, uint64_t stats[6], int threads) {
  build_tables();
  size_t spf = (size_t)p->block_len * p->blocks_per_frame;
  if (spf == 0) return X3O_ERR_REFERENCE_PANIC;
  size_t frames = (n + spf - 1) / spf;
  if (threads < 1) threads = 1;
  if ((size_t)threads > frames) threads = frames ? (int)frames : 1;
  enc_job *jobs = (enc_job *)calloc((size_t)threads, sizeof(enc_job));
  pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; t++) {
    enc_job *j = &jobs[t];
    j->wav = wav; j->n = n; j->p = p; j->spf = spf;
    j->f0 = frames * (size_t)t / (size_t)threads;
    j->f1 = frames * (size_t)(t + 1) / (size_t)threads;
    size_t ns = (j->f1 - j->f0) * spf;
    j->tmp_cap = x3o_encode_bound(ns, p);
    j->tmp = (uint8_t *)malloc(j->tmp_cap ? j->tmp_cap : 1);
    pthread_create(&th[t], NULL, enc_worker, j);
  }
  int rc = X3O_OK;
  size_t pos = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    enc_job *j = &jobs[t];
    if (j->rc && !rc) rc = j->rc;
    if (!rc) {
      if (pos + j->tmp_len > cap) rc = X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY;
      else { memcpy(buf + pos, j->tmp, j->tmp_len); pos += j->tmp_len; }
    }
    for (int k = 0; k < 6; k++) stats[k] += j->stats[k];
    free(j->tmp);
  }
  free(jobs); free(th);
  *out_len = pos;
  return rc;
}

typedef struct {
  const uint8_t *bytes; const x3o_params *p; int16_t *wav;
  const size_t *fpos; const size_t *spos; size_t f0, f1;
  int rc;
} dec_job;

static void *dec_worker(void *arg) {
  dec_job *j = (dec_job *)arg;
  j->rc = X3O_OK;
  for (size_t f = j->f0; f < j->f1; f++) {
    const uint8_t *hb = j->bytes + j->fpos[f];
    x3o_frame_header h;
    int r = x3o_read_frame_header(hb, FRAME_HEADER_LEN, &h);
    if (r) { j->rc = r; return NULL; }
    if (x3o_crc16(hb + FRAME_HEADER_LEN, h.payload_len) != h.payload_crc) {
      j->rc = X3O_ERR_FRAME_HEADER_INVALID_PAYLOAD_CRC; return NULL;
    }
    size_t got;
    r = x3o_decode_frame(hb + FRAME_HEADER_LEN, h.payload_len, j->wav + j->spos[f], h.samples, j->p,
                         h.samples, &got);
    if (r) { j->rc = r; return NULL; }
  }
  return NULL;
}

int x3o_decode_stream_mt(const uint8_t *bytes, size_t len, const x3o_params *p, int16_t *wav,
                         size_t wav_cap, size_t *n_out, int threads) {
  /* well-formed streams only: serial header walk (cheap), then frame-parallel payload decode */
  build_tables();
  size_t cap = 1024, nf = 0, cursor = 0, total = 0;
  size_t *fpos = (size_t *)malloc(cap * sizeof(size_t)), *spos = (size_t *)malloc(cap * sizeof(size_t));
  while (len - cursor > FRAME_HEADER_LEN) {
    x3o_frame_header h;
    int r = x3o_read_frame_header(bytes + cursor, FRAME_HEADER_LEN, &h);
    if (r) { free(fpos); free(spos); return r; }
    if (cursor + FRAME_HEADER_LEN + h.payload_len > len) break;
    if (total + h.samples > wav_cap) { free(fpos); free(spos); return X3O_ERR_BYTEWRITER_INSUFFICIENT_MEMORY; }
    if (nf == cap) { cap *= 2; fpos = (size_t *)realloc(fpos, cap * sizeof(size_t)); spos = (size_t *)realloc(spos, cap * sizeof(size_t)); }
    fpos[nf] = cursor; spos[nf] = total; nf++;
    total += h.samples;
    cursor += FRAME_HEADER_LEN + h.payload_len;
  }
  if (threads < 1) threads = 1;
  dec_job *jobs = (dec_job *)calloc((size_t)threads, sizeof(dec_job));
  pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; t++) {
    dec_job *j = &jobs[t];
    j->bytes = bytes; j->p = p; j->wav = wav; j->fpos = fpos; j->spos = spos;
    j->f0 = nf * (size_t)t / (size_t)threads; j->f1 = nf * (size_t)(t + 1) / (size_t)threads;
    pthread_create(&th[t], NULL, dec_worker, j);
  }
  int rc = X3O_OK;
  for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); if (jobs[t].rc && !rc) rc = jobs[t].rc; }
  free(jobs); free(th); free(fpos); free(spos);
  *n_out = total;
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Synthetic signals, SURVEY.md section 8(d).  Integer only; sample n is a pure function of     */
/* (kind, seed, fs, n).                                                                        */
/* ------------------------------------------------------------------------------------------ */

static uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint64_t hh(uint32_t seed, uint64_t n) { return splitmix64(((uint64_t)seed << 32) ^ n); }
static int32_t uni(uint32_t seed, uint64_t n, int32_t a) {
  return (int32_t)(hh(seed, n) % (uint64_t)(2 * a + 1)) - a;
}
static int32_t colored(uint32_t seed, uint64_t n, int32_t a) {
  int32_t s = 0;
  for (uint64_t j = 0; j < 16 && j <= n; j++) s += uni(seed, n - j, a);
  return s >> 2; /* arithmetic shift */
}
static int32_t isin(uint64_t p, int32_t amp) { return ((int32_t)X3_SIN1024[p & 1023] * amp) >> 15; }
static int16_t clamp16(int32_t v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

static int32_t s4_kind_sample(uint32_t seed, uint64_t n, uint32_t kind) {
  switch (kind) {
    case 0: return uni(seed, n, 32767);
    case 1: return uni(seed, n, 300);
    case 2: return uni(seed, n, 12);
    case 3: return 32767;
    default: return -32768;
  }
}

int x3o_synth(int kind, uint32_t seed, uint32_t fs, uint64_t n0, uint64_t count, int16_t *out) {
  static const int32_t S2_A[4] = {2, 8, 24, 40};
  if (fs == 0) return X3O_ERR_REFERENCE_PANIC;
  for (uint64_t i = 0; i < count; i++) {
    uint64_t n = n0 + i;
    int32_t v;
    if (kind == 1) {
      v = -3460 + colored(seed, n, 8) + isin((n * 50ull * 1024ull) / fs, 200);
    } else if (kind == 2) {
      int32_t a = S2_A[(n / fs) % 4];
      v = colored(seed, n, a);
      if (n % 196608ull < 64) v += isin((n * 48000ull * 1024ull) / fs, 6000);
    } else if (kind == 4) {
      uint64_t seg = n / 4096;
      uint32_t k = (uint32_t)(hh(seed ^ 0xABCDu, seg) % 6);
      if (k == 5) k = (uint32_t)(hh(seed ^ 0x1234u, n / 20) % 5);
      v = s4_kind_sample(seed, n, k);
    } else {
      return X3O_ERR_REFERENCE_PANIC;
    }
    out[i] = clamp16(v);
  }
  return X3O_OK;
}
