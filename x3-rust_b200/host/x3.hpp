// x3.hpp -- C++17 host-side mirror of the reference crate's public API over the C ABI (include/x3_b200.h).
//
// The reference is a Rust crate; no Rust toolchain exists in the build image, so this header is the compiled
// host layer that is actually built and tested (rust/ holds the equivalent Rust crate as source only).
// Names, argument meaning and error behaviour follow the reference:
//   x3::Parameters / Channel / IterChannel / FrameHeader          src/x3.rs
//   x3::encoder::encode / encode_frame / write_frame_header        src/encoder.rs:51,175,122
//   x3::decoder::decode_frame / read_frame_header                  src/decoder.rs:36,69
//   x3::bytewriter::SliceByteWriter / StreamByteWriter             src/bytewriter.rs
//   x3::encodefile::wav_to_x3a, x3::decodefile::x3a_to_wav / X3aReader   src/encodefile.rs:48, src/decodefile.rs:189,47
// Errors: the reference returns Result<T, X3Error>; here X3Error is thrown (code() is the C-ABI code).
// Panics of the reference (unreadable file, non-16-bit or non-mono WAV) are std::runtime_error.
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/x3_b200.h"

namespace x3 {

class X3Error : public std::runtime_error {
 public:
  explicit X3Error(int code) : std::runtime_error(describe(code)), code_(code) {}
  int code() const { return code_; }

 private:
  static std::string describe(int c) {
    std::string s = x3_strerror(c);
    if (c == X3_ERR_CUDA) s += std::string(": ") + x3_last_cuda_error();
    return s;
  }
  int code_;
};
inline void check(int rc) { if (rc != X3_OK) throw X3Error(rc); }

struct Archive { static constexpr const char *ID = "X3ARCHIV"; static constexpr size_t ID_LEN = 8; };  // x3.rs:136-141
struct Frame { static constexpr size_t MAX_LENGTH = 0x7fe0; };                                          // x3.rs:143-146

struct FrameHeader {  // x3.rs:147-184
  static constexpr size_t LENGTH = 20;
  static constexpr uint16_t KEY = 30771;
  uint8_t source_id = 0;
  uint16_t samples = 0;
  uint8_t channels = 0;
  size_t payload_len = 0;
  uint16_t payload_crc = 0;
};

struct Parameters {  // x3.rs:81-134
  static constexpr size_t MAX_BLOCK_LENGTH = 60, DEFAULT_BLOCK_LENGTH = 20, DEFAULT_BLOCKS_PER_FRAME = 500;
  size_t block_len = 20, blocks_per_frame = 500;
  std::array<size_t, 3> codes{{0, 1, 3}}, thresholds{{3, 8, 20}};
  Parameters() = default;  // Parameters::default()
  Parameters(size_t bl, size_t bpf, std::array<size_t, 3> c, std::array<size_t, 3> t)  // Parameters::new -> InvalidEncodingThresh
      : block_len(bl), blocks_per_frame(bpf), codes(c), thresholds(t) {
    x3_params p = c_struct();
    check(x3_params_validate(&p));
  }
  x3_params c_struct() const {
    x3_params p;
    p.block_len = (uint32_t)block_len;
    p.blocks_per_frame = (uint32_t)blocks_per_frame;
    for (int i = 0; i < 3; i++) { p.codes[i] = (uint32_t)codes[i]; p.thresholds[i] = (uint32_t)thresholds[i]; }
    return p;
  }
};

struct Channel {  // x3.rs:29-45
  uint16_t id;
  const int16_t *wav;
  size_t len;
  uint32_t sample_rate;
  Parameters params;
  Channel(uint16_t id_, const int16_t *w, size_t n, uint32_t fs, Parameters p) : id(id_), wav(w), len(n), sample_rate(fs), params(p) {}
};

template <class It>
struct IterChannel {  // x3.rs:47-69; drained into contiguous memory when encoded (the GPU needs a slice)
  uint16_t id;
  It first, last;
  uint32_t sample_rate;
  Parameters params;
  IterChannel(uint16_t id_, It b, It e, uint32_t fs, Parameters p) : id(id_), first(b), last(e), sample_rate(fs), params(p) {}
};

struct X3aSpec { uint32_t sample_rate = 0; Parameters params; uint8_t channels = 0; };  // x3.rs:70-79

namespace bytewriter {
struct SliceByteWriter {  // bytewriter.rs:27-100
  uint8_t *slice;
  size_t len, p_byte = 0, stream_length = 0;
  SliceByteWriter(uint8_t *buf, size_t n) : slice(buf), len(n) {}
  void write_all(const uint8_t *v, size_t n) {
    if (n > len - p_byte) throw X3Error(X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY);
    std::memcpy(slice + p_byte, v, n);
    p_byte += n;
    if (p_byte > stream_length) stream_length = p_byte;
  }
  void align2() { if (p_byte % 2) { uint8_t z = 0; write_all(&z, 1); } }
  size_t stream_position() const { return p_byte; }
  size_t capacity_left() const { return len - p_byte; }
  bool bounded() const { return true; }
};
struct StreamByteWriter {  // bytewriter.rs:115-164
  std::ostream &w;
  explicit StreamByteWriter(std::ostream &s) : w(s) {}
  void write_all(const uint8_t *v, size_t n) { w.write(reinterpret_cast<const char *>(v), (std::streamsize)n); }
  void align2() { if (stream_position() % 2) { uint8_t z = 0; write_all(&z, 1); } }
  size_t stream_position() const { return (size_t)w.tellp(); }
  size_t capacity_left() const { return SIZE_MAX; }
  bool bounded() const { return false; }
};
}  // namespace bytewriter

namespace encoder {
inline std::array<uint64_t, 6> &last_stats() { static std::array<uint64_t, 6> s{}; return s; }

inline void print_stats(const std::array<uint64_t, 6> &s) {  // encoder.rs:96-108
  float t = 0;
  for (auto v : s) t += (float)v;
  std::printf("\nStatistics:\n  Rice-0: %.4f%%\n  Rice-1: %.4f%%\n  Rice-2: %.4f%%\n  Rice-3: %.4f%%\n  BFP: %.4f%%\n  Pass-through %.4f%%\n\n",
              s[0] / t * 100, s[1] / t * 100, s[2] / t * 100, s[3] / t * 100, s[4] / t * 100, s[5] / t * 100);
}

// contiguous samples -> frame bytes (GPU)
inline std::vector<uint8_t> encode_slice(const int16_t *wav, size_t n, const Parameters &params, std::array<uint64_t, 6> *stats = nullptr) {
  x3_params p = params.c_struct();
  check(x3_params_validate(&p));
  std::vector<uint8_t> out(x3_encode_bound(n, &p));
  size_t len = 0;
  x3_stats st;
  check(x3_encode_host(wav, n, &p, out.data(), out.size(), &len, &st));
  out.resize(len);
  if (stats) for (int i = 0; i < 6; i++) (*stats)[i] = st.samples_by_mode[i];
  return out;
}

// encoder::encode(&[&Channel], writer) -- README.md:43-50 form and encoder.rs:51 (one channel only)
template <class W>
void encode(const std::vector<const Channel *> &channels, W &writer, bool quiet = false) {
  if (channels.size() > 1) throw X3Error(X3_ERR_MORE_THAN_ONE_CHANNEL);  // encoder.rs:55-57
  const Channel &ch = *channels.at(0);
  std::array<uint64_t, 6> stats{};
  if (ch.len) {
    writer.align2();  // encode_frame, encoder.rs:182
    std::vector<uint8_t> data = encode_slice(ch.wav, ch.len, ch.params, &stats);
    if (writer.bounded() && data.size() > writer.capacity_left()) throw X3Error(X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY);
    writer.write_all(data.data(), data.size());
  }
  last_stats() = stats;
  if (!quiet) print_stats(stats);
}
template <class It, class W>
void encode(IterChannel<It> &ch, W &writer, bool quiet = false) {  // encoder.rs:51 form
  std::vector<int16_t> pcm(ch.first, ch.last);
  Channel c(ch.id, pcm.data(), pcm.size(), ch.sample_rate, ch.params);
  encode(std::vector<const Channel *>{&c}, writer, quiet);
}

template <class W>
void encode_frame(const int16_t *wav, size_t n, W &writer, const Parameters &params, std::array<uint64_t, 6> &stats) {  // encoder.rs:175
  x3_params p = params.c_struct();
  std::vector<uint8_t> out(std::max<size_t>(x3_encode_frame_bound(n, &p), 32));   // 2.75 bytes per sample at block_len 1
  size_t len = 0;
  x3_stats st;
  writer.align2();
  check(x3_encode_frame_host(wav, n, &p, out.data(), out.size(), &len, &st));
  if (writer.bounded() && len > writer.capacity_left()) throw X3Error(X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY);
  writer.write_all(out.data(), len);
  for (int i = 0; i < 6; i++) stats[i] += st.samples_by_mode[i];
}

inline std::array<uint8_t, 20> write_frame_header(size_t num_samples, uint8_t id, size_t payload_len, uint16_t payload_crc) {  // encoder.rs:122
  std::array<uint8_t, 20> h{};
  check(x3_write_frame_header(num_samples, id, payload_len, payload_crc, h.data()));
  return h;
}
}  // namespace encoder

namespace decoder {
inline FrameHeader read_frame_header(const uint8_t *bytes, size_t len) {  // decoder.rs:69-118
  x3_frame_header h;
  check(x3_read_frame_header(bytes, len, &h));
  FrameHeader r;
  r.source_id = h.source_id; r.samples = h.samples; r.channels = h.channels; r.payload_len = h.payload_len; r.payload_crc = h.payload_crc;
  return r;
}
// decoder::decode_frame (decoder.rs:36-58): returns samples written
inline size_t decode_frame(const uint8_t *x3_bytes, size_t len, int16_t *wav_buf, size_t wav_cap, const Parameters &params, size_t samples) {
  x3_params p = params.c_struct();
  size_t n = 0;
  check(x3_decode_frame_host(x3_bytes, len, &p, wav_buf, wav_cap, samples, &n));
  return n;
}
// frame loop of decodefile.rs:105-136 over a bare frame stream; rc = what the reference would propagate
inline std::vector<int16_t> decode_stream(const uint8_t *frames, size_t len, const Parameters &params, int *rc, x3_decode_result *res = nullptr) {
  size_t total = 0, pos = 0;
  while (len - pos > 20) {  // sizing walk
    x3_frame_header h;
    if (x3_read_frame_header(frames + pos, 20, &h) != X3_OK) break;
    total += h.samples;
    pos += 20 + h.payload_len;
  }
  std::vector<int16_t> pcm(total ? total : 1);
  x3_params p = params.c_struct();
  size_t n = 0;
  x3_decode_result r;
  *rc = x3_decode_host(frames, len, &p, pcm.data(), total, &n, &r);
  if (*rc == X3_ERR_CUDA || *rc == X3_ERR_INVALID_ARGUMENT) throw X3Error(*rc);
  if (res) *res = r;
  pcm.resize(n);
  return pcm;
}
}  // namespace decoder

// ---- canonical 16-bit mono PCM WAV (what hound reads / writes for this crate) ----
namespace wav {
inline uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline std::vector<int16_t> read_mono16(const std::string &path, uint32_t *fs) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);  // WavReader::open(..).unwrap(), encodefile.rs:49
  std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (d.size() < 12 || std::memcmp(d.data(), "RIFF", 4) || std::memcmp(d.data() + 8, "WAVE", 4)) throw std::runtime_error("not a RIFF/WAVE file");
  size_t pos = 12;
  bool have_fmt = false;
  std::vector<int16_t> pcm;
  while (pos + 8 <= d.size()) {
    uint32_t sz = rd32(&d[pos + 4]);
    const uint8_t *body = &d[pos + 8];
    if (!std::memcmp(&d[pos], "fmt ", 4) && sz >= 16) {
      if (rd16(body + 14) != 16) throw std::runtime_error("assertion failed: bits_per_sample == 16 (encodefile.rs:52)");
      if (rd16(body + 2) != 1) throw std::runtime_error("assertion failed: channels == 1 (encodefile.rs:55)");
      *fs = rd32(body + 4);
      have_fmt = true;
    } else if (!std::memcmp(&d[pos], "data", 4)) {
      size_t avail = d.size() - (pos + 8);
      size_t n = (sz < avail ? sz : avail) / 2;
      pcm.resize(n);
      std::memcpy(pcm.data(), body, n * 2);
      break;
    }
    pos += 8 + sz + (sz & 1);
  }
  if (!have_fmt) throw std::runtime_error("WAV has no fmt chunk");
  return pcm;
}
inline void write_header(std::ostream &f, uint32_t fs, size_t n) {   // canonical 44-byte PCM header for n mono 16-bit samples
  uint8_t h[44] = {'R', 'I', 'F', 'F', 0, 0, 0, 0, 'W', 'A', 'V', 'E', 'f', 'm', 't', ' ', 16, 0, 0, 0, 1, 0, 1, 0};
  auto w32 = [&](int o, uint32_t v) { h[o] = v; h[o + 1] = v >> 8; h[o + 2] = v >> 16; h[o + 3] = v >> 24; };
  w32(4, (uint32_t)(36 + 2 * n));
  w32(24, fs);
  w32(28, fs * 2);
  h[32] = 2; h[33] = 0; h[34] = 16; h[35] = 0;
  std::memcpy(h + 36, "data", 4);
  w32(40, (uint32_t)(2 * n));
  f.write(reinterpret_cast<const char *>(h), 44);
}
inline void write_mono16(const std::string &path, uint32_t fs, const int16_t *pcm, size_t n) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw X3Error(X3_ERR_IO);
  write_header(f, fs, n);
  f.write(reinterpret_cast<const char *>(pcm), (std::streamsize)(2 * n));
}
}  // namespace wav

namespace encodefile {
inline std::string archive_xml(uint32_t fs, const Parameters &p) {  // encodefile.rs:93-117
  char b[1024];
  std::snprintf(b, sizeof b,
                "<X3ARCH PROG=\"x3new.m\" VERSION=\"2.0\" /><CFG ID=\"0\" FTYPE=\"XML\" /><CFG ID=\"1\" FTYPE=\"WAV\">"
                "<FS UNIT=\"Hz\">%u</FS><SUFFIX>wav</SUFFIX><CODEC TYPE=\"X3\" VERS=\"2\"><BLKLEN>%zu</BLKLEN>"
                "<CODES N=\"4\">RICE%zu,RICE%zu,RICE%zu,BFP</CODES><FILTER>DIFF</FILTER><NBITS>16</NBITS>"
                "<T N=\"3\">%zu,%zu,%zu</T></CODEC></CFG>",
                fs, p.block_len, p.codes[0], p.codes[1], p.codes[2], p.thresholds[0], p.thresholds[1], p.thresholds[2]);
  return b;
}
inline std::vector<uint8_t> create_archive_header(uint32_t fs, const Parameters &p) {  // encodefile.rs:82-138
  std::string xml = archive_xml(fs, p);
  std::vector<uint8_t> payload(xml.begin(), xml.end());
  if (payload.size() % 2) payload.push_back(0);  // pad byte is inside the CRC, :123-128
  auto h = encoder::write_frame_header(0, 0, payload.size(), x3_crc16(payload.data(), payload.size()));
  std::vector<uint8_t> out(Archive::ID, Archive::ID + 8);
  out.insert(out.end(), h.begin(), h.end());
  out.insert(out.end(), payload.begin(), payload.end());
  return out;
}
inline void wav_to_x3a(const std::string &wav_filename, const std::string &x3a_filename, bool quiet = false) {  // encodefile.rs:48-78
  uint32_t fs = 0;
  std::vector<int16_t> pcm = wav::read_mono16(wav_filename, &fs);
  Parameters params;  // always default, encodefile.rs:57
  std::ofstream f(x3a_filename, std::ios::binary);
  if (!f) throw X3Error(X3_ERR_IO);
  bytewriter::StreamByteWriter w(f);
  auto hdr = create_archive_header(fs, params);
  w.write_all(hdr.data(), hdr.size());
  Channel ch(0, pcm.data(), pcm.size(), fs, params);
  encoder::encode(std::vector<const Channel *>{&ch}, w, quiet);
}
}  // namespace encodefile

namespace decodefile {
inline std::string xml_first(const std::string &xml, const std::string &tag) {
  size_t a = 0;
  while ((a = xml.find("<" + tag, a)) != std::string::npos) {
    char c = xml[a + 1 + tag.size()];
    size_t gt = xml.find('>', a);
    if ((c == '>' || c == ' ') && gt != std::string::npos && xml[gt - 1] != '/') {
      size_t e = xml.find('<', gt);
      std::string t = xml.substr(gt + 1, e - gt - 1);
      size_t s0 = t.find_first_not_of(" \t\r\n"), s1 = t.find_last_not_of(" \t\r\n");
      return s0 == std::string::npos ? "" : t.substr(s0, s1 - s0 + 1);
    }
    a++;
  }
  throw X3Error(X3_ERR_REFERENCE_PANIC);  // fs[0] index panic in the reference
}
inline void parse_xml(const std::string &xml, uint32_t *fs, Parameters *params, bool quiet) {  // decodefile.rs:232-303
  std::string sfs = xml_first(xml, "FS"), sbl = xml_first(xml, "BLKLEN"), sc = xml_first(xml, "CODES"), st = xml_first(xml, "T");
  if (!quiet) std::printf("sample rate: %s\nblock length: %s\nRice codes: %s\nthresholds: %s\n", sfs.c_str(), sbl.c_str(), sc.c_str(), st.c_str());
  std::vector<size_t> ids, ths;
  size_t p = 0;
  while (p <= sc.size()) {
    size_t q = sc.find(',', p);
    std::string w = sc.substr(p, q == std::string::npos ? std::string::npos : q - p);
    if (w == "RICE0" || w == "RICE1" || w == "RICE2" || w == "RICE3") ids.push_back((size_t)(w[4] - '0'));
    else if (w != "BFP") throw X3Error(X3_ERR_ARCHIVE_XML_RICE_CODE);
    if (q == std::string::npos) break;
    p = q + 1;
  }
  p = 0;
  while (p <= st.size()) {
    size_t q = st.find(',', p);
    ths.push_back((size_t)std::stoul(st.substr(p, q == std::string::npos ? std::string::npos : q - p)));
    if (q == std::string::npos) break;
    p = q + 1;
  }
  if (ids.size() < 3 || ths.size() < 3) throw X3Error(X3_ERR_REFERENCE_PANIC);
  *fs = (uint32_t)std::stoul(sfs);
  *params = Parameters(std::stoul(sbl), Parameters::DEFAULT_BLOCKS_PER_FRAME, {{ids[0], ids[1], ids[2]}}, {{ths[0], ths[1], ths[2]}});
}

// decodefile.rs:47-137.  The file is read in pieces of `chunk_bytes` (32 MiB unless X3_STREAM_CHUNK says otherwise), each
// piece is cut at the last whole frame by the reference's own walk (header, payload_len, next header), decoded on the
// GPU in one call, and the frames are handed out one at a time; only the current piece is in memory.  The piece that
// reaches the end of the file is decoded as it is (a truncated last frame, a junk tail, a bad header: decode_stream
// reports what the reference would).
class X3aReader {
 public:
  static X3aReader open(const std::string &filename, bool quiet = false, size_t chunk_bytes = 0) {
    return X3aReader(filename, quiet, chunk_bytes);
  }
  const X3aSpec &spec() const { return spec_; }
  size_t frame_errors() const { return frame_errors_; }
  // returns false at the end of the stream (Ok(None)); throws what the reference would return as Err
  bool decode_next_frame(std::vector<int16_t> &wav_buf, size_t *samples) {
    while (next_ >= sizes_.size()) {
      if (final_rc_ != X3_OK) { int rc = final_rc_; final_rc_ = X3_OK; done_ = true; throw X3Error(rc); }
      if (done_ || !next_batch()) return false;
    }
    wav_buf.assign(pcm_.begin() + (std::ptrdiff_t)starts_[next_], pcm_.begin() + (std::ptrdiff_t)(starts_[next_] + sizes_[next_]));
    *samples = sizes_[next_++];
    return true;
  }
  // the frames of the current batch that have not been handed out yet, as one run of samples (x3a_to_wav)
  bool decode_batch(const int16_t **pcm, size_t *samples) {
    while (next_ >= sizes_.size()) {
      if (final_rc_ != X3_OK) { int rc = final_rc_; final_rc_ = X3_OK; done_ = true; throw X3Error(rc); }
      if (done_ || !next_batch()) return false;
    }
    *pcm = pcm_.data() + starts_[next_];
    *samples = pcm_.size() - starts_[next_];
    next_ = sizes_.size();
    return true;
  }

 private:
  X3aReader(const std::string &filename, bool quiet, size_t chunk_bytes) : f_(filename, std::ios::binary) {
    if (!f_) throw std::runtime_error("cannot open " + filename);  // File::open(..).unwrap(), decodefile.rs:60
    const char *env = std::getenv("X3_STREAM_CHUNK");
    chunk_ = chunk_bytes ? chunk_bytes : (env && *env ? (size_t)std::strtoull(env, nullptr, 10) : (size_t)32 << 20);
    if (chunk_ < 64 * 1024) chunk_ = 64 * 1024;   // a piece must hold at least one frame (< 32 KiB)
    uint8_t head[28];
    if (!f_.read(reinterpret_cast<char *>(head), 8)) throw X3Error(X3_ERR_IO);
    if (std::memcmp(head, Archive::ID, 8)) throw X3Error(X3_ERR_ARCHIVE_INVALID_KEY);  // decodefile.rs:147-149
    if (!f_.read(reinterpret_cast<char *>(head) + 8, 20)) throw X3Error(X3_ERR_IO);
    FrameHeader h = decoder::read_frame_header(head + 8, 20);
    std::string xml(h.payload_len, '\0');
    if (h.payload_len && !f_.read(&xml[0], (std::streamsize)h.payload_len)) throw X3Error(X3_ERR_IO);
    parse_xml(xml, &spec_.sample_rate, &spec_.params, quiet);
    spec_.channels = h.channels;
  }
  // whole frames at the front of buf: their bytes; `stopped` when a header is bad or a frame is longer than the
  // reference's read buffer -- the decode of the rest of the file reports that
  static size_t walk(const std::vector<uint8_t> &buf, bool *stopped) {
    size_t pos = 0;
    *stopped = false;
    while (buf.size() - pos > 20) {
      x3_frame_header h;
      if (x3_read_frame_header(buf.data() + pos, 20, &h) != X3_OK) { *stopped = true; return pos; }
      if (buf.size() - pos - 20 < h.payload_len) break;   // continues in the part of the file not read yet
      pos += 20 + h.payload_len;
      if (h.payload_len > 24 * 1024) { *stopped = true; return pos; }   // X3_READ_BUFFER_SIZE, decodefile.rs:44,118-121
    }
    return pos;
  }
  bool next_batch() {
    for (;;) {
      if (!eof_) {
        const size_t old = carry_.size();
        carry_.resize(old + chunk_);
        f_.read(reinterpret_cast<char *>(carry_.data() + old), (std::streamsize)chunk_);
        const size_t got = (size_t)f_.gcount();
        carry_.resize(old + got);
        if (got < chunk_) eof_ = true;
      }
      bool stopped = false;
      size_t take = carry_.size();
      if (!eof_) {
        take = walk(carry_, &stopped);
        if (stopped) {   // hand over everything that is left: the decode reports the bad frame
          std::vector<uint8_t> rest((std::istreambuf_iterator<char>(f_)), std::istreambuf_iterator<char>());
          carry_.insert(carry_.end(), rest.begin(), rest.end());
          eof_ = true;
          take = carry_.size();
        } else if (take == 0) {
          continue;      // not one whole frame yet
        }
      }
      if (eof_) done_ = true;
      x3_decode_result r;
      pcm_ = decoder::decode_stream(carry_.data(), take, spec_.params, &final_rc_, &r);
      frame_errors_ += (size_t)r.frame_errors;
      sizes_.clear();
      starts_.clear();
      next_ = 0;
      size_t pos = 0, start = 0;
      for (uint64_t i = 0; i < r.frames; i++) {
        FrameHeader fh = decoder::read_frame_header(carry_.data() + pos, 20);
        sizes_.push_back(fh.samples);
        starts_.push_back(start);
        start += fh.samples;
        pos += 20 + fh.payload_len;
      }
      if (final_rc_ != X3_OK || r.frame_errors) done_ = true;   // the reference stops at the first bad frame
      carry_.erase(carry_.begin(), carry_.begin() + (std::ptrdiff_t)take);
      return !sizes_.empty() || final_rc_ != X3_OK;
    }
  }
  std::ifstream f_;
  X3aSpec spec_;
  std::vector<uint8_t> carry_;           // bytes read but not decoded yet (a partial frame)
  std::vector<int16_t> pcm_;             // PCM of the current batch
  std::vector<size_t> sizes_, starts_;
  size_t next_ = 0, frame_errors_ = 0, chunk_ = 0;
  bool eof_ = false, done_ = false;
  int final_rc_ = X3_OK;
};

inline void x3a_to_wav(const std::string &x3a_filename, const std::string &wav_filename, bool quiet = false) {  // decodefile.rs:189-212
  X3aReader rd = X3aReader::open(x3a_filename, quiet);
  // the WAV is written batch by batch; its two length fields are patched at the end (hound does the same on finalize)
  std::ofstream f(wav_filename, std::ios::binary);
  if (!f) throw X3Error(X3_ERR_IO);
  wav::write_header(f, rd.spec().sample_rate, 0);
  size_t total = 0, n = 0;
  const int16_t *pcm = nullptr;
  int pending = X3_OK;
  try {
    while (rd.decode_batch(&pcm, &n)) {
      f.write(reinterpret_cast<const char *>(pcm), (std::streamsize)(2 * n));
      total += n;
    }
  } catch (const X3Error &e) { pending = e.code(); }
  f.seekp(0);
  wav::write_header(f, rd.spec().sample_rate, total);   // samples before the error are kept
  if (pending != X3_OK) throw X3Error(pending);
}
}  // namespace decodefile

}  // namespace x3
