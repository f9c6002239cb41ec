// x3_api.cu -- the C ABI of include/x3_b200.h over the CUDA kernels.  No CPU fallback: every computing
// entry point fails with X3_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/x3_b200.h"
#include "x3_crc_host.h"
#include "x3_dec_core.cuh"
#include "x3_enc_core.cuh"
#include "x3_kernels.h"

using namespace x3;

namespace {

std::atomic<uint64_t> g_launches{0};
thread_local char tl_cuda_err[256] = "";
thread_local float tl_ms[4] = {0.f, 0.f, 0.f, 0.f};
thread_local int tl_enc_kernel = -1;   // kernel kind of the most recent x3_encode_device on this thread (x3_last_encode_kernel)
thread_local uint32_t tl_max_payload = x3::kReadBufferSize;   // x3_decode_frame_host lifts the stream reader's limit

int cuda_fail(cudaError_t e, const char *what) {
  snprintf(tl_cuda_err, sizeof tl_cuda_err, "%s: %s", what, cudaGetErrorString(e));
  cudaGetLastError();
  return X3_ERR_CUDA;
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// ---- CRC table bank (layout in x3_common.cuh) -------------------------------------------------
uint16_t g_crc_host[kCrcBankEntries3];
std::once_flag g_crc_once;

void build_crc_host() { build_crc_bank(g_crc_host); }
const uint16_t *crc_host() {
  std::call_once(g_crc_once, build_crc_host);
  return g_crc_host;
}

struct DeviceState {
  uint16_t *crc_dev = nullptr;
  int sms = 0;
  bool pool_ready = false;
  int occ[3] = {0, 0, 0};          // encode kernels: resident CTAs per SM ...
  size_t occ_smem[3] = {0, 0, 0};  // ... at this dynamic shared-memory size
};
std::mutex g_dev_mu;
DeviceState g_dev[64];

int device_state(DeviceState **out) {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return X3_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DeviceState &d = g_dev[dev];
  if (!d.crc_dev) {
    CU(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    uint16_t *p = nullptr;
    CU(cudaMalloc(&p, sizeof(uint16_t) * kCrcBankEntries3));
    CU(cudaMemcpy(p, crc_host(), sizeof(uint16_t) * kCrcBankEntries3, cudaMemcpyHostToDevice));
    d.crc_dev = p;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;  // keep freed workspace cached in the pool between calls
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      d.pool_ready = true;
    }
    cudaGetLastError();
  }
  *out = &d;
  return X3_OK;
}

// small pinned buffer for result read-back, one per thread
struct Pinned {
  unsigned long long *p = nullptr;
  ~Pinned() { if (p) cudaFreeHost(p); }
};
thread_local Pinned tl_pinned;
int pinned(unsigned long long **out) {
  if (!tl_pinned.p) CU(cudaMallocHost(&tl_pinned.p, 64 * sizeof(unsigned long long)));
  *out = tl_pinned.p;
  return X3_OK;
}

// CUDA event pairs for the kernel timings behind x3_last_kernel_ms.  Creating and destroying eight events per call
// cost more host time than the calls' own launches, so each thread keeps a small pool.
struct EventPool {
  cudaEvent_t ev[16];
  int n = 0, used = 0, device = -1;
};
thread_local EventPool tl_events;
cudaEvent_t pooled_event() {
  EventPool &p = tl_events;
  int dev = 0;
  cudaGetDevice(&dev);
  if (p.device != dev) { p.device = dev; p.n = 0; p.used = 0; }  // events of another device are left to teardown
  if (p.used < p.n) return p.ev[p.used++];
  cudaEvent_t e = nullptr;
  if (p.n < 16 && cudaEventCreate(&e) == cudaSuccess) {
    p.ev[p.n++] = e;
    p.used = p.n;
    return e;
  }
  cudaGetLastError();
  return nullptr;
}
void release_pooled_events() { tl_events.used = 0; }   // at the start of every call that times kernels

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  explicit Timer(cudaStream_t st) : s(st) {
    a = pooled_event();
    b = pooled_event();
  }
  void start() { if (a) cudaEventRecord(a, s); }
  void stop() { if (b) cudaEventRecord(b, s); }
  float ms() {
    float t = 0.f;
    if (!a || !b || cudaEventElapsedTime(&t, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
    return t;
  }
};

// Payload CRC and frame decode are independent given the frame table: the CRC kernel runs on a second stream beside
// the decode kernel and fills the SMs that kernel leaves idle while its last (serial) frames finish.
struct DecodeFork {
  int device = -1;
  cudaStream_t s2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
thread_local DecodeFork tl_fork;
cudaStream_t fork_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  DecodeFork &f = tl_fork;
  if (f.device != dev || !f.s2) {
    f.device = dev;
    if (cudaStreamCreateWithFlags(&f.s2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) != cudaSuccess) {
      f.s2 = nullptr;
      cudaGetLastError();
      return nullptr;
    }
  }
  return f.s2;
}

// ---- parameter handling -----------------------------------------------------------------------
constexpr size_t kMaxDynSmem = 227 * 1024;

struct Derived {
  CodecParams P;
  uint32_t max_blocks;
  uint32_t max_block_bits;
  uint32_t out_words_cap;
  int kind;        // kEncKernel*: which encode kernel takes these parameters
  size_t smem;
};

// Parameters::default() inputs have two kernels: the strip kernel (ordinary audio) and the block-per-thread kernel of
// round 1 (inputs whose frames overflow the strip kernel's windows); encode_probe_kernel picks per call.
// X3_ENC_KERNEL=fast / =strip forces one of them (A/B measurements).
int forced_kernel() {
  static const int v = [] {
    const char *e = getenv("X3_ENC_KERNEL");
    return !e ? -1 : !strcmp(e, "fast") ? (int)kEncKernelFast : !strcmp(e, "strip") ? (int)kEncKernelStrip : -1;
  }();
  return v;
}
bool use_old_fast_kernel() { return forced_kernel() == (int)kEncKernelFast; }
size_t encode_smem_of(int kind, const CodecParams &P, uint32_t max_blocks, uint32_t out_words_cap) {
  return kind == kEncKernelStrip ? encode_strip_smem_bytes()
         : kind == kEncKernelFast ? encode_fast_smem_bytes(P, out_words_cap)
                                  : encode_smem_bytes(P, max_blocks, out_words_cap);
}

uint32_t max_block_bits_of(const x3_params *p) {
  const bool sorted = p->thresholds[0] <= p->thresholds[1] && p->thresholds[1] <= p->thresholds[2];
  uint32_t rice_max = 0;
  for (int f = 0; f < 3; f++) {
    uint32_t t = sorted ? p->thresholds[f] : p->thresholds[2];
    if (t > p->thresholds[2]) t = p->thresholds[2];
    uint32_t l = ((2u * t) >> p->codes[f]) + p->codes[f] + 1u;
    if (l > rice_max) rice_max = l;
  }
  uint32_t a = 2u + p->block_len * rice_max, b = 6u + 16u * p->block_len;
  return a > b ? a : b;
}

thread_local bool tl_one_frame = false;   // set by the encode_frame entry points: the frame is the whole input
int derive(const x3_params *p, Derived *d) {
  const bool one_frame = tl_one_frame;
  if (!p) return X3_ERR_INVALID_ARGUMENT;
  if (p->block_len == 0 || p->blocks_per_frame == 0) return X3_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < 3; k++)
    if (p->codes[k] > 3) return X3_ERR_INVALID_ARGUMENT;  // RiceCodes::get would index out of bounds, x3.rs:256
  for (int k = 0; k < 2; k++)                              // x3.rs:107-112
    if (p->thresholds[k] > rice_offset(p->codes[k])) return X3_ERR_INVALID_ENCODING_THRESH;
  if (p->block_len > (uint32_t)kMaxBlockLen) return X3_ERR_UNSUPPORTED_PARAMS;  // reference panics (wav_diff[60])
  if (p->thresholds[2] > 4096u) return X3_ERR_UNSUPPORTED_PARAMS;
  const unsigned long long spf = (unsigned long long)p->block_len * p->blocks_per_frame;
  // header field is u16 (encoder.rs:141); a single frame may round its block count up past it, its sample count cannot
  if (spf > 65535ull + (one_frame ? p->block_len - 1u : 0u)) return X3_ERR_UNSUPPORTED_PARAMS;
  d->P.block_len = p->block_len;
  d->P.spf = (uint32_t)spf;
  for (int k = 0; k < 3; k++) { d->P.codes[k] = p->codes[k]; d->P.thresholds[k] = p->thresholds[k]; }
  d->max_blocks = p->blocks_per_frame;
  d->max_block_bits = max_block_bits_of(p);
  const unsigned long long bits = 16ull + (unsigned long long)d->max_blocks * d->max_block_bits;
  d->out_words_cap = (uint32_t)((bits + 31ull) / 32ull) + 1u;
  d->kind = kEncKernelGeneric;
  if (params_are_default(d->P) && d->max_blocks <= 512u) d->kind = use_old_fast_kernel() ? kEncKernelFast : kEncKernelStrip;
  d->smem = encode_smem_of(d->kind, d->P, d->max_blocks, d->out_words_cap);
  if (d->smem > kMaxDynSmem) return X3_ERR_UNSUPPORTED_PARAMS;
  return X3_OK;
}

size_t frame_bound(uint32_t n, const Derived &d) {
  const uint32_t nblk = n > 1 ? (n - 2u) / d.P.block_len + 1u : 1u;
  const unsigned long long bits = 16ull + (unsigned long long)nblk * d.max_block_bits;
  return (size_t)kFrameHeaderLen + (size_t)(((bits + 15ull) >> 4) << 1);
}

int map_frame_error(int st) { return st; }  // kDec* codes are the public X3_ERR_* values

}  // namespace

// ================================================================================================
extern "C" {

int x3_abi_version(void) { return X3_B200_ABI_VERSION; }

int x3_params_default(x3_params *p) {
  if (!p) return X3_ERR_INVALID_ARGUMENT;
  p->block_len = 20;
  p->blocks_per_frame = 500;
  p->codes[0] = 0; p->codes[1] = 1; p->codes[2] = 3;
  p->thresholds[0] = 3; p->thresholds[1] = 8; p->thresholds[2] = 20;
  return X3_OK;
}

int x3_params_validate(const x3_params *p) {
  Derived d;
  return derive(p, &d);
}

size_t x3_encode_bound(size_t n_samples, const x3_params *p) {
  Derived d;
  if (derive(p, &d) != X3_OK || n_samples == 0) return 0;
  const size_t full = n_samples / d.P.spf, rem = n_samples % d.P.spf;
  size_t b = full * frame_bound(d.P.spf, d);
  if (rem) b += frame_bound((uint32_t)rem, d);
  return b;
}

const char *x3_strerror(int code) {
  switch (code) {
    case X3_OK: return "ok";
    case X3_ERR_INVALID_ENCODING_THRESH: return "InvalidEncodingThresh: threshold must be <= code.offset";
    case X3_ERR_OUT_OF_BOUNDS_INVERSE: return "OutOfBoundsInverse";
    case X3_ERR_MORE_THAN_ONE_CHANNEL: return "MoreThanOneChannel";
    case X3_ERR_ARCHIVE_XML_INVALID: return "ArchiveHeaderXMLInvalid";
    case X3_ERR_ARCHIVE_XML_RICE_CODE: return "ArchiveHeaderXMLRiceCode";
    case X3_ERR_ARCHIVE_INVALID_KEY: return "ArchiveHeaderXMLInvalidKey";
    case X3_ERR_FRAME_LENGTH: return "FrameLength";
    case X3_ERR_FRAME_HEADER_INVALID_KEY: return "FrameHeaderInvalidKey";
    case X3_ERR_FRAME_HEADER_INVALID_PAYLOAD_LEN: return "FrameHeaderInvalidPayloadLen";
    case X3_ERR_FRAME_HEADER_INVALID_HEADER_CRC: return "FrameHeaderInvalidHeaderCRC";
    case X3_ERR_FRAME_HEADER_INVALID_PAYLOAD_CRC: return "FrameHeaderInvalidPayloadCRC";
    case X3_ERR_FRAME_DECODE_INVALID_FTYPE: return "FrameDecodeInvalidFType";
    case X3_ERR_FRAME_DECODE_INVALID_BPF: return "FrameDecodeInvalidBPF";
    case X3_ERR_FRAME_DECODE_UNEXPECTED_END: return "FrameDecodeUnexpectedEnd";
    case X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY: return "ByteWriterInsufficientMemory";
    case X3_ERR_IO: return "Io: unexpected end of stream";
    case X3_ERR_INVALID_ARGUMENT: return "invalid argument";
    case X3_ERR_UNSUPPORTED_PARAMS: return "parameters outside the GPU path's limits";
    case X3_ERR_CUDA: return "CUDA error (see x3_last_cuda_error)";
    case X3_ERR_REFERENCE_PANIC: return "input on which the reference panics";
    default: return "unknown error";
  }
}

const char *x3_last_cuda_error(void) { return tl_cuda_err; }
uint64_t x3_kernel_launch_count(void) { return g_launches.load(); }
int x3_last_encode_kernel(void) { return tl_enc_kernel; }
int x3_last_kernel_ms(float ms[4]) {
  if (!ms) return X3_ERR_INVALID_ARGUMENT;
  ms[0] = tl_ms[0]; ms[1] = tl_ms[1]; ms[2] = tl_ms[2]; ms[3] = tl_ms[3];
  return X3_OK;
}

uint16_t x3_crc16(const uint8_t *data, size_t len) {
  const uint16_t *T = crc_host();
  uint32_t s = 0xffffu;
  for (size_t i = 0; i < len; i++) s = crc16_byte(T, s, data[i]);
  return (uint16_t)s;
}

int x3_write_frame_header(size_t num_samples, uint8_t id, size_t payload_len, uint16_t payload_crc,
                          uint8_t header[20]) {
  if (!header) return X3_ERR_INVALID_ARGUMENT;
  memset(header, 0, 20);
  header[0] = 0x78; header[1] = 0x33;
  header[2] = id; header[3] = id;  // encoder.rs:131,135
  header[4] = (uint8_t)(num_samples >> 8); header[5] = (uint8_t)num_samples;
  header[6] = (uint8_t)(payload_len >> 8); header[7] = (uint8_t)payload_len;
  const uint16_t hc = x3_crc16(header, 16);
  header[16] = (uint8_t)(hc >> 8); header[17] = (uint8_t)hc;
  header[18] = (uint8_t)(payload_crc >> 8); header[19] = (uint8_t)payload_crc;
  return X3_OK;
}

int x3_read_frame_header(const uint8_t *b, size_t len, x3_frame_header *h) {
  if (!b || !h) return X3_ERR_INVALID_ARGUMENT;
  if (len < 20) return X3_ERR_FRAME_DECODE_UNEXPECTED_END;                                   // decoder.rs:70
  if (x3_crc16(b, 16) != (uint16_t)((b[16] << 8) | b[17])) return X3_ERR_FRAME_HEADER_INVALID_HEADER_CRC;
  if (((b[0] << 8) | b[1]) != (int)kFrameKey) return X3_ERR_FRAME_HEADER_INVALID_KEY;
  h->source_id = b[2];
  h->channels = b[3];
  if (h->channels > 1) return X3_ERR_MORE_THAN_ONE_CHANNEL;
  h->samples = (uint16_t)((b[4] << 8) | b[5]);
  h->payload_len = (uint32_t)((b[6] << 8) | b[7]);
  if (h->payload_len >= kFrameMaxLength) return X3_ERR_FRAME_LENGTH;
  h->payload_crc = (uint16_t)((b[18] << 8) | b[19]);
  return X3_OK;
}

// ------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------
namespace {

size_t encode_ws_bytes(unsigned long long nf) { return (128 + 8 * (size_t)nf + 256 + 15) & ~(size_t)15; }  // + optional phase timing words

// per-thread resources of the pipelined host path (x3_encode_host)
struct HostPipe {
  int device = -1;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  int16_t *d_pcm[2] = {nullptr, nullptr};
  uint8_t *d_out[2] = {nullptr, nullptr};
  unsigned char *ws[2] = {nullptr, nullptr};
  size_t cap_pcm[2] = {0, 0}, cap_out[2] = {0, 0}, cap_ws[2] = {0, 0}, cap_res = 0;
  unsigned long long *h_res = nullptr;
  void release() {
    for (int b = 0; b < 2; b++) {
      if (d_pcm[b]) cudaFree(d_pcm[b]);
      if (d_out[b]) cudaFree(d_out[b]);
      if (ws[b]) cudaFree(ws[b]);
      if (ev_in[b]) cudaEventDestroy(ev_in[b]);
      if (ev_cmp[b]) cudaEventDestroy(ev_cmp[b]);
      if (ev_out[b]) cudaEventDestroy(ev_out[b]);
      d_pcm[b] = nullptr; d_out[b] = nullptr; ws[b] = nullptr; ev_in[b] = ev_cmp[b] = ev_out[b] = nullptr;
      cap_pcm[b] = cap_out[b] = cap_ws[b] = 0;
    }
    if (h_res) cudaFreeHost(h_res);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_cmp) cudaStreamDestroy(s_cmp);
    if (s_out) cudaStreamDestroy(s_out);
    h_res = nullptr; cap_res = 0; s_in = s_cmp = s_out = nullptr;
    cudaGetLastError();
  }
  ~HostPipe() {}  // left to process teardown: the CUDA context may already be gone when thread-locals die
};
thread_local HostPipe tl_pipe;

// zero the workspace and launch the encode kernel for n_samples (> 0) on `st`; no synchronisation
cudaError_t enqueue_encode(Derived d, DeviceState *ds, const int16_t *d_pcm, size_t n_samples, uint8_t *d_out, size_t out_cap,
                           unsigned char *ws, cudaStream_t st, int *rc_out, unsigned long long *result_dev = nullptr) {
  *rc_out = X3_OK;
  const unsigned long long nf = (n_samples + d.P.spf - 1) / d.P.spf;
  EncodeArgs a;
  a.pcm = d_pcm;
  a.n_samples = n_samples;
  a.out = d_out;
  a.out_cap = out_cap;
  a.P = d.P;
  a.n_frames = (uint32_t)nf;
  a.max_blocks = d.max_blocks;
  a.out_words_cap = d.out_words_cap;
  a.result = result_dev ? result_dev : reinterpret_cast<unsigned long long *>(ws);        // 8 words (the caller's, zeroed by the caller, in the stream-ordered API)
  a.ticket = reinterpret_cast<unsigned int *>(ws + 64);
  a.status = reinterpret_cast<unsigned long long *>(ws + 128);
  a.timing = reinterpret_cast<unsigned long long *>(ws + 128 + 8 * (size_t)nf);
  a.crc_tables = ds->crc_dev;
  a.neg_one = -1;
  a.choice = nullptr;
  // the fast kernels stage frames with 16-byte async copies: they need a 16-byte aligned base and frame size (and the
  // strip kernel a 2-byte aligned output: it writes halfwords and 16-byte aligned bulk stores)
  if (d.kind != kEncKernelGeneric && ((((uintptr_t)d_pcm) & 15u) != 0 || ((d.P.spf * 2u) & 15u) != 0 || (((uintptr_t)d_out) & 1u) != 0)) {
    d.kind = kEncKernelGeneric;
    d.smem = encode_smem_of(d.kind, d.P, d.max_blocks, d.out_words_cap);
    if (d.smem > kMaxDynSmem) { *rc_out = X3_ERR_UNSUPPORTED_PARAMS; return cudaSuccess; }
  }
  // one launch of kernel `kind` with its own shared-memory size and grid
  auto launch_kind = [&](int kind) -> cudaError_t {
    const size_t smem = encode_smem_of(kind, d.P, d.max_blocks, d.out_words_cap);
    // resident CTAs per SM of this kernel at this shared-memory size (cached: the query costs more than the launch)
    int occ;
    {
      std::lock_guard<std::mutex> lk(g_dev_mu);
      int &slot = ds->occ[kind];
      size_t &slot_smem = ds->occ_smem[kind];
      if (slot == 0 || slot_smem != smem) { slot = encode_occupancy(kind, smem); slot_smem = smem; }
      occ = slot;
    }
    unsigned long long grid = (unsigned long long)ds->sms * (unsigned)occ;
    if (kind != kEncKernelGeneric) {  // CTA 0 is the scanner, the others encode
      if (grid > nf + 1) grid = nf + 1;
      if (grid < 2) grid = 2;
    } else if (grid > nf) {
      grid = nf;
    }
    g_launches++;
    return launch_encode(a, kind, (int)grid, smem, st);
  };
  const size_t ws_bytes = encode_ws_bytes(nf);
  // Parameters::default(): the probe zeroes the workspace and picks between the strip kernel and the block-per-thread
  // kernel; both are launched, the one not picked returns at once (no host round trip).  A call too small to matter,
  // or a frame image too large for the block-per-thread kernel, just takes the strip kernel.
  const bool adaptive = d.kind == kEncKernelStrip && forced_kernel() < 0 && nf >= 64 &&
                        encode_smem_of(kEncKernelFast, d.P, d.max_blocks, d.out_words_cap) <= kMaxDynSmem;
  cudaError_t e;
  if (adaptive) {
    unsigned int *choice = reinterpret_cast<unsigned int *>(ws + 72);
    e = launch_encode_probe(d_pcm, n_samples, ws, ws_bytes, choice, st);
    g_launches++;
    if (e != cudaSuccess) return e;
    a.choice = choice;
    e = launch_kind(kEncKernelStrip);
    if (e != cudaSuccess) return e;
    return launch_kind(kEncKernelFast);
  }
  e = cudaMemsetAsync(ws, 0, ws_bytes, st);
  if (e != cudaSuccess) return e;
  return launch_kind(d.kind);
}

}  // namespace

int x3_encode_device(const int16_t *d_pcm, size_t n_samples, const x3_params *p, uint8_t *d_out, size_t out_cap,
                     size_t *out_len, x3_stats *stats, void *cuda_stream) {
  Derived d;
  int rc = derive(p, &d);
  if (rc) return rc;
  if (!out_len || (n_samples && (!d_pcm || !d_out))) return X3_ERR_INVALID_ARGUMENT;
  *out_len = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  tl_ms[0] = tl_ms[1] = tl_ms[2] = tl_ms[3] = 0.f;
  if (n_samples == 0) return X3_OK;
  if (((uintptr_t)d_pcm & 1u) != 0) return X3_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  DeviceState *ds;
  if ((rc = device_state(&ds))) return rc;
  unsigned long long *host_res;
  if ((rc = pinned(&host_res))) return rc;

  const unsigned long long nf = (n_samples + d.P.spf - 1) / d.P.spf;
  if (nf > 0xfffffff0ull) return X3_ERR_UNSUPPORTED_PARAMS;
  const size_t ws_bytes = encode_ws_bytes(nf);
  unsigned char *ws = nullptr;
  CU(cudaMallocAsync(&ws, ws_bytes, st));
  release_pooled_events();
  Timer tm(st);
  tm.start();
  cudaError_t e = enqueue_encode(d, ds, d_pcm, n_samples, d_out, out_cap, ws, st, &rc);
  tm.stop();
  if (rc) { cudaFreeAsync(ws, st); return rc; }
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_res, ws, 80, cudaMemcpyDeviceToHost, st);   // results + the probe's choice
#ifdef X3_ENC_TIMING
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_res + 16, ws + 128 + 8 * (size_t)nf, 256, cudaMemcpyDeviceToHost, st);
#endif
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFreeAsync(ws, st);
  if (e != cudaSuccess) return cuda_fail(e, "encode_frames_kernel");
  tl_ms[0] = tl_ms[2] = tm.ms();
  {
    const unsigned int picked = (unsigned int)(host_res[9] & 0xffffffffull);   // ws + 72: 0 unless the probe ran
    tl_enc_kernel = picked ? (int)picked : (d.kind == kEncKernelStrip && forced_kernel() == (int)kEncKernelFast ? (int)kEncKernelFast : -2);
  }
#ifdef X3_ENC_TIMING
  {
    const char *names[] = {"waitA", "measure", "waitB", "scan2", "pack", "waitD", "or", "waitE", "crc", "waitOff", "copy", "frames",
                           "c_waitSize", "c_hist", "c_poll", "c_waitCrc", "c_final", "c_frames"};
    fprintf(stderr, "[x3 timing] cycles per frame:");
    for (int k = 0; k < 11; k++) fprintf(stderr, " %s=%.0f", names[k], (double)host_res[16 + k] / (double)(host_res[16 + 11] ? host_res[16 + 11] : 1));
    fprintf(stderr, " |");
    for (int k = 12; k < 17; k++) fprintf(stderr, " %s=%.0f", names[k], (double)host_res[16 + k] / (double)(host_res[16 + 17] ? host_res[16 + 17] : 1));
    fprintf(stderr, "\n");
  }
#endif
  *out_len = (size_t)host_res[0];
  if (stats)
    for (int k = 0; k < 6; k++) stats->samples_by_mode[k] = host_res[2 + k];
  if (host_res[1]) return X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY;
  return X3_OK;
}

int x3_encode_host(const int16_t *pcm, size_t n_samples, const x3_params *p, uint8_t *out, size_t out_cap,
                   size_t *out_len, x3_stats *stats) {
  Derived d;
  int rc = derive(p, &d);
  if (rc) return rc;
  if (!out_len || (n_samples && (!pcm || !out))) return X3_ERR_INVALID_ARGUMENT;
  *out_len = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  if (n_samples == 0) return X3_OK;
  DeviceState *ds;
  if ((rc = device_state(&ds))) return rc;
  const unsigned long long nf = (n_samples + d.P.spf - 1) / d.P.spf;
  if (nf > 0xfffffff0ull) return X3_ERR_UNSUPPORTED_PARAMS;

  // Chunks of whole frames flow through three streams: H2D of chunk c+1, the kernel of chunk c and the D2H of chunk
  // c-1 overlap (PCIe is full duplex).  Frames are independent, so a chunk's stream is simply appended.
  unsigned long long chunk_mb = 48;
  if (const char *env = getenv("X3_HOST_CHUNK_MB")) chunk_mb = std::max(1, atoi(env));
  const unsigned long long chunk_frames = std::max<unsigned long long>(1, (chunk_mb << 20) / ((unsigned long long)d.P.spf * 2ull));
  const unsigned long long nchunks = (nf + chunk_frames - 1) / chunk_frames;
  const size_t chunk_samples = (size_t)chunk_frames * d.P.spf;
  const size_t chunk_bound = x3_encode_bound(std::min(chunk_samples, n_samples), p);
  if (nchunks <= 1) {
    // small input: one H2D, one kernel, one D2H on the default stream (pool allocations, no extra streams)
    cudaStream_t st = nullptr;
    const size_t dcap = chunk_bound < out_cap ? chunk_bound : out_cap;  // never write more than the caller has room for
    int16_t *dp = nullptr;
    uint8_t *dout = nullptr;
    CU(cudaMallocAsync(&dp, n_samples * sizeof(int16_t), st));
    cudaError_t e1 = cudaMallocAsync(&dout, dcap ? dcap : 1, st);
    if (e1 != cudaSuccess) { cudaFreeAsync(dp, st); return cuda_fail(e1, "cudaMallocAsync"); }
    e1 = cudaMemcpyAsync(dp, pcm, n_samples * sizeof(int16_t), cudaMemcpyHostToDevice, st);
    if (e1 == cudaSuccess) {
      rc = x3_encode_device(dp, n_samples, p, dout, dcap, out_len, stats, st);
      if (rc == X3_OK && *out_len) {
        e1 = cudaMemcpyAsync(out, dout, *out_len, cudaMemcpyDeviceToHost, st);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(st);
      }
    }
    cudaFreeAsync(dp, st);
    cudaFreeAsync(dout, st);
    if (e1 != cudaSuccess) return cuda_fail(e1, "x3_encode_host copies");
    return rc;
  }
  // streams, events, device buffers and the pinned result words are cached per host thread and device and only
  // grow; creating and destroying them per call costs more than the copies they overlap
  HostPipe &hp = tl_pipe;
  cudaError_t e = cudaSuccess;
#define CUP(call) do { e = (call); if (e != cudaSuccess) { hp.release(); return cuda_fail(e, #call); } } while (0)
  {
    int dev = 0;
    CUP(cudaGetDevice(&dev));
    if (hp.device != dev) { hp.release(); hp.device = dev; }
    if (!hp.s_in) {
      CUP(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
      CUP(cudaStreamCreateWithFlags(&hp.s_cmp, cudaStreamNonBlocking));
      CUP(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
      for (int b = 0; b < 2; b++) {
        CUP(cudaEventCreateWithFlags(&hp.ev_in[b], cudaEventDisableTiming));
        CUP(cudaEventCreateWithFlags(&hp.ev_cmp[b], cudaEventDisableTiming));
        CUP(cudaEventCreateWithFlags(&hp.ev_out[b], cudaEventDisableTiming));
      }
    }
    const size_t need_pcm = std::min(chunk_samples, n_samples) * sizeof(int16_t);
    const size_t need_out = chunk_bound ? chunk_bound : 1;
    const size_t need_ws = encode_ws_bytes(std::min(chunk_frames, nf));
    for (int b = 0; b < 2; b++) {
      if (hp.cap_pcm[b] < need_pcm) { if (hp.d_pcm[b]) cudaFree(hp.d_pcm[b]); hp.d_pcm[b] = nullptr; hp.cap_pcm[b] = 0; CUP(cudaMalloc(&hp.d_pcm[b], need_pcm)); hp.cap_pcm[b] = need_pcm; }
      if (hp.cap_out[b] < need_out) { if (hp.d_out[b]) cudaFree(hp.d_out[b]); hp.d_out[b] = nullptr; hp.cap_out[b] = 0; CUP(cudaMalloc(&hp.d_out[b], need_out)); hp.cap_out[b] = need_out; }
      if (hp.cap_ws[b] < need_ws) { if (hp.ws[b]) cudaFree(hp.ws[b]); hp.ws[b] = nullptr; hp.cap_ws[b] = 0; CUP(cudaMalloc(&hp.ws[b], need_ws)); hp.cap_ws[b] = need_ws; }
    }
    if (hp.cap_res < (size_t)nchunks * 64) {
      if (hp.h_res) cudaFreeHost(hp.h_res);
      hp.h_res = nullptr; hp.cap_res = 0;
      CUP(cudaMallocHost(&hp.h_res, (size_t)nchunks * 64));
      hp.cap_res = (size_t)nchunks * 64;
    }
  }
  cudaStream_t s_in = hp.s_in, s_cmp = hp.s_cmp, s_out = hp.s_out;
  cudaEvent_t *ev_in = hp.ev_in, *ev_cmp = hp.ev_cmp, *ev_out = hp.ev_out;
  int16_t **d_pcm = hp.d_pcm;
  uint8_t **d_out = hp.d_out;
  unsigned char **ws = hp.ws;
  unsigned long long *h_res = hp.h_res;
  auto cleanup = [&]() {};

  size_t off = 0;
  bool overflow = false;
  unsigned long long st6[6] = {0, 0, 0, 0, 0, 0};
  // D2H of a finished chunk (its size is known only after its kernel): waits for the kernel, then queues the copy
  auto finish = [&](unsigned long long c) -> cudaError_t {
    const int b = (int)(c & 1);
    cudaError_t ee = cudaEventSynchronize(ev_cmp[b]);
    if (ee != cudaSuccess) return ee;
    const unsigned long long *r = h_res + c * 8;
    for (int k = 0; k < 6; k++) st6[k] += r[2 + k];
    const size_t len = (size_t)r[0];
    if (r[1] || off + len > out_cap) { overflow = true; off += len; return cudaEventRecord(ev_out[b], s_out); }
    if (len) ee = cudaMemcpyAsync(out + off, d_out[b], len, cudaMemcpyDeviceToHost, s_out);
    off += len;
    if (ee == cudaSuccess) ee = cudaEventRecord(ev_out[b], s_out);
    return ee;
  };
  for (unsigned long long c = 0; c < nchunks; c++) {
    const int b = (int)(c & 1);
    const size_t s0 = (size_t)c * chunk_samples;
    const size_t ns = std::min(chunk_samples, n_samples - s0);
    if (c >= 2) CUP(cudaStreamWaitEvent(s_in, ev_cmp[b], 0));   // d_pcm[b] is free once chunk c-2 was encoded
    CUP(cudaMemcpyAsync(d_pcm[b], pcm + s0, ns * sizeof(int16_t), cudaMemcpyHostToDevice, s_in));
    CUP(cudaEventRecord(ev_in[b], s_in));
    CUP(cudaStreamWaitEvent(s_cmp, ev_in[b], 0));
    if (c >= 2) CUP(cudaStreamWaitEvent(s_cmp, ev_out[b], 0));  // d_out[b] is free once chunk c-2 was copied out
    int erc = X3_OK;
    CUP(enqueue_encode(d, ds, d_pcm[b], ns, d_out[b], chunk_bound, ws[b], s_cmp, &erc));
    if (erc) { cudaStreamSynchronize(s_in); cudaStreamSynchronize(s_cmp); cudaStreamSynchronize(s_out); return erc; }
    CUP(cudaMemcpyAsync(h_res + c * 8, ws[b], 64, cudaMemcpyDeviceToHost, s_cmp));
    CUP(cudaEventRecord(ev_cmp[b], s_cmp));
    if (c >= 1) CUP(finish(c - 1));
  }
  CUP(finish(nchunks - 1));
  CUP(cudaStreamSynchronize(s_out));
#undef CUP
  cleanup();
  *out_len = off;
  if (stats)
    for (int k = 0; k < 6; k++) stats->samples_by_mode[k] = st6[k];
  return overflow ? X3_ERR_BYTEWRITER_INSUFFICIENT_MEMORY : X3_OK;
}

int x3_encode_frame_host(const int16_t *pcm, size_t n_samples, const x3_params *p, uint8_t *out, size_t out_cap,
                         size_t *out_len, x3_stats *stats) {
  if (!p) return X3_ERR_INVALID_ARGUMENT;
  if (n_samples == 0) return X3_ERR_REFERENCE_PANIC;  // wav[0], encoder.rs:189
  if (n_samples > 65535) return X3_ERR_UNSUPPORTED_PARAMS;
  // one frame = encode() with a frame length that covers the whole input
  x3_params q = *p;
  if (q.block_len == 0) return X3_ERR_INVALID_ARGUMENT;
  q.blocks_per_frame = (uint32_t)((n_samples + q.block_len - 1) / q.block_len);   // may round up past 65535 samples
  tl_one_frame = true;
  const int rc = x3_encode_host(pcm, n_samples, &q, out, out_cap, out_len, stats);
  tl_one_frame = false;
  return rc;
}

size_t x3_encode_frame_bound(size_t n_samples, const x3_params *p) {
  if (!p || p->block_len == 0 || n_samples == 0 || n_samples > 65535) return 0;
  x3_params q = *p;
  q.blocks_per_frame = (uint32_t)((n_samples + q.block_len - 1) / q.block_len);
  tl_one_frame = true;
  const size_t b = x3_encode_bound(n_samples, &q);
  tl_one_frame = false;
  return b;
}

// ------------------------------------------------------------------------------------------------
// frame-range sharding helpers (no GPU needed except the copy)
// ------------------------------------------------------------------------------------------------
int x3_shard_range(uint64_t n_samples, const x3_params *p, uint32_t rank, uint32_t world, uint64_t *s0, uint64_t *s1) {
  if (!p || !s0 || !s1 || world == 0 || rank >= world) return X3_ERR_INVALID_ARGUMENT;
  const uint64_t spf = (uint64_t)p->block_len * p->blocks_per_frame;
  if (spf == 0) return X3_ERR_INVALID_ARGUMENT;
  const uint64_t frames = (n_samples + spf - 1) / spf;
  const unsigned __int128 f0 = (unsigned __int128)frames * rank / world, f1 = (unsigned __int128)frames * (rank + 1u) / world;
  const uint64_t a = (uint64_t)f0 * spf, b = (uint64_t)f1 * spf;
  *s0 = a < n_samples ? a : n_samples;
  *s1 = b < n_samples ? b : n_samples;
  return X3_OK;
}

int x3_deal_files(const uint64_t *frames_per_file, size_t n_files, uint32_t world, uint32_t *rank_of_file) {
  if ((!frames_per_file || !rank_of_file) && n_files) return X3_ERR_INVALID_ARGUMENT;
  if (world == 0) return X3_ERR_INVALID_ARGUMENT;
  unsigned __int128 total = 0, acc = 0;
  for (size_t i = 0; i < n_files; i++) total += frames_per_file[i];
  for (size_t i = 0; i < n_files; i++) {
    uint64_t r = total ? (uint64_t)(acc * world / total) : 0;   // the rank whose equal share holds the file's first frame
    if (r > world - 1u) r = world - 1u;
    rank_of_file[i] = (uint32_t)r;
    acc += frames_per_file[i];
  }
  return X3_OK;
}

int x3_shard_base(const uint64_t *shard_bytes, uint32_t world, uint32_t rank, uint64_t *base) {
  if (!shard_bytes || !base || rank >= world) return X3_ERR_INVALID_ARGUMENT;
  uint64_t b = 0;
  for (uint32_t r = 0; r < rank; r++) b += shard_bytes[r];
  *base = b;
  return X3_OK;
}

int x3_place_shard_device(uint8_t *d_stream, uint64_t base, const uint8_t *d_shard, size_t shard_bytes, void *cuda_stream) {
  if ((!d_stream || !d_shard) && shard_bytes) return X3_ERR_INVALID_ARGUMENT;
  if (shard_bytes == 0) return X3_OK;
  CU(cudaMemcpyAsync(d_stream + base, d_shard, shard_bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
  return X3_OK;
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
namespace {

// The sequential header walk of decodefile.rs:105-126 on a host copy of the stream.
// Returns the frames to decode and the condition that ended the walk (0 = clean stop).
int host_walk(const uint8_t *s, size_t len, std::vector<FrameRec> &frames, unsigned long long *total_samples) {
  size_t cursor = 0, remaining = len;
  unsigned long long samples = 0;
  for (;;) {
    if (remaining <= 20) return X3_OK;                                 // :107-109
    if (cursor + 20 > len) return X3_ERR_IO;
    remaining -= 20;
    x3_frame_header h;
    int rc = x3_read_frame_header(s + cursor, 20, &h);                 // :112
    cursor += 20;
    if (rc) return rc;
    if (remaining < h.payload_len) return X3_OK;                       // :114-116
    FrameRec fr;
    fr.pos = cursor - 20;
    fr.out_off = samples;
    fr.samples = h.samples;
    fr.payload_len = h.payload_len;
    fr.payload_crc = h.payload_crc;
    fr.pad = 0;
    frames.push_back(fr);
    samples += h.samples;
    *total_samples = samples;
    if (h.payload_len > tl_max_payload) return X3_OK;  // the frame itself reports InvalidPayloadLen (crc kernel)
    remaining -= h.payload_len;
    cursor += h.payload_len;
  }
}

}  // namespace

namespace {
// x3_decode_device; `consumed` (optional) receives the stream bytes covered by the frames found when the device
// index was used and proven (0 otherwise)
constexpr int kRetryLargerTable = 1000;   // internal: the frame table / tile candidate capacity was too small
int decode_device_attempt(const uint8_t *d_frames, size_t len, const x3_params *p, int16_t *d_pcm, size_t pcm_cap,
                          size_t *n_out, x3_decode_result *res, void *cuda_stream, unsigned long long *consumed,
                          uint32_t tile_bytes, bool hop, unsigned long long max_frames);

// First attempt: the hop index (64 KiB tiles, at most 64 frames in a tile) with a frame table sized for frames of 256
// bytes and more (the default frame is ~4.7 KB).  A stream of smaller frames (small blocks_per_frame: the reference
// accepts any, decoder.rs:49-55) overflows one or the other; it is then indexed by the scan kernel with 16 KiB tiles
// and a table for the smallest frame there is (20 + 2 bytes).  X3_INDEX=scan skips the hop index (for comparison).
static bool use_scan_index() {
  static const bool v = [] { const char *e = getenv("X3_INDEX"); return e && strcmp(e, "scan") == 0; }();
  return v;
}
static uint32_t hop_tile_bytes() {  // X3_HOP_TILE=<KiB> for tuning runs
  static const uint32_t v = [] {
    const char *e = getenv("X3_HOP_TILE");
    const long k = e ? atol(e) : 0;
    return k >= 1 && k <= 1024 ? (uint32_t)k * 1024u : kHopTileBytes;
  }();
  return v;
}
int decode_device_impl(const uint8_t *d_frames, size_t len, const x3_params *p, int16_t *d_pcm, size_t pcm_cap,
                       size_t *n_out, x3_decode_result *res, void *cuda_stream, unsigned long long *consumed) {
  unsigned long long max_frames = len / 256 + 4096;
  const unsigned long long most = len / 22 + 1;
  if (max_frames > most) max_frames = most;
  int rc = decode_device_attempt(d_frames, len, p, d_pcm, pcm_cap, n_out, res, cuda_stream, consumed,
                                 use_scan_index() ? kScanTileBytes : hop_tile_bytes(), !use_scan_index(), max_frames);
  if (rc == kRetryLargerTable)
    rc = decode_device_attempt(d_frames, len, p, d_pcm, pcm_cap, n_out, res, cuda_stream, consumed, kScanTileBytesSmall, false, most);
  return rc;
}

int decode_device_attempt(const uint8_t *d_frames, size_t len, const x3_params *p, int16_t *d_pcm, size_t pcm_cap,
                          size_t *n_out, x3_decode_result *res, void *cuda_stream, unsigned long long *consumed,
                          uint32_t tile_bytes, bool hop, unsigned long long max_frames) {
  if (consumed) *consumed = 0;
  Derived d;
  int rc = derive(p, &d);
  if (rc == X3_ERR_UNSUPPORTED_PARAMS) {
    // decode does not depend on the encoder-side limits (frame length comes from the headers)
    if (!p || p->block_len == 0) return X3_ERR_INVALID_ARGUMENT;
    d.P.block_len = p->block_len;
    d.P.spf = 0;
    for (int k = 0; k < 3; k++) { d.P.codes[k] = p->codes[k]; d.P.thresholds[k] = p->thresholds[k]; }
    rc = X3_OK;
  }
  if (rc) return rc;
  if (!n_out) return X3_ERR_INVALID_ARGUMENT;
  *n_out = 0;
  x3_decode_result local;
  if (!res) res = &local;
  memset(res, 0, sizeof *res);
  res->first_bad_frame = UINT64_MAX;
  tl_ms[0] = tl_ms[1] = tl_ms[2] = tl_ms[3] = 0.f;
  if (len <= 20) return X3_OK;  // decodefile.rs:107-109
  if (!d_frames || (!d_pcm && pcm_cap)) return X3_ERR_INVALID_ARGUMENT;
  if (((uintptr_t)d_pcm & 1u) != 0) return X3_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  DeviceState *ds;
  if ((rc = device_state(&ds))) return rc;
  unsigned long long *host_res;
  if ((rc = pinned(&host_res))) return rc;

  const uint32_t n_tiles = (uint32_t)((len + tile_bytes - 1) / tile_bytes);
  const bool can_retry = hop || tile_bytes != kScanTileBytesSmall;
  // workspace layout
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_res = take(64), o_dres = take(64), o_tick = take(64);
  const size_t zero_bytes = off;
  const size_t o_tiles = take(8 * (size_t)n_tiles), o_trecs = take(8 * (size_t)n_tiles);
  const size_t o_frames = take(sizeof(FrameRec) * max_frames), o_fstat = take(sizeof(int) * max_frames);
  const size_t o_cstat = take(sizeof(int) * max_frames);
  const size_t o_recs = take(sizeof(FrameRec) * max_frames);
  unsigned char *ws = nullptr;
  CU(cudaMallocAsync(&ws, off, st));
  cudaError_t e = cudaMemsetAsync(ws, 0, zero_bytes, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws + o_dres, 0xff, 8, st);
  auto fail = [&](cudaError_t ee, const char *what) { cudaFreeAsync(ws, st); return cuda_fail(ee, what); };
  if (e != cudaSuccess) return fail(e, "cudaMemsetAsync");

  ScanArgs sa;
  sa.stream = d_frames;
  sa.stream_len = len;
  sa.frames = reinterpret_cast<FrameRec *>(ws + o_frames);
  sa.max_frames = max_frames;
  sa.recs = reinterpret_cast<FrameRec *>(ws + o_recs);
  sa.tile_status = reinterpret_cast<unsigned long long *>(ws + o_tiles);
  sa.ticket = reinterpret_cast<unsigned int *>(ws + o_tick);
  sa.rec_cursor = reinterpret_cast<unsigned long long *>(ws + o_tick + 8);
  sa.tile_recs = reinterpret_cast<unsigned long long *>(ws + o_trecs);
  sa.result = reinterpret_cast<unsigned long long *>(ws + o_res);
  sa.crc_tables = ds->crc_dev;
  sa.n_tiles = n_tiles;
  sa.tile_bytes = tile_bytes;
  sa.len_dev = nullptr;

  DecodeArgs da;
  da.stream = d_frames;
  da.stream_len = len;
  da.pcm = d_pcm;
  da.pcm_cap = pcm_cap;
  da.P = d.P;
  da.frames = sa.frames;
  da.n_frames = sa.result + 4;
  da.max_frames = max_frames;
  da.frame_status = reinterpret_cast<int *>(ws + o_fstat);
  da.crc_status = reinterpret_cast<int *>(ws + o_cstat);
  da.one = 1u;
  da.max_payload = tl_max_payload;
  da.result = reinterpret_cast<unsigned long long *>(ws + o_dres);
  da.crc_tables = ds->crc_dev;
  da.len_dev = nullptr;

  cudaStream_t s2 = fork_stream();
  if (!s2) s2 = st;  // no second stream: the two kernels simply run one after the other
  release_pooled_events();
  Timer t_all(st), t_idx(st), t_crc(s2), t_dec(st);
  // crc_frames on s2 beside decode_frames on st; st continues only when both are done
  auto crc_and_decode = [&](unsigned long long hint) -> cudaError_t {
    cudaError_t ee = cudaSuccess;
    if (s2 != st) {
      ee = cudaEventRecord(tl_fork.fork, st);
      if (ee == cudaSuccess) ee = cudaStreamWaitEvent(s2, tl_fork.fork, 0);
      if (ee != cudaSuccess) return ee;
    }
    // the decode kernel first: its CTAs take their places, the CRC kernel's CTAs get what is left and then every
    // slot a finished decode CTA frees
    static const bool crc_first = [] { const char *e = getenv("X3_CRC_FIRST"); return e && e[0] == '1'; }();
    if (crc_first) {
      t_crc.start();
      ee = launch_crc(da, hint, s2);
      t_crc.stop();
      if (ee != cudaSuccess) return ee;
    }
    t_dec.start();
    ee = launch_decode(da, hint, st);
    t_dec.stop();
    if (ee != cudaSuccess) return ee;
    if (!crc_first) {
      t_crc.start();
      ee = launch_crc(da, hint, s2);
      t_crc.stop();
    }
    g_launches += 2;
    if (ee != cudaSuccess) return ee;
    if (s2 != st) {
      ee = cudaEventRecord(tl_fork.join, s2);
      if (ee == cudaSuccess) ee = cudaStreamWaitEvent(st, tl_fork.join, 0);
    }
    return ee;
  };
  bool need_walk = (((uintptr_t)d_frames) & 15u) != 0;  // the scan uses 16-byte loads from the stream base
  int walk_rc = X3_OK;
  unsigned long long total_samples = 0, n_frames = 0;
  t_all.start();
  if (!need_walk) {
    t_idx.start();
    e = launch_scan(sa, hop, st);
    if (e == cudaSuccess) e = launch_chain_check(sa, st);
    t_idx.stop();
    g_launches += 4;
    if (e != cudaSuccess) return fail(e, "scan_headers_kernel");
    e = crc_and_decode(max_frames);
    if (e != cudaSuccess) return fail(e, "crc_frames_kernel / decode_frames_kernel");
    t_all.stop();
    e = cudaMemcpyAsync(host_res, ws + o_res, 64, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_res + 8, ws + o_dres, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(e, "decode (device index)");
    if ((host_res[2] & 2ull) != 0 && can_retry) {  // a capacity was exceeded: smaller tiles, larger table
      cudaFreeAsync(ws, st);
      return kRetryLargerTable;
    }
    if (host_res[2] != 0) need_walk = true;  // the table could not be proven equal to the reference's walk
    n_frames = host_res[4];
    total_samples = host_res[5];
    if (consumed && !need_walk) *consumed = host_res[6];
    tl_ms[0] = t_dec.ms();
    tl_ms[1] = t_idx.ms();
    tl_ms[2] = t_all.ms();
    tl_ms[3] = t_crc.ms();
  }
  if (need_walk) {
    // sequential host walk (rare: corrupt / foreign / unaligned streams)
    res->used_host_walk = 1;
    std::vector<uint8_t> host(len);
    e = cudaMemcpyAsync(host.data(), d_frames, len, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(e, "host walk copy");
    std::vector<FrameRec> frames;
    total_samples = 0;
    walk_rc = host_walk(host.data(), len, frames, &total_samples);
    n_frames = frames.size();
    if (n_frames > max_frames) {
      cudaFreeAsync(ws, st);
      return can_retry ? kRetryLargerTable : X3_ERR_UNSUPPORTED_PARAMS;   // the retry's table holds any stream
    }
    host_res[4] = n_frames;
    e = cudaMemcpyAsync(ws + o_res + 32, host_res + 4, 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_frames)
      e = cudaMemcpyAsync(ws + o_frames, frames.data(), sizeof(FrameRec) * n_frames, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws + o_dres, 0xff, 8, st);
    if (e == cudaSuccess && n_frames) e = crc_and_decode(n_frames);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_res + 8, ws + o_dres, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(e, "decode (host walk)");
    if (n_frames) { tl_ms[0] = t_dec.ms(); tl_ms[3] = t_crc.ms(); tl_ms[2] = tl_ms[0] + tl_ms[3]; }
  }

  // ---- what the reference would have reported ----
  const unsigned long long first_bad = host_res[8];
  int ret = walk_rc;
  unsigned long long good_frames = n_frames, good_samples = total_samples;
  if (first_bad != ~0ull && first_bad < n_frames) {
    int fstat = 0, cstat = 0;
    FrameRec fr;
    e = cudaMemcpyAsync(&fstat, ws + o_fstat + sizeof(int) * first_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(&cstat, ws + o_cstat + sizeof(int) * first_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(&fr, ws + o_frames + sizeof(FrameRec) * first_bad, sizeof(FrameRec), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(e, "decode status read-back");
    if (cstat != kDecOk) fstat = cstat;  // the reference checks the payload CRC before it decodes (decodefile.rs:93-103)
    good_frames = first_bad;
    good_samples = fr.out_off;
    res->first_bad_frame = first_bad;
    res->first_bad_code = map_frame_error(fstat);
    if (fstat == kDecErrOutOfBoundsInverse || fstat == kDecErrInvalidBpf) {
      res->frame_errors = 1;  // decodefile.rs:130-134: counted, printed, decode stops, Ok(None)
      ret = X3_OK;
    } else {
      ret = map_frame_error(fstat);  // payload CRC / payload length / capacity: propagated as Err
    }
  }
  cudaFreeAsync(ws, st);
  res->frames = good_frames;
  res->samples = good_samples;
  *n_out = (size_t)good_samples;
  return ret;
}

// ---- pipelined host decode ----------------------------------------------------------------------
// The stream is cut into a few pieces of geometrically growing size at frame starts GUESSED on the host (key 'x3'
// + valid header at an even offset).  Pieces are uploaded back to back on one stream; each is indexed and decoded
// as a stream of its own as soon as it has arrived, and its PCM goes back on a third stream while the next piece
// is still uploading / decoding -- PCIe runs in both directions.  A guess is proven afterwards: piece j is the
// reference's walk of its bytes (device chain check), it must decode without any error and must consume exactly
// its byte range, i.e. end where piece j+1 starts.  Piece 0 starts at 0, so by induction every piece starts at a
// true frame start.  If anything is off -- a bad frame, a piece that needed the host walk, a byte left over -- the
// attempt is dropped and the whole stream is decoded by the plain path, which reports what the reference would.
struct DecPipe {
  int device = -1;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
thread_local DecPipe tl_dec;
// tunables (read on every call so that tests can exercise the pipeline on small streams):
// X3_DEC_PIPE_MIN_MB -- streams shorter than this take the plain path (default 48);
// X3_DEC_PIPE_FIRST_KB -- size of the first piece (default 8192); each following piece is 4x larger, because the
// PCM of a piece is ~4x its bytes, so its download hides the next piece's upload
size_t env_size(const char *name, size_t dflt, unsigned shift) {
  const char *s = getenv(name);
  return s && *s ? (size_t)strtoull(s, nullptr, 10) << shift : dflt;
}
size_t decode_pipe_min() { return env_size("X3_DEC_PIPE_MIN_MB", (size_t)48 << 20, 20); }
size_t decode_pipe_first() {
  const size_t v = env_size("X3_DEC_PIPE_FIRST_KB", (size_t)8 << 20, 10);
  return v < 4096 ? 4096 : v;
}

// first even offset >= from at which a plausible frame header starts (validated later), or len if none
size_t guess_frame_start(const uint8_t *s, size_t len, size_t from) {
  for (size_t pos = (from + 1) & ~(size_t)1; pos + 20 <= len; pos += 2) {
    if (s[pos] != 'x' || s[pos + 1] != '3') continue;
    x3_frame_header h;
    if (x3_read_frame_header(s + pos, 20, &h) != X3_OK) continue;
    const size_t next = pos + 20 + h.payload_len;
    if (next + 20 <= len) {  // the following header must be valid too (cheap filter; the proof comes later)
      x3_frame_header h2;
      if (x3_read_frame_header(s + next, 20, &h2) != X3_OK) continue;
    }
    return pos;
  }
  return len;
}

// returns 1 if the pipelined attempt produced the final result (in *rc_out), 0 if the caller must use the plain path
int decode_host_pipelined(const uint8_t *frames, size_t len, const x3_params *p, int16_t *pcm, size_t pcm_cap,
                          size_t *n_out, x3_decode_result *res, int *rc_out) {
  size_t cut[9];
  int np = 0;
  cut[0] = 0;
  size_t want = decode_pipe_first();
  while (np < 7) {
    const size_t target = cut[np] + want;
    if (target + (want >> 1) >= len) break;
    const size_t c = guess_frame_start(frames, len, target);
    if (c >= len || c <= cut[np]) break;
    cut[++np] = c;
    want *= 4;
  }
  cut[++np] = len;  // pieces [cut[j], cut[j+1]), j = 0..np-1
  if (np < 2) return 0;

  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  DecPipe &dp = tl_dec;
  if (dp.device != dev || !dp.s_in) {
    dp.device = dev;
    if (cudaStreamCreateWithFlags(&dp.s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&dp.s_cmp, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&dp.s_out, cudaStreamNonBlocking) != cudaSuccess) {
      dp.s_in = nullptr;
      return 0;
    }
    for (int i = 0; i < 8; i++)
      if (cudaEventCreateWithFlags(&dp.ev[i], cudaEventDisableTiming) != cudaSuccess) { dp.s_in = nullptr; return 0; }
  }
  // device placement: every piece starts at a 256-byte aligned address (the index kernel loads 16 bytes at a time)
  size_t doff[9], total = 0;
  for (int j = 0; j < np; j++) { doff[j] = total; total += ((cut[j + 1] - cut[j]) + 16 + 255) & ~(size_t)255; }
  uint8_t *d_in = nullptr;
  int16_t *d_pcm = nullptr;
  if (cudaMallocAsync(&d_in, total, dp.s_cmp) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaMallocAsync(&d_pcm, (pcm_cap ? pcm_cap : 1) * sizeof(int16_t), dp.s_cmp) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeAsync(d_in, dp.s_cmp);
    return 0;
  }
  cudaEvent_t ev_alloc = dp.ev[7];
  cudaError_t e = cudaEventRecord(ev_alloc, dp.s_cmp);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(dp.s_in, ev_alloc, 0);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(dp.s_out, ev_alloc, 0);
  for (int j = 0; j < np && e == cudaSuccess; j++) {
    e = cudaMemcpyAsync(d_in + doff[j], frames + cut[j], cut[j + 1] - cut[j], cudaMemcpyHostToDevice, dp.s_in);
    if (e == cudaSuccess) e = cudaEventRecord(dp.ev[j], dp.s_in);
  }
  bool clean = e == cudaSuccess;
  unsigned long long samples = 0, nframes = 0;
  float ms[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < np && clean; j++) {
    const size_t plen = cut[j + 1] - cut[j];
    if (cudaStreamWaitEvent(dp.s_cmp, dp.ev[j], 0) != cudaSuccess) { clean = false; break; }
    size_t got = 0;
    x3_decode_result r;
    unsigned long long consumed = 0;
    const int rc = decode_device_impl(d_in + doff[j], plen, p, d_pcm + samples, pcm_cap - samples, &got, &r, dp.s_cmp, &consumed);
    for (int k = 0; k < 4; k++) ms[k] += tl_ms[k];
    // interior pieces must be consumed to the last byte; the last one may end in the <= 20 stray bytes or the
    // truncated frame the reference tolerates -- but then it is the plain path that reports it
    if (rc != X3_OK || r.first_bad_frame != UINT64_MAX || r.used_host_walk || consumed != plen) { clean = false; break; }
    if (got && cudaMemcpyAsync(pcm + samples, d_pcm + samples, got * sizeof(int16_t), cudaMemcpyDeviceToHost, dp.s_out) != cudaSuccess) {
      clean = false;
      break;
    }
    samples += got;
    nframes += r.frames;
  }
  e = cudaStreamSynchronize(dp.s_out);
  const cudaError_t e2 = cudaStreamSynchronize(dp.s_in);
  cudaFreeAsync(d_in, dp.s_cmp);
  cudaFreeAsync(d_pcm, dp.s_cmp);
  if (e != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); return 0; }
  if (!clean) return 0;
  if (res) {
    memset(res, 0, sizeof *res);
    res->first_bad_frame = UINT64_MAX;
    res->frames = nframes;
    res->samples = samples;
  }
  for (int k = 0; k < 4; k++) tl_ms[k] = ms[k];
  *n_out = (size_t)samples;
  *rc_out = X3_OK;
  return 1;
}
}  // namespace

int x3_decode_device(const uint8_t *d_frames, size_t len, const x3_params *p, int16_t *d_pcm, size_t pcm_cap,
                     size_t *n_out, x3_decode_result *res, void *cuda_stream) {
  return decode_device_impl(d_frames, len, p, d_pcm, pcm_cap, n_out, res, cuda_stream, nullptr);
}

int x3_decode_host(const uint8_t *frames, size_t len, const x3_params *p, int16_t *pcm, size_t pcm_cap, size_t *n_out,
                   x3_decode_result *res) {
  if (!n_out) return X3_ERR_INVALID_ARGUMENT;
  *n_out = 0;
  if (len <= 20) {
    if (res) { memset(res, 0, sizeof *res); res->first_bad_frame = UINT64_MAX; }
    return p ? X3_OK : X3_ERR_INVALID_ARGUMENT;
  }
  if (!frames || (!pcm && pcm_cap)) return X3_ERR_INVALID_ARGUMENT;
  cudaStream_t st = nullptr;
  DeviceState *ds;
  int rc = device_state(&ds);
  if (rc) return rc;
  if (len >= decode_pipe_min() && p) {
    int prc = X3_OK;
    if (decode_host_pipelined(frames, len, p, pcm, pcm_cap, n_out, res, &prc)) return prc;
    *n_out = 0;
  }
  uint8_t *d_in = nullptr;
  int16_t *d_pcm = nullptr;
  CU(cudaMallocAsync(&d_in, len + 16, st));
  cudaError_t e = cudaMallocAsync(&d_pcm, (pcm_cap ? pcm_cap : 1) * sizeof(int16_t), st);
  if (e != cudaSuccess) { cudaFreeAsync(d_in, st); return cuda_fail(e, "cudaMallocAsync"); }
  e = cudaMemcpyAsync(d_in, frames, len, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = x3_decode_device(d_in, len, p, d_pcm, pcm_cap, n_out, res, st);
    if (rc != X3_ERR_CUDA && *n_out) {
      e = cudaMemcpyAsync(pcm, d_pcm, *n_out * sizeof(int16_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
  }
  cudaFreeAsync(d_in, st);
  cudaFreeAsync(d_pcm, st);
  if (e != cudaSuccess) return cuda_fail(e, "x3_decode_host copies");
  return rc;
}

int x3_decode_frame_host(const uint8_t *payload, size_t payload_len, const x3_params *p, int16_t *pcm, size_t pcm_cap,
                         size_t samples, size_t *n_out) {
  if (!payload || !p || !n_out) return X3_ERR_INVALID_ARGUMENT;
  *n_out = 0;
  if (payload_len < 2 || samples == 0) return X3_ERR_REFERENCE_PANIC;  // decoder.rs:42,47
  if (payload_len >= kFrameMaxLength || samples > 65535) return X3_ERR_FRAME_LENGTH;
  if (samples > pcm_cap) return X3_ERR_REFERENCE_PANIC;  // slice index out of range, decoder.rs:51
  // wrap the payload in a frame so the stream path can be used
  std::vector<uint8_t> buf(20 + payload_len);
  x3_write_frame_header(samples, 1, payload_len, x3_crc16(payload, payload_len), buf.data());
  memcpy(buf.data() + 20, payload, payload_len);
  x3_decode_result r;
  tl_max_payload = kFrameMaxLength;   // decoder::decode_frame has no X3_READ_BUFFER_SIZE limit (that is X3aReader's)
  int rc = x3_decode_host(buf.data(), buf.size(), p, pcm, pcm_cap, n_out, &r);
  tl_max_payload = kReadBufferSize;
  if (rc == X3_OK && r.frame_errors) return r.first_bad_code;  // decode_frame itself returns the Err
  return rc;
}


// ---- stream-ordered entry points: nothing is read back, nothing synchronises -------------------------------------
int x3_encode_device_async(const int16_t *d_pcm, size_t n_samples, const x3_params *p, uint8_t *d_out, size_t out_cap,
                           x3_device_result *d_res, void *cuda_stream) {
  Derived d;
  int rc = derive(p, &d);
  if (rc) return rc;
  if (!d_res || (n_samples && (!d_pcm || !d_out))) return X3_ERR_INVALID_ARGUMENT;
  if (((uintptr_t)d_pcm & 1u) != 0 || ((uintptr_t)d_res & 7u) != 0) return X3_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CU(cudaMemsetAsync(d_res, 0, sizeof *d_res, st));
  if (n_samples == 0) return X3_OK;
  DeviceState *ds;
  if ((rc = device_state(&ds))) return rc;
  const unsigned long long nf = (n_samples + d.P.spf - 1) / d.P.spf;
  if (nf > 0xfffffff0ull) return X3_ERR_UNSUPPORTED_PARAMS;
  unsigned char *ws = nullptr;
  CU(cudaMallocAsync(&ws, encode_ws_bytes(nf), st));
  cudaError_t e = enqueue_encode(d, ds, d_pcm, n_samples, d_out, out_cap, ws, st, &rc, reinterpret_cast<unsigned long long *>(d_res));
  cudaFreeAsync(ws, st);   // stream-ordered: after the kernels
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "encode_frames_kernel");
  return X3_OK;
}

int x3_decode_device_async(const uint8_t *d_frames, size_t len_cap, const uint64_t *d_len, const x3_params *p, int16_t *d_pcm,
                           size_t pcm_cap, x3_device_result *d_res, void *cuda_stream) {
  Derived d;
  int rc = derive(p, &d);
  if (rc == X3_ERR_UNSUPPORTED_PARAMS) {
    if (!p || p->block_len == 0) return X3_ERR_INVALID_ARGUMENT;
    d.P.block_len = p->block_len;
    d.P.spf = 0;
    for (int k = 0; k < 3; k++) { d.P.codes[k] = p->codes[k]; d.P.thresholds[k] = p->thresholds[k]; }
    rc = X3_OK;
  }
  if (rc) return rc;
  if (!d_res || !d_frames || (!d_pcm && pcm_cap)) return X3_ERR_INVALID_ARGUMENT;
  if (((uintptr_t)d_pcm & 1u) != 0 || ((uintptr_t)d_res & 7u) != 0 || ((uintptr_t)d_len & 7u) != 0) return X3_ERR_INVALID_ARGUMENT;
  if (((uintptr_t)d_frames & 15u) != 0) return X3_ERR_INVALID_ARGUMENT;   // the device index needs a 16-byte aligned base
  cudaStream_t st = (cudaStream_t)cuda_stream;
  DeviceState *ds;
  if ((rc = device_state(&ds))) return rc;
  const uint32_t tile_bytes = hop_tile_bytes();
  unsigned long long max_frames = len_cap / 256 + 4096;
  const unsigned long long most = len_cap / 22 + 1;
  if (max_frames > most) max_frames = most;
  const uint32_t n_tiles = (uint32_t)((len_cap + tile_bytes - 1) / tile_bytes);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_res = take(64), o_dres = take(64), o_tick = take(64);
  const size_t zero_bytes = off;
  const size_t o_tiles = take(8 * (size_t)(n_tiles ? n_tiles : 1)), o_trecs = take(8 * (size_t)(n_tiles ? n_tiles : 1));
  const size_t o_frames = take(sizeof(FrameRec) * max_frames), o_fstat = take(sizeof(int) * max_frames);
  const size_t o_cstat = take(sizeof(int) * max_frames);
  const size_t o_recs = take(sizeof(FrameRec) * max_frames);
  unsigned char *ws = nullptr;
  CU(cudaMallocAsync(&ws, off, st));
  cudaError_t e = cudaMemsetAsync(ws, 0, zero_bytes, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws + o_dres, 0xff, 8, st);
  auto fail = [&](cudaError_t ee, const char *what) { cudaFreeAsync(ws, st); return cuda_fail(ee, what); };
  if (e != cudaSuccess) return fail(e, "cudaMemsetAsync");
  ScanArgs sa;
  sa.stream = d_frames;
  sa.stream_len = len_cap;
  sa.frames = reinterpret_cast<FrameRec *>(ws + o_frames);
  sa.max_frames = max_frames;
  sa.recs = reinterpret_cast<FrameRec *>(ws + o_recs);
  sa.tile_status = reinterpret_cast<unsigned long long *>(ws + o_tiles);
  sa.ticket = reinterpret_cast<unsigned int *>(ws + o_tick);
  sa.rec_cursor = reinterpret_cast<unsigned long long *>(ws + o_tick + 8);
  sa.tile_recs = reinterpret_cast<unsigned long long *>(ws + o_trecs);
  sa.result = reinterpret_cast<unsigned long long *>(ws + o_res);
  sa.crc_tables = ds->crc_dev;
  sa.n_tiles = n_tiles;
  sa.tile_bytes = tile_bytes;
  sa.len_dev = reinterpret_cast<const unsigned long long *>(d_len);
  DecodeArgs da;
  da.stream = d_frames;
  da.stream_len = len_cap;
  da.pcm = d_pcm;
  da.pcm_cap = pcm_cap;
  da.P = d.P;
  da.frames = sa.frames;
  da.n_frames = sa.result + 4;
  da.max_frames = max_frames;
  da.frame_status = reinterpret_cast<int *>(ws + o_fstat);
  da.crc_status = reinterpret_cast<int *>(ws + o_cstat);
  da.one = 1u;
  da.max_payload = kReadBufferSize;
  da.result = reinterpret_cast<unsigned long long *>(ws + o_dres);
  da.crc_tables = ds->crc_dev;
  da.len_dev = sa.len_dev;
  e = launch_scan(sa, true, st);
  if (e == cudaSuccess) e = launch_chain_check(sa, st);
  g_launches += 4;
  if (e != cudaSuccess) return fail(e, "hop_index_kernel");
  cudaStream_t s2 = fork_stream();
  if (!s2) s2 = st;
  if (s2 != st) {
    e = cudaEventRecord(tl_fork.fork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s2, tl_fork.fork, 0);
  }
  if (e == cudaSuccess) e = launch_decode(da, max_frames, st);
  if (e == cudaSuccess) e = launch_crc(da, max_frames, s2);
  g_launches += 2;
  if (e == cudaSuccess && s2 != st) {
    e = cudaEventRecord(tl_fork.join, s2);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, tl_fork.join, 0);
  }
  if (e == cudaSuccess) e = launch_decode_finalize(sa, da, reinterpret_cast<unsigned long long *>(d_res), st);
  g_launches++;
  if (e != cudaSuccess) return fail(e, "decode kernels");
  cudaFreeAsync(ws, st);
  return X3_OK;
}

int x3_synth_device(int kind, uint32_t seed, uint32_t fs, uint64_t n0, uint64_t count, int16_t *d_out,
                    void *cuda_stream) {
  if ((kind != 1 && kind != 2 && kind != 4) || fs == 0 || (!d_out && count)) return X3_ERR_INVALID_ARGUMENT;
  cudaError_t e = launch_synth(kind, seed, fs, n0, count, d_out, (cudaStream_t)cuda_stream);
  if (e != cudaSuccess) return cuda_fail(e, "synth_kernel");
  g_launches++;
  return X3_OK;
}

}  // extern "C"
