"""Mirror of the reference's `decoder` module (src/decoder.rs) over the CUDA path."""
import ctypes as C

import numpy as np

from . import _lib, error, x3


def read_frame_header(data):
    """decoder::read_frame_header (decoder.rs:69-118)."""
    b = bytes(data[:20]) if len(data) >= 20 else bytes(data)
    h = _lib.x3_frame_header()
    buf = (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b or b"\0")
    error.check(_lib.lib().x3_read_frame_header(buf, len(b), C.byref(h)))
    return x3.FrameHeader(h.source_id, h.samples, h.channels, h.payload_len, h.payload_crc)


def decode_frame(x3_bytes, wav_buf, params, samples):
    """decoder::decode_frame (decoder.rs:36-58): payload (no header) -> `samples` PCM values in wav_buf.
    Returns the number of samples written (Ok(Some(n)))."""
    pl = np.frombuffer(bytes(x3_bytes), dtype=np.uint8)
    if not (isinstance(wav_buf, np.ndarray) and wav_buf.dtype == np.int16 and wav_buf.flags.c_contiguous):
        raise error.X3Error(error.INVALID_ARGUMENT, "wav_buf must be a contiguous int16 array")
    n = C.c_size_t()
    ps = params.c_struct()
    error.check(_lib.lib().x3_decode_frame_host(pl.ctypes.data, pl.size, C.byref(ps), wav_buf.ctypes.data,
                                                wav_buf.size, samples, C.byref(n)))
    return n.value


class DecodeResult:
    def __init__(self, code, r):
        self.code = code
        self.samples, self.frames, self.frame_errors = int(r.samples), int(r.frames), int(r.frame_errors)
        self.first_bad_frame = None if r.first_bad_frame == 0xFFFFFFFFFFFFFFFF else int(r.first_bad_frame)
        self.first_bad_code = int(r.first_bad_code)
        self.used_host_walk = bool(r.used_host_walk)


def decode_stream(data, params, max_samples=None):
    """The frame loop of decodefile.rs:105-136 / :202-209 over a bare frame stream held in host memory.
    Returns (pcm int16 array, DecodeResult); DecodeResult.code is the error the reference would propagate."""
    L = _lib.lib()
    d = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    if max_samples is None:
        # every frame is at least 22 bytes and carries at most 65535 samples; walk the headers for the real total
        max_samples, pos = 0, 0
        while d.size - pos > 20:
            try:
                h = read_frame_header(d[pos:pos + 20])
            except error.X3Error:
                break
            max_samples += h.samples
            pos += 20 + h.payload_len
    pcm = np.empty(max(max_samples, 1), dtype=np.int16)
    n = C.c_size_t()
    r = _lib.x3_decode_result()
    ps = params.c_struct()
    code = L.x3_decode_host(d.ctypes.data, d.size, C.byref(ps), pcm.ctypes.data, max_samples, C.byref(n), C.byref(r))
    if code in (error.CUDA, error.INVALID_ARGUMENT):
        error.check(code)
    return pcm[:n.value], DecodeResult(code, r)
