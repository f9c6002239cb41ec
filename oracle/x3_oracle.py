"""ctypes binding of the CPU oracle (oracle/libx3oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libx3oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "x3_oracle.c")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Params(C.Structure):
    _fields_ = [("block_len", C.c_uint32), ("blocks_per_frame", C.c_uint32),
                ("codes", C.c_uint32 * 3), ("thresholds", C.c_uint32 * 3)]

    @classmethod
    def default(cls):
        p = cls()
        lib().x3o_params_default(C.byref(p))
        return p

    @classmethod
    def make(cls, block_len=20, blocks_per_frame=500, codes=(0, 1, 3), thresholds=(3, 8, 20)):
        p = cls()
        p.block_len, p.blocks_per_frame = block_len, blocks_per_frame
        p.codes[:] = codes
        p.thresholds[:] = thresholds
        return p


class FrameHeader(C.Structure):
    _fields_ = [("source_id", C.c_uint8), ("samples", C.c_uint16), ("channels", C.c_uint8),
                ("payload_len", C.c_uint32), ("payload_crc", C.c_uint16)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.x3o_crc16.restype = C.c_uint16
        _lib.x3o_update_crc16.restype = C.c_uint16
        _lib.x3o_inv_rice.restype = C.c_int16
        _lib.x3o_encode_bound.restype = C.c_size_t
        _lib.x3o_encode_bound.argtypes = [C.c_size_t, C.POINTER(Params)]
    return _lib


def _u8(a):
    return np.ascontiguousarray(np.frombuffer(bytes(a), dtype=np.uint8)) if not isinstance(a, np.ndarray) \
        else np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle error %d" % code)
        self.code = code


def crc16(data):
    d = _u8(data)
    return int(lib().x3o_crc16(_p(d, C.c_uint8), C.c_size_t(d.size)))


def rice_table(code_id):
    code = (C.c_uint32 * 56)()
    nbits = (C.c_uint32 * 56)()
    nsubs, off, n, inv_len = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    r = lib().x3o_rice_table(C.c_uint32(code_id), C.byref(nsubs), C.byref(off), C.byref(n), code, nbits,
                             C.byref(inv_len))
    if r:
        raise OracleError(r)
    return dict(nsubs=nsubs.value, offset=off.value, code=list(code[:n.value]),
                num_bits=list(nbits[:n.value]), inv_len=inv_len.value)


def inv_rice(i):
    return int(lib().x3o_inv_rice(C.c_uint32(i)))


def bitpack(pairs, cap):
    vals = np.array([v for v, _ in pairs], dtype=np.uint64)
    nb = np.array([n for _, n in pairs], dtype=np.uint32)
    buf = np.zeros(cap, dtype=np.uint8)
    out_len = C.c_size_t()
    r = lib().x3o_bitpack(_p(vals, C.c_uint64), _p(nb, C.c_uint32), C.c_size_t(len(pairs)),
                          _p(buf, C.c_uint8), C.c_size_t(cap), C.byref(out_len))
    if r:
        raise OracleError(r)
    return buf, out_len.value


def bitread(data, ops):
    """ops: list of ints; n>0 read_nbits(n); 0 count_zero_bits(); -1 report initial state."""
    d = _u8(data)
    o = np.array([0xffffffff if x < 0 else x for x in ops], dtype=np.uint32)
    res = np.zeros(len(ops), dtype=np.uint32)
    lead = np.zeros(len(ops), dtype=np.uint32)
    rem = np.zeros(len(ops), dtype=np.uint32)
    lib().x3o_bitread(_p(d, C.c_uint8), C.c_size_t(d.size), _p(o, C.c_uint32), C.c_size_t(len(ops)),
                      _p(res, C.c_uint32), _p(lead, C.c_uint32), _p(rem, C.c_uint32))
    return [(int(a), int(b), int(c)) for a, b, c in zip(res, lead, rem)]


def write_frame_header(num_samples, id_, payload_len, payload_crc):
    h = (C.c_uint8 * 20)()
    lib().x3o_write_frame_header(C.c_size_t(num_samples), C.c_uint8(id_), C.c_size_t(payload_len),
                                 C.c_uint16(payload_crc), h)
    return bytes(h)


def encode_block_test(wav, params=None, lead_zero_bits=0, cap=8192):
    params = params or Params.default()
    w = np.ascontiguousarray(wav, dtype=np.int16)
    buf = np.zeros(cap, dtype=np.uint8)
    out_len = C.c_size_t()
    r = lib().x3o_encode_block_test(_p(w, C.c_int16), C.c_size_t(w.size), C.byref(params),
                                    C.c_uint32(lead_zero_bits), _p(buf, C.c_uint8), C.c_size_t(cap),
                                    C.byref(out_len))
    if r:
        raise OracleError(r)
    return bytes(buf[:out_len.value])


def encode_frame(wav, params=None, cap=None, pos=0):
    params = params or Params.default()
    w = np.ascontiguousarray(wav, dtype=np.int16)
    cap = cap if cap is not None else 2 * w.size + 64 + pos
    buf = np.zeros(cap, dtype=np.uint8)
    p = C.c_size_t(pos)
    stats = (C.c_uint64 * 6)()
    r = lib().x3o_encode_frame(_p(w, C.c_int16), C.c_size_t(w.size), C.byref(params), _p(buf, C.c_uint8),
                               C.c_size_t(cap), C.byref(p), stats)
    if r:
        raise OracleError(r)
    return bytes(buf[:p.value]), list(stats)


def encode_bound(n, params=None):
    params = params or Params.default()
    return int(lib().x3o_encode_bound(C.c_size_t(n), C.byref(params)))


def encode(wav, params=None, cap=None, threads=1):
    """encoder::encode over a contiguous channel -> (frame bytes as np.uint8 array, stats[6])."""
    params = params or Params.default()
    w = np.ascontiguousarray(wav, dtype=np.int16)
    cap = cap if cap is not None else encode_bound(w.size, params)
    buf = np.empty(cap, dtype=np.uint8)
    stats = (C.c_uint64 * 6)()
    if threads > 1:
        out_len = C.c_size_t()
        r = lib().x3o_encode_mt(_p(w, C.c_int16), C.c_size_t(w.size), C.byref(params), _p(buf, C.c_uint8),
                                C.c_size_t(cap), C.byref(out_len), stats, C.c_int(threads))
        n = out_len.value
    else:
        p = C.c_size_t(0)
        r = lib().x3o_encode(_p(w, C.c_int16), C.c_size_t(w.size), C.byref(params), _p(buf, C.c_uint8),
                             C.c_size_t(cap), C.byref(p), stats)
        n = p.value
    if r:
        raise OracleError(r)
    return buf[:n], list(stats)


def read_frame_header(data):
    d = _u8(data)
    h = FrameHeader()
    r = lib().x3o_read_frame_header(_p(d, C.c_uint8), C.c_size_t(d.size), C.byref(h))
    if r:
        raise OracleError(r)
    return h


def decode_block_test(data, last_wav, block_len, skip_bits=0, params=None):
    params = params or Params.default()
    d = _u8(data)
    wav = np.zeros(block_len, dtype=np.int16)
    r = lib().x3o_decode_block_test(_p(d, C.c_uint8), C.c_size_t(d.size), C.c_uint32(skip_bits),
                                    C.c_int16(last_wav), C.byref(params), _p(wav, C.c_int16),
                                    C.c_size_t(block_len))
    if r:
        raise OracleError(r)
    return wav


def decode_frame(payload, samples, params=None):
    params = params or Params.default()
    d = _u8(payload)
    wav = np.zeros(max(samples, 1), dtype=np.int16)
    n_out = C.c_size_t()
    r = lib().x3o_decode_frame(_p(d, C.c_uint8), C.c_size_t(d.size), _p(wav, C.c_int16), C.c_size_t(wav.size),
                               C.byref(params), C.c_size_t(samples), C.byref(n_out))
    if r:
        raise OracleError(r)
    return wav[:n_out.value]


def decode_stream(data, wav_cap, params=None, remaining0=None, threads=1):
    """Frame loop of decodefile.rs over a bare frame stream.
    Returns (rc, pcm, frames_ok, frame_errors)."""
    params = params or Params.default()
    d = _u8(data)
    wav = np.zeros(max(wav_cap, 1), dtype=np.int16)
    n_out, frames_ok, frame_errors = C.c_size_t(), C.c_size_t(), C.c_size_t()
    if threads > 1:
        r = lib().x3o_decode_stream_mt(_p(d, C.c_uint8), C.c_size_t(d.size), C.byref(params), _p(wav, C.c_int16),
                                       C.c_size_t(wav_cap), C.byref(n_out), C.c_int(threads))
        return r, wav[:n_out.value], None, None
    rem = d.size if remaining0 is None else remaining0
    r = lib().x3o_decode_stream(_p(d, C.c_uint8), C.c_size_t(d.size), C.c_size_t(rem), C.byref(params),
                                _p(wav, C.c_int16), C.c_size_t(wav_cap), C.byref(n_out), C.byref(frames_ok),
                                C.byref(frame_errors))
    return r, wav[:n_out.value], frames_ok.value, frame_errors.value


def archive_header(sample_rate, params=None):
    params = params or Params.default()
    buf = np.zeros(2048, dtype=np.uint8)
    n = C.c_size_t()
    r = lib().x3o_archive_header(C.c_uint32(sample_rate), C.byref(params), _p(buf, C.c_uint8), C.c_size_t(2048),
                                 C.byref(n))
    if r:
        raise OracleError(r)
    return bytes(buf[:n.value])


def x3a_encode(wav, sample_rate):
    w = np.ascontiguousarray(wav, dtype=np.int16)
    cap = encode_bound(w.size) + 1024
    buf = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    stats = (C.c_uint64 * 6)()
    r = lib().x3o_x3a_encode(_p(w, C.c_int16), C.c_size_t(w.size), C.c_uint32(sample_rate), _p(buf, C.c_uint8),
                             C.c_size_t(cap), C.byref(n), stats)
    if r:
        raise OracleError(r)
    return buf[:n.value], list(stats)


def x3a_decode(data, wav_cap):
    d = _u8(data)
    wav = np.zeros(max(wav_cap, 1), dtype=np.int16)
    n_out, fs, frames_ok, frame_errors = C.c_size_t(), C.c_uint32(), C.c_size_t(), C.c_size_t()
    r = lib().x3o_x3a_decode(_p(d, C.c_uint8), C.c_size_t(d.size), _p(wav, C.c_int16), C.c_size_t(wav_cap),
                             C.byref(n_out), C.byref(fs), C.byref(frames_ok), C.byref(frame_errors))
    return r, wav[:n_out.value], fs.value, frames_ok.value, frame_errors.value


def synth(kind, seed, fs, n0, count):
    out = np.empty(count, dtype=np.int16)
    r = lib().x3o_synth(C.c_int(kind), C.c_uint32(seed), C.c_uint32(fs), C.c_uint64(n0), C.c_uint64(count),
                        _p(out, C.c_int16))
    if r:
        raise OracleError(r)
    return out
