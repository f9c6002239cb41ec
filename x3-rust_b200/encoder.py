"""Mirror of the reference's `encoder` module (src/encoder.rs) over the CUDA path.

encode() is encoder::encode (encoder.rs:51-111): it writes frames only (no archive header) at the writer's
current position and prints the mode statistics like the `std` build does (encoder.rs:96-108).
"""
import ctypes as C

import numpy as np

from . import _lib, error, x3

last_stats = [0] * 6   # stats[6] of the most recent encode() (encoder.rs:63), surfaced for callers


def _as_pcm(wav):
    if isinstance(wav, np.ndarray):
        if wav.dtype != np.int16:
            raise error.X3Error(error.INVALID_ARGUMENT, "PCM must be int16")
        return np.ascontiguousarray(wav)
    return np.fromiter(wav, dtype=np.int16) if not isinstance(wav, (bytes, bytearray, memoryview, list, tuple)) \
        else np.ascontiguousarray(np.asarray(wav, dtype=np.int16))


def encode_bound(n_samples, params):
    return int(_lib.lib().x3_encode_bound(n_samples, C.byref(params.c_struct())))


def encode_array(pcm, params, quiet=True):
    """Contiguous int16 samples -> (np.uint8 frame bytes, stats[6]).  Host buffers, GPU compute."""
    L = _lib.lib()
    pcm = _as_pcm(pcm)
    ps = params.c_struct()
    error.check(L.x3_params_validate(C.byref(ps)))
    cap = int(L.x3_encode_bound(pcm.size, C.byref(ps)))
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = C.c_size_t()
    st = _lib.x3_stats()
    error.check(L.x3_encode_host(pcm.ctypes.data, pcm.size, C.byref(ps), out.ctypes.data, cap, C.byref(n), C.byref(st)))
    stats = [int(v) for v in st.samples_by_mode]
    if not quiet:
        print_stats(stats)
    return out[:n.value], stats


def print_stats(stats):
    """encoder.rs:96-108"""
    t = float(sum(stats))
    pct = [(s / t * 100.0) if t else float("nan") for s in stats]
    print("\nStatistics:\n  Rice-0: %.4f%%\n  Rice-1: %.4f%%\n  Rice-2: %.4f%%\n  Rice-3: %.4f%%\n  BFP: %.4f%%\n"
          "  Pass-through %.4f%%\n" % tuple(pct))


def encode(channels, writer, quiet=False):
    """encoder::encode(&mut [&mut IterChannel], &mut W) (encoder.rs:51).  Also accepts x3.Channel (the README's
    slice-backed form, README.md:43-50)."""
    global last_stats
    if len(channels) > 1:
        raise error.X3Error(error.MORE_THAN_ONE_CHANNEL)          # encoder.rs:55-57
    ch = channels[0]
    pcm = _as_pcm(ch.wav)
    if pcm.size == 0:
        last_stats = [0] * 6
        if not quiet:
            print_stats(last_stats)
        return
    writer.align(2)                                               # encode_frame, encoder.rs:182
    data, stats = encode_array(pcm, ch.params)
    left = writer.capacity_left()
    if left is not None and data.size > left:
        raise error.X3Error(error.BYTEWRITER_INSUFFICIENT_MEMORY)  # bytewriter.rs:88-90
    writer.write_all(data.tobytes() if left is None else memoryview(data))
    last_stats = stats
    if not quiet:
        print_stats(stats)


def encode_frame(wav, writer, params, stats):
    """encoder::encode_frame (encoder.rs:175-214): exactly one frame; stats[6] is accumulated in place."""
    L = _lib.lib()
    pcm = _as_pcm(wav)
    ps = params.c_struct()
    writer.align(2)
    cap = max(int(L.x3_encode_frame_bound(pcm.size, C.byref(ps))), 32)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    st = _lib.x3_stats()
    error.check(L.x3_encode_frame_host(pcm.ctypes.data, pcm.size, C.byref(ps), out.ctypes.data, cap, C.byref(n),
                                       C.byref(st)))
    left = writer.capacity_left()
    if left is not None and n.value > left:
        raise error.X3Error(error.BYTEWRITER_INSUFFICIENT_MEMORY)
    writer.write_all(memoryview(out[:n.value]))
    for i in range(6):
        stats[i] += int(st.samples_by_mode[i])


def write_frame_header(num_samples, id, payload_len, payload_crc):
    """encoder::write_frame_header (encoder.rs:122-162) -> 20 bytes."""
    h = (C.c_uint8 * 20)()
    error.check(_lib.lib().x3_write_frame_header(num_samples, id, payload_len, payload_crc, h))
    return bytes(h)
