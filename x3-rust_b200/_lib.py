"""ctypes loader for libx3b200.so -- the only compute backend of this package.

There is no CPU fallback: if the library is missing or no CUDA device is usable, calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libx3b200.so")


class x3_params(C.Structure):
    _fields_ = [("block_len", C.c_uint32), ("blocks_per_frame", C.c_uint32),
                ("codes", C.c_uint32 * 3), ("thresholds", C.c_uint32 * 3)]


class x3_stats(C.Structure):
    _fields_ = [("samples_by_mode", C.c_uint64 * 6)]


class x3_frame_header(C.Structure):
    _fields_ = [("source_id", C.c_uint8), ("channels", C.c_uint8), ("samples", C.c_uint16),
                ("payload_len", C.c_uint32), ("payload_crc", C.c_uint16)]


class x3_decode_result(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("frames", C.c_uint64), ("frame_errors", C.c_uint64),
                ("first_bad_frame", C.c_uint64), ("first_bad_code", C.c_int32), ("used_host_walk", C.c_int32)]


# every symbol include/x3_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "x3_abi_version": (C.c_int, []),
    "x3_params_default": (C.c_int, [_P(x3_params)]),
    "x3_params_validate": (C.c_int, [_P(x3_params)]),
    "x3_encode_bound": (C.c_size_t, [C.c_size_t, _P(x3_params)]),
    "x3_encode_frame_bound": (C.c_size_t, [C.c_size_t, _P(x3_params)]),
    "x3_shard_range": (C.c_int, [C.c_uint64, _P(x3_params), C.c_uint32, C.c_uint32, _P(C.c_uint64), _P(C.c_uint64)]),
    "x3_deal_files": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
    "x3_shard_base": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _P(C.c_uint64)]),
    "x3_place_shard_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "x3_strerror": (C.c_char_p, [C.c_int]),
    "x3_last_cuda_error": (C.c_char_p, []),
    "x3_write_frame_header": (C.c_int, [C.c_size_t, C.c_uint8, C.c_size_t, C.c_uint16, C.c_void_p]),
    "x3_read_frame_header": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_frame_header)]),
    "x3_crc16": (C.c_uint16, [C.c_void_p, C.c_size_t]),
    "x3_encode_host": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t,
                                 _P(C.c_size_t), _P(x3_stats)]),
    "x3_encode_device": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t,
                                   _P(C.c_size_t), _P(x3_stats), C.c_void_p]),
    "x3_encode_frame_host": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t,
                                       _P(C.c_size_t), _P(x3_stats)]),
    "x3_decode_host": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t,
                                 _P(C.c_size_t), _P(x3_decode_result)]),
    "x3_decode_device": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t,
                                   _P(C.c_size_t), _P(x3_decode_result), C.c_void_p]),
    "x3_decode_frame_host": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t,
                                       C.c_size_t, _P(C.c_size_t)]),
    "x3_synth_device": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "x3_kernel_launch_count": (C.c_uint64, []),
    "x3_last_kernel_ms": (C.c_int, [_P(C.c_float * 4)]),
    "x3_last_encode_kernel": (C.c_int, []),
    "x3_encode_device_async": (C.c_int, [C.c_void_p, C.c_size_t, _P(x3_params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "x3_decode_device_async": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, _P(x3_params), C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_void_p]),
}

_lib = None


def lib():
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "x3-rust_b200: %s is missing -- build it with `python x3-rust_b200/build.py` "
                "(nvcc, sm_100a).  There is no CPU fallback." % SO_PATH)
        h = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib
