"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out (zero-copy over the C ABI).

PyTorch is plumbing only here: it owns the device memory and the stream; all compute is libx3b200.so.
"""
import ctypes as C

import torch

from . import _lib, error, x3


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _on_device:
    """`with torch.cuda.device(d)` only when d is not already current (the context manager costs ~10 us a call,
    which is GPU idle time between two stream-synchronous ABI calls)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def encode_tensor(pcm, params=None, out=None):
    """int16 CUDA tensor -> (uint8 CUDA tensor holding the frame stream, length, stats[6])."""
    params = params or x3.Parameters.default()
    assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()
    L = _lib.lib()
    ps = params.c_struct()
    n = pcm.numel()
    with _on_device(pcm.device):
        if out is None:
            out = torch.empty(max(int(L.x3_encode_bound(n, C.byref(ps))), 1), dtype=torch.uint8, device=pcm.device)
        out_len = C.c_size_t()
        st = _lib.x3_stats()
        error.check(L.x3_encode_device(C.c_void_p(pcm.data_ptr()), n, C.byref(ps), C.c_void_p(out.data_ptr()),
                                       out.numel(), C.byref(out_len), C.byref(st), _stream_ptr()))
    return out, out_len.value, [int(v) for v in st.samples_by_mode]


def encode_tensor_async(pcm, out, result, params=None):
    """Stream-ordered encode (x3_encode_device_async): only enqueues work on the current stream.  `result` is an int64
    CUDA tensor of 8 words (x3_device_result: [0] bytes written, [1] flags, [2..8) samples by mode)."""
    params = params or x3.Parameters.default()
    assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()
    assert result.is_cuda and result.dtype == torch.int64 and result.numel() >= 8 and out.dtype == torch.uint8
    ps = params.c_struct()
    with _on_device(pcm.device):
        error.check(_lib.lib().x3_encode_device_async(C.c_void_p(pcm.data_ptr()), pcm.numel(), C.byref(ps), C.c_void_p(out.data_ptr()),
                                                     out.numel(), C.c_void_p(result.data_ptr()), _stream_ptr()))


def decode_tensor_async(frames, length_dev, out, result, params=None):
    """Stream-ordered decode (x3_decode_device_async): the stream's length is read from device memory (`length_dev`: an
    int64 CUDA tensor, e.g. the encode's result[0:1]); frames.numel() is the upper bound.  `result` as above:
    [0] samples written, [1] flags, [2] frames, [3] first bad frame or -1, [4] its status, [5] bytes consumed."""
    params = params or x3.Parameters.default()
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous() and out.dtype == torch.int16
    assert length_dev.is_cuda and length_dev.dtype == torch.int64 and result.is_cuda and result.dtype == torch.int64
    ps = params.c_struct()
    with _on_device(frames.device):
        error.check(_lib.lib().x3_decode_device_async(C.c_void_p(frames.data_ptr()), frames.numel(), C.c_void_p(length_dev.data_ptr()),
                                                     C.byref(ps), C.c_void_p(out.data_ptr()), out.numel(),
                                                     C.c_void_p(result.data_ptr()), _stream_ptr()))


def decode_tensor(frames, length, params=None, out=None, max_samples=None):
    """uint8 CUDA tensor with a frame stream of `length` bytes -> (int16 CUDA tensor, n_samples, result, code)."""
    params = params or x3.Parameters.default()
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.is_contiguous()
    L = _lib.lib()
    ps = params.c_struct()
    with _on_device(frames.device):
        if out is None:
            assert max_samples is not None
            out = torch.empty(max(max_samples, 1), dtype=torch.int16, device=frames.device)
        n = C.c_size_t()
        r = _lib.x3_decode_result()
        code = L.x3_decode_device(C.c_void_p(frames.data_ptr()), length, C.byref(ps), C.c_void_p(out.data_ptr()),
                                  out.numel(), C.byref(n), C.byref(r), _stream_ptr())
    if code in (error.CUDA, error.INVALID_ARGUMENT):
        error.check(code)
    from .decoder import DecodeResult
    return out, n.value, DecodeResult(code, r), code


def synth(kind, seed, fs, n0, count, device="cuda", out=None):
    """Synthetic signal S1/S2/S4 generated on the device."""
    L = _lib.lib()
    if out is None:
        out = torch.empty(count, dtype=torch.int16, device=device)
    with torch.cuda.device(out.device):
        error.check(L.x3_synth_device(kind, seed, fs, n0, count, C.c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def kernel_launch_count():
    return int(_lib.lib().x3_kernel_launch_count())


def last_kernel_ms():
    ms = (C.c_float * 4)()
    _lib.lib().x3_last_kernel_ms(C.byref(ms))
    return [float(v) for v in ms]
