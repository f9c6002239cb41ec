"""x3-rust_b200 -- B200-native (sm_100a CUDA) implementation of the X3 lossless audio codec's frame
encode / frame decode path, shaped as a drop-in for the public surface of psiphi75/x3-rust:

    x3.Parameters, x3.Channel, x3.IterChannel        (src/x3.rs)
    encoder.encode / encode_frame / write_frame_header (src/encoder.rs)
    decoder.decode_frame / read_frame_header           (src/decoder.rs)
    encodefile.wav_to_x3a, decodefile.x3a_to_wav / X3aReader

All compute runs in libx3b200.so (x3-rust_b200/csrc, C ABI in include/x3_b200.h).  There is no CPU
fallback: without the built library and a CUDA device the calls raise.
The directory name contains a hyphen; import it with importlib.import_module("x3-rust_b200").
"""
from . import _lib, error  # noqa: F401
from . import x3, bytewriter, encoder, decoder, encodefile, decodefile  # noqa: F401
from .error import X3Error  # noqa: F401
from .encodefile import wav_to_x3a  # noqa: F401
from .decodefile import x3a_to_wav, X3aReader  # noqa: F401

__all__ = ["x3", "bytewriter", "encoder", "decoder", "encodefile", "decodefile", "X3Error", "wav_to_x3a",
           "x3a_to_wav", "X3aReader"]
