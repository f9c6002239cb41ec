"""X3Error (error.rs:27-62) as a Python exception carrying the C-ABI code."""
from . import _lib

OK = 0
INVALID_ENCODING_THRESH = -1
OUT_OF_BOUNDS_INVERSE = -2
MORE_THAN_ONE_CHANNEL = -3
ARCHIVE_XML_INVALID = -4
ARCHIVE_XML_RICE_CODE = -5
ARCHIVE_INVALID_KEY = -6
FRAME_LENGTH = -7
FRAME_HEADER_INVALID_KEY = -8
FRAME_HEADER_INVALID_PAYLOAD_LEN = -9
FRAME_HEADER_INVALID_HEADER_CRC = -10
FRAME_HEADER_INVALID_PAYLOAD_CRC = -11
FRAME_DECODE_INVALID_FTYPE = -12
FRAME_DECODE_INVALID_BPF = -13
FRAME_DECODE_UNEXPECTED_END = -14
BYTEWRITER_INSUFFICIENT_MEMORY = -15
IO = -16
INVALID_ARGUMENT = -101
UNSUPPORTED_PARAMS = -102
CUDA = -103
REFERENCE_PANIC = -104


class X3Error(Exception):
    def __init__(self, code, detail=""):
        self.code = int(code)
        msg = _lib.lib().x3_strerror(self.code).decode()
        if self.code == CUDA:
            msg += ": " + _lib.lib().x3_last_cuda_error().decode()
        if detail:
            msg += " (" + detail + ")"
        super().__init__(msg)


def check(code, detail=""):
    if code != OK:
        raise X3Error(code, detail)
