"""Mirror of src/bytewriter.rs: the output-buffer contract of the encoder."""
import io

from . import error


class SliceByteWriter:
    """bytewriter.rs:27-100: fixed-size buffer; overflow -> ByteWriterInsufficientMemory."""

    def __init__(self, buf):
        self.slice = memoryview(buf).cast("B")
        self.p_byte = 0
        self.stream_length = 0

    def align(self, n):
        residual = self.p_byte % n
        if residual == 0:
            return 0
        self.write_all(bytes(n - residual))
        return n - residual

    def write_all(self, value):
        value = bytes(value) if not isinstance(value, (bytes, bytearray, memoryview)) else value
        n = len(value)
        if n > len(self.slice) - self.p_byte:
            raise error.X3Error(error.BYTEWRITER_INSUFFICIENT_MEMORY)
        self.slice[self.p_byte:self.p_byte + n] = value
        self.p_byte += n
        self.stream_length = max(self.stream_length, self.p_byte)

    def capacity_left(self):
        return len(self.slice) - self.p_byte

    def flush(self):
        pass

    def stream_position(self):
        return self.p_byte

    def as_bytes(self):
        return bytes(self.slice[:self.stream_length])


class StreamByteWriter:
    """bytewriter.rs:115-164: any seekable binary stream."""

    def __init__(self, writer):
        self.writer = writer

    def align(self, n):
        residual = self.writer.tell() % n
        if residual == 0:
            return 0
        self.writer.write(bytes(n - residual))
        return n - residual

    def write_all(self, value):
        self.writer.write(value)

    def capacity_left(self):
        return None

    def flush(self):
        self.writer.flush()

    def stream_position(self):
        return self.writer.tell()
