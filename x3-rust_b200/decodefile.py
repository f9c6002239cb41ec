"""Mirror of src/decodefile.rs: X3aReader and x3a_to_wav."""
import os
import re
import wave

import numpy as np

from . import decoder, error, x3

X3_READ_BUFFER_SIZE = 1024 * 24                 # decodefile.rs:44
X3_WRITE_BUFFER_SIZE = X3_READ_BUFFER_SIZE * 8  # decodefile.rs:45


def parse_xml(xml, quiet=False):
    """decodefile.rs:232-303: first FS / BLKLEN / CODES / T elements -> (sample_rate, Parameters)."""
    def first(tag):
        m = re.search(r"<%s(?:\s[^>]*)?>([^<]*)</%s>" % (tag, tag), xml)
        if m is None:
            raise error.X3Error(error.REFERENCE_PANIC, "missing <%s> (the reference indexes fs[0] and panics)" % tag)
        return m.group(1).strip()
    fs, bl, codes, th = first("FS"), first("BLKLEN"), first("CODES"), first("T")
    if not quiet:
        print("sample rate: %s" % fs)           # decodefile.rs:267-270
        print("block length: %s" % bl)
        print("Rice codes: %s" % codes)
        print("thresholds: %s" % th)
    ids = []
    for word in codes.split(","):
        if word in ("RICE0", "RICE1", "RICE2", "RICE3"):
            ids.append(int(word[4]))
        elif word == "BFP":
            pass
        else:
            raise error.X3Error(error.ARCHIVE_XML_RICE_CODE)
    ths = [int(s) for s in th.split(",")]
    params = x3.Parameters(int(bl), x3.Parameters.DEFAULT_BLOCKS_PER_FRAME, ids[:3], ths[:3])  # :290-299
    return int(fs), params


def read_archive_header(data, quiet=False):
    """decodefile.rs:142-176.  Returns (X3aSpec, header_size, frames_offset); header_size = 20 + xml length and
    does NOT include the 8-byte archive id, reproducing decodefile.rs:61-65,166."""
    if len(data) < 8:
        raise error.X3Error(error.IO)
    if bytes(data[:8]) != x3.Archive.ID:
        raise error.X3Error(error.ARCHIVE_INVALID_KEY)
    if len(data) < 28:
        raise error.X3Error(error.IO)
    h = decoder.read_frame_header(data[8:28])
    if len(data) < 28 + h.payload_len:
        raise error.X3Error(error.IO)
    xml = bytes(data[28:28 + h.payload_len]).decode("utf-8", errors="replace")
    fs, params = parse_xml(xml, quiet=quiet)
    return x3.X3aSpec(fs, params, h.channels), 20 + h.payload_len, 28 + h.payload_len


class X3aReader:
    """decodefile.rs:47-137.  The whole file is decoded on the GPU on first use; decode_next_frame then hands
    the frames out one at a time with the reference's semantics."""

    def __init__(self, filename, quiet=False):
        with open(str(filename), "rb") as f:       # File::open(...).unwrap(): a missing file raises
            self._data = np.frombuffer(f.read(), dtype=np.uint8)
        self._spec, header_size, self._off = read_archive_header(self._data, quiet=quiet)
        self.remaing_bytes = self._data.size - header_size   # runs 8 bytes high, like the reference
        self.frame_errors = 0
        self._decoded = None
        self._frame_sizes = None
        self._next = 0
        self._quiet = quiet

    @classmethod
    def open(cls, filename, quiet=False):
        return cls(filename, quiet=quiet)

    def spec(self):
        return self._spec

    def _ensure(self):
        if self._decoded is not None:
            return
        frames = self._data[self._off:]
        # the reference's end-of-file test uses remaing_bytes, which is 8 larger than the bytes really left:
        # a final frame cut short by fewer than 8 bytes passes the length guard and dies in read_exact (Io).
        pcm, res = decoder.decode_stream(frames, self._spec.params)
        sizes, pos = [], 0
        for _ in range(res.frames):
            h = decoder.read_frame_header(frames[pos:pos + 20])
            sizes.append(h.samples)
            pos += 20 + h.payload_len
        code = res.code
        if code == error.OK and res.frame_errors == 0 and frames.size - pos > 20 - 8:
            # what follows the last whole frame, seen through the reference's inflated counter
            rem = frames.size - pos + 8
            if rem > 20:
                if frames.size - pos < 20:
                    code = error.IO
                else:
                    try:
                        h = decoder.read_frame_header(frames[pos:pos + 20])
                        if rem - 20 >= h.payload_len and h.payload_len <= X3_READ_BUFFER_SIZE:
                            code = error.IO
                    except error.X3Error as e:
                        code = e.code
        self._decoded, self._frame_sizes, self._final_code = pcm, sizes, code
        self.frame_errors = res.frame_errors
        self._first_bad_code = res.first_bad_code

    def decode_next_frame(self, wav_buf):
        """decodefile.rs:105-136: returns the number of samples written, or None at the end of the stream."""
        self._ensure()
        if self._next < len(self._frame_sizes):
            n = self._frame_sizes[self._next]
            start = sum(self._frame_sizes[:self._next]) if self._next < 4 else self._starts()[self._next]
            wav_buf[:n] = self._decoded[start:start + n]
            self._next += 1
            return n
        if self._final_code != error.OK:
            code, self._final_code = self._final_code, error.OK
            raise error.X3Error(code)
        if self.frame_errors and not self._quiet and self._first_bad_code:
            print("Frame error: %s" % error.X3Error(self._first_bad_code))   # decodefile.rs:132
            self._first_bad_code = 0
        return None

    def _starts(self):
        if not hasattr(self, "_starts_cache"):
            self._starts_cache = np.concatenate([[0], np.cumsum(self._frame_sizes)]).tolist()
        return self._starts_cache


def x3a_to_wav(x3a_filename, wav_filename, quiet=False):
    """decodefile::x3a_to_wav (decodefile.rs:189-212).  Samples already decoded are written even when a later
    frame raises, as the reference's streaming loop does."""
    rd = X3aReader.open(x3a_filename, quiet=quiet)
    spec = rd.spec()
    rd._ensure()
    with wave.open(str(wav_filename), "wb") as w:
        w.setnchannels(1)                          # decodefile.rs:195
        w.setsampwidth(2)
        w.setframerate(spec.sample_rate)
        n = sum(rd._frame_sizes)
        w.writeframes(rd._decoded[:n].astype("<i2", copy=False).tobytes())
    rd._next = len(rd._frame_sizes)
    rd.decode_next_frame(np.empty(X3_WRITE_BUFFER_SIZE, dtype=np.int16))  # raises what the reference would return
