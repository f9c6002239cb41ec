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
    """decodefile.rs:47-137 with the reference's memory behaviour: the file is STREAMED.  The reader holds one batch
    of whole frames at a time (X3_STREAM_CHUNK bytes of frame stream, 32 MiB by default, and their PCM), decodes it
    with one GPU call, and hands the frames out one at a time through decode_next_frame with the reference's
    semantics; the rest of a 118 GB batch of recordings is never in memory.  (The reference reads 24 KiB at a time,
    decodefile.rs:44; a GPU call needs tens of thousands of frames to fill the device.)

    Batches end on frame boundaries found by the reference's own walk -- header, payload_len, next header
    (decodefile.rs:105-126) -- so decoding batch by batch is decoding the file.  The last batch (the one that reaches
    the end of the file) goes through the whole-buffer logic below, which reproduces the reference's end-of-file
    quirks (remaing_bytes runs 8 bytes high, decodefile.rs:61-65)."""

    def __init__(self, filename, quiet=False, chunk_bytes=None):
        self._f = open(str(filename), "rb")            # File::open(...).unwrap(): a missing file raises
        self._size = os.fstat(self._f.fileno()).st_size
        head = self._f.read(28)
        if len(head) >= 28 and bytes(head[:8]) == x3.Archive.ID:
            try:
                plen = decoder.read_frame_header(np.frombuffer(head[8:28], dtype=np.uint8)).payload_len
            except error.X3Error:
                plen = 0
            head += self._f.read(plen)
        self._spec, header_size, off = read_archive_header(np.frombuffer(head, dtype=np.uint8), quiet=quiet)
        self._f.seek(off)
        self.remaing_bytes = self._size - header_size        # runs 8 bytes high, like the reference
        self.frame_errors = 0
        self._chunk = int(chunk_bytes or os.environ.get("X3_STREAM_CHUNK", 32 << 20))
        self._carry = np.empty(0, dtype=np.uint8)            # bytes read but not yet decoded (a partial frame)
        self._eof = False                                    # the file has been read to its end
        self._done = False                                   # the last batch has been decoded
        self._decoded = np.empty(0, dtype=np.int16)          # PCM of the current batch
        self._frame_sizes, self._starts_cache, self._next = [], [0], 0
        self._final_code, self._first_bad_code = error.OK, 0
        self._quiet = quiet

    @classmethod
    def open(cls, filename, quiet=False, chunk_bytes=None):
        return cls(filename, quiet=quiet, chunk_bytes=chunk_bytes)

    def spec(self):
        return self._spec

    def close(self):
        self._f.close()

    # -- one batch ---------------------------------------------------------------------------------
    def _walk(self, buf):
        """The reference's header walk over buf: (bytes of the whole frames, their sample counts, stopped)."""
        pos, sizes = 0, []
        while buf.size - pos > 20:
            try:
                h = decoder.read_frame_header(buf[pos:pos + 20])
            except error.X3Error:
                return pos, sizes, True        # the decode of this very frame reports the error: everything goes to it
            if buf.size - pos - 20 < h.payload_len:
                break                           # the frame continues in the part of the file not read yet
            sizes.append(h.samples)
            pos += 20 + h.payload_len
            if h.payload_len > X3_READ_BUFFER_SIZE:
                return pos, sizes, True        # FrameHeaderInvalidPayloadLen comes from the decode (decodefile.rs:118-121)
        return pos, sizes, False

    def _next_batch(self):
        """Decode the next batch of whole frames; False when the stream is finished."""
        if self._done:
            return False
        while True:
            if not self._eof:
                more = self._f.read(self._chunk)
                if len(more) < self._chunk:
                    self._eof = True
                buf = np.concatenate([self._carry, np.frombuffer(more, dtype=np.uint8)]) if self._carry.size else \
                    np.frombuffer(more, dtype=np.uint8)
            else:
                buf = self._carry
            if self._eof:
                self._finish(buf)
                return bool(self._frame_sizes) or self._final_code != error.OK or bool(self.frame_errors)
            consumed, sizes, stopped = self._walk(buf)
            if stopped:
                # a bad header or an oversized frame: hand over everything read so far, the decode reports it
                rest = self._f.read()
                self._eof = True
                self._finish(np.concatenate([buf, np.frombuffer(rest, dtype=np.uint8)]) if rest else buf)
                return True
            if consumed == 0:
                self._carry = buf                      # not one whole frame yet (never with chunks >= 32 KiB)
                continue
            pcm, res = decoder.decode_stream(buf[:consumed], self._spec.params, max_samples=sum(sizes))
            self._carry = buf[consumed:].copy()
            self._install(pcm, sizes[:res.frames], res.code, res)
            if res.code != error.OK or res.frame_errors:
                self._done = True                      # the reference stops at the first bad frame
            return True

    def _install(self, pcm, sizes, code, res):
        self._decoded, self._frame_sizes, self._next = pcm, list(sizes), 0
        self._starts_cache = np.concatenate([[0], np.cumsum(self._frame_sizes)]).astype(np.int64).tolist()
        self._final_code = code
        self.frame_errors += res.frame_errors
        self._first_bad_code = res.first_bad_code

    def _finish(self, frames):
        """The last batch: all remaining bytes, with the reference's end-of-file behaviour."""
        self._done = True
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
        self._install(pcm, sizes, code, res)

    # -- the reference's interface -----------------------------------------------------------------
    def decode_next_frame(self, wav_buf):
        """decodefile.rs:105-136: returns the number of samples written, or None at the end of the stream."""
        while self._next >= len(self._frame_sizes):
            if self._final_code != error.OK:
                code, self._final_code = self._final_code, error.OK
                self._done = True
                raise error.X3Error(code)
            if self._done or not self._next_batch():
                if self.frame_errors and not self._quiet and self._first_bad_code:
                    print("Frame error: %s" % error.X3Error(self._first_bad_code))   # decodefile.rs:132
                    self._first_bad_code = 0
                return None
        n = self._frame_sizes[self._next]
        start = self._starts_cache[self._next]
        wav_buf[:n] = self._decoded[start:start + n]
        self._next += 1
        return n

    def decode_batch(self):
        """All frames not handed out yet of the current batch (or of the next one) as one int16 array, or None at the end
        of the stream: what x3a_to_wav writes per GPU call.  Raises what decode_next_frame would raise."""
        while self._next >= len(self._frame_sizes):
            if self._final_code != error.OK:
                code, self._final_code = self._final_code, error.OK
                self._done = True
                raise error.X3Error(code)
            if self._done or not self._next_batch():
                if self.frame_errors and not self._quiet and self._first_bad_code:
                    print("Frame error: %s" % error.X3Error(self._first_bad_code))
                    self._first_bad_code = 0
                return None
        a, b = self._starts_cache[self._next], self._starts_cache[len(self._frame_sizes)]
        self._next = len(self._frame_sizes)
        return self._decoded[a:b]


def x3a_to_wav(x3a_filename, wav_filename, quiet=False):
    """decodefile::x3a_to_wav (decodefile.rs:189-212): frames are decoded and written batch by batch, so samples
    already decoded are in the file even when a later frame raises, as in the reference's streaming loop."""
    rd = X3aReader.open(x3a_filename, quiet=quiet)
    spec = rd.spec()
    try:
        with wave.open(str(wav_filename), "wb") as w:
            w.setnchannels(1)                          # decodefile.rs:195
            w.setsampwidth(2)
            w.setframerate(spec.sample_rate)
            while True:
                pcm = rd.decode_batch()                # raises what the reference would return, after the good samples
                if pcm is None:
                    break
                w.writeframes(pcm.astype("<i2", copy=False).tobytes())
    finally:
        rd.close()
